from ml_audio_restoration_b200.models.super_resolution import AudioSuperResolution  # noqa: F401
