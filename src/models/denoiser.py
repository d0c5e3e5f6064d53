from ml_audio_restoration_b200.models.denoiser import AudioDenoiser  # noqa: F401
