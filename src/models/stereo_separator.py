from ml_audio_restoration_b200.models.stereo_separator import StereoSeparator  # noqa: F401
