from ml_audio_restoration_b200.audio_processing import normalize_audio, chunk_audio, load_audio, save_audio  # noqa: F401
