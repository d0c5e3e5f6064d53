"""Import-path shim: `from src.models import AudioDenoiser`, `from src.inference import restore_audio`
resolve to the B200-native implementations, exactly as they resolve to the PyTorch ones in the reference tree."""
