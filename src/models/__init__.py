"""Audio restoration models (same import surface as the reference's src/models/__init__.py)."""
from ml_audio_restoration_b200.models import AudioDenoiser, StereoSeparator, AudioSuperResolution

__all__ = ['AudioDenoiser', 'StereoSeparator', 'AudioSuperResolution']
