"""Audio restoration models (drop-in for the reference's `src.models`, src/models/__init__.py:2-6)."""
from .denoiser import AudioDenoiser
from .stereo_separator import StereoSeparator
from .super_resolution import AudioSuperResolution

__all__ = ["AudioDenoiser", "StereoSeparator", "AudioSuperResolution"]
