"""B200-native (sm_100a) inference chain of ml-audio-restoration: drop-in models + pipeline.

Host side is Python/PyTorch plumbing (parameters, device memory, streams); all compute runs in
`libaudiorestore_sm100.so` (csrc/), reached through the C-ABI of include/audiorestore.h.
"""
from .models import AudioDenoiser, AudioSuperResolution, StereoSeparator  # noqa: F401
from .inference import (RestorationPipeline, restore_audio, plan_chunks, shard_range, restore_sharded,  # noqa: F401
                        chunked_model_eval, generate_test_output)
from .audio_processing import normalize_audio, chunk_audio, load_audio, save_audio  # noqa: F401

__all__ = ["AudioDenoiser", "AudioSuperResolution", "StereoSeparator", "RestorationPipeline", "restore_audio",
           "plan_chunks", "shard_range", "restore_sharded", "normalize_audio", "chunk_audio", "load_audio", "save_audio"]
