import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
import oracle
from oracle.weights import make_state_dict
from ml_audio_restoration_b200 import RestorationPipeline, _lib
sds = {n: make_state_dict(n) for n in oracle.MODEL_NAMES}
pipe = RestorationPipeline.from_state_dicts(sds["denoiser"], sds["super_resolution"], sds["stereo"], "cuda")
need = C.c_size_t()
for B in (1, 1184, 1776, 2368):
    _lib.check(_lib.lib().ar_chain_workspace_bytes(pipe.chain(), B, 44100, C.byref(need)))
    print("chain workspace B=%d: %.1f MB per chunk, %.1f GB total" % (B, need.value / B / 1e6, need.value / 1e9))
