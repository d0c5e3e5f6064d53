"""CPU oracle for the restoration chain -- TEST INFRASTRUCTURE ONLY.

This package restates, op by op, the reference's PyTorch forward passes
(`/root/reference/src/models/*.py`, `src/utils/audio_processing.py::normalize_audio`)
and this repo's own chunk/stitch scheme, on the CPU.  It is the checker for the
CUDA path: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  Nothing under
`ml_audio_restoration_b200/` or `src/` imports it, and the product path raises when the
CUDA library is missing rather than falling back to this code.

Parity status
  * model forwards, normalize_audio: PINNED -- checked against the reference modules
    imported from /root/reference (tests/golden/make_golden.py) and the committed
    golden vectors under tests/golden/.
  * load_audio front end (PCM16 decode, mono mix, torchaudio sinc resampler -- oracle/audio_io.py): PINNED against
    golden vectors generated from the installed torchaudio 2.11 (tests/golden/make_golden_io.py); the dependency is
    not vendored in the reference (requirements.txt: torchaudio>=2.0.0), its published algorithm is restated.
  * chunk -> batch -> overlap-add stitching: PARITY UNPINNED by the reference (it has no
    such function); the oracle restates THIS repo's scheme from the
    oracle model forwards.  With overlap=0 it reduces to the reference's
    `Trainer.generate_test_output` loop (trainer.py:652-681), which is pinned.
  * whole-file-exact chunking (`restore_exact`: conv halos, chunk starts = 0 mod 8, LSTM state carried): PINNED -- it
    must reproduce `restore_whole`, i.e. the reference's whole-file forwards (inference.py:59-95), and the golden
    `chain_whole` vector.
"""
from .weights import make_state_dict, MODEL_NAMES  # noqa: F401
from .models import denoiser_forward, super_resolution_forward, stereo_forward, stereo_forward_window  # noqa: F401
from .pipeline import (normalize_audio, chain_forward, restore_chunked, restore_whole, restore_exact, plan_chunks,  # noqa: F401
                       crossfade_window)
