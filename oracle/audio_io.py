"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the front end of `load_audio`
(reference: src/utils/audio_processing.py:10-42): PCM16 -> float (soundfile's float32 read, :24), mono mix
(`torch.mean(dim=0, keepdim=True)`, :33) and `torchaudio.transforms.Resample(sr, sample_rate)` (:38).

The resampler is a third-party dependency that is not vendored in the reference (requirements.txt pins only
`torchaudio>=2.0.0`; installed here: torchaudio 2.11.0).  This file restates its published algorithm
(`torchaudio.functional.functional._get_sinc_resample_kernel` / `_apply_sinc_resample_kernel`, method
"sinc_interp_hann", lowpass_filter_width 6, rolloff 0.99 -- the defaults `Resample` uses) in numpy and is pinned by
golden vectors generated from the installed torchaudio (tests/golden/make_golden_io.py -> golden_io_v1.npz).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this package.
"""
import math

import numpy as np


def sinc_resample_taps(orig_sr: int, new_sr: int):
    """(taps fp32 [new/g][2*width + orig/g], width, orig/g, new/g) -- float64 math, rounded to fp32 at the end."""
    g = math.gcd(int(orig_sr), int(new_sr))
    o, n = int(orig_sr) // g, int(new_sr) // g
    base = min(o, n) * 0.99
    width = math.ceil(6 * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    # torchaudio divides the integer phase index by new_freq in the default dtype (float32) and only then promotes to
    # float64 (`torch.arange(0, -new_freq, -1, dtype=None) / new_freq + idx`): restated as is, it moves taps by up to 1e-5
    phase = (np.arange(0, -n, -1).astype(np.float32) / np.float32(n)).astype(np.float64)
    t = (phase[:, None] + idx) * base
    t = np.clip(t, -6.0, 6.0)
    window = np.cos(t * math.pi / 6 / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / o)
    return k.astype(np.float32), width, o, n


def mono_mix(x: np.ndarray) -> np.ndarray:
    """[C,N] -> [1,N], fp32 mean over channels (audio_processing.py:32-33)."""
    x = np.asarray(x, dtype=np.float32)
    return x if x.shape[0] == 1 else (x.sum(axis=0, dtype=np.float32) / np.float32(x.shape[0]))[None]


def resample(x: np.ndarray, orig_sr: int, new_sr: int) -> np.ndarray:
    """[C,N] fp32 -> [C, ceil(new*N/orig)] fp32; identity when the rates match (as torchaudio does)."""
    x = np.asarray(x, dtype=np.float32)
    if int(orig_sr) == int(new_sr):
        return x
    taps, width, o, n = sinc_resample_taps(orig_sr, new_sr)
    C, N = x.shape
    xp = np.pad(x, ((0, 0), (width, width + o)))
    K = taps.shape[1]
    frames = (xp.shape[1] - K) // o + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, K, axis=1)[:, ::o][:, :frames]     # [C, frames, K]
    # fp32 taps and samples, products accumulated in float64 and rounded once: the checker is then at least as accurate
    # as any fp32 summation order (torch's conv1d, the CUDA kernel's serial fmaf chain) it is compared with
    y = np.einsum("cfk,pk->cfp", win.astype(np.float64), taps.astype(np.float64)).astype(np.float32).reshape(C, -1)
    target = -(-n * N // o)
    return np.ascontiguousarray(y[:, :target])


def pcm16_to_float(pcm: np.ndarray) -> np.ndarray:
    """interleaved int16 frames [N, C] -> planar fp32 [C, N] in [-1, 1) (soundfile dtype='float32')."""
    return np.ascontiguousarray((np.asarray(pcm, dtype=np.int16).astype(np.float32) / np.float32(32768.0)).T)


def load_front_end(x: np.ndarray, sr: int, sample_rate: int = 22050, mono: bool = True) -> np.ndarray:
    """mono mix then resample, in the reference's order (audio_processing.py:32-39)."""
    if mono and x.shape[0] > 1:
        x = mono_mix(x)
    return resample(x, sr, sample_rate)
