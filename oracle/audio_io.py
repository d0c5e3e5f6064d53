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


PCM_U8, PCM_S16, PCM_S24, PCM_S32, PCM_F32, PCM_F64 = 1, 2, 3, 4, 5, 6     # AR_PCM_* of include/audiorestore.h


def pcm_to_float(raw: bytes, fmt: int, channels: int) -> np.ndarray:
    """The sample bytes of a WAV `data` chunk (interleaved little-endian frames) -> planar fp32 [C, N], scaled as
    soundfile's dtype='float32' read scales them (audio_processing.py:24): integers by 2^(bits-1), unsigned 8-bit with a
    bias of 128, floats unchanged (float64 rounded to nearest)."""
    if fmt == PCM_U8:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif fmt == PCM_S16:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    elif fmt == PCM_S24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        a = v.astype(np.float32) / np.float32(8388608.0)
    elif fmt == PCM_S32:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / np.float32(2147483648.0)
    elif fmt == PCM_F32:
        a = np.frombuffer(raw, dtype="<f4").astype(np.float32)
    elif fmt == PCM_F64:
        a = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    else:
        raise ValueError(f"unknown sample format {fmt}")
    return np.ascontiguousarray(a.reshape(-1, channels).T)


def write_wav(path: str, samples: np.ndarray, fmt: int, sample_rate: int, extensible: bool = False, junk: bool = False) -> bytes:
    """Test helper: `samples` [N, C] (already of the encoding's numpy type; for PCM_S24 int32 values in [-2^23, 2^23)) as a
    RIFF/WAVE file with format tag 1 / 3 or WAVE_FORMAT_EXTENSIBLE, optionally with an odd-sized LIST chunk in front of
    `data`.  Returns the bytes of the data chunk."""
    import struct
    n, ch = samples.shape
    bits = {PCM_U8: 8, PCM_S16: 16, PCM_S24: 24, PCM_S32: 32, PCM_F32: 32, PCM_F64: 64}[fmt]
    tag = 3 if fmt in (PCM_F32, PCM_F64) else 1
    if fmt == PCM_S24:
        v = samples.astype(np.int32) & 0xFFFFFF
        data = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=-1).astype(np.uint8).tobytes()
    else:
        dt = {PCM_U8: np.uint8, PCM_S16: "<i2", PCM_S32: "<i4", PCM_F32: "<f4", PCM_F64: "<f8"}[fmt]
        data = np.ascontiguousarray(samples).astype(dt).tobytes()
    align = ch * bits // 8
    if extensible:
        guid = struct.pack("<H", tag) + bytes.fromhex("000000001000800000aa00389b71")
        fmt_body = struct.pack("<HHIIHHHHI", 0xFFFE, ch, sample_rate, sample_rate * align, align, bits, 22, bits, 0) + guid
    else:
        fmt_body = struct.pack("<HHIIHH", tag, ch, sample_rate, sample_rate * align, align, bits)
    chunks = struct.pack("<4sI", b"fmt ", len(fmt_body)) + fmt_body
    if junk:
        chunks += struct.pack("<4sI", b"LIST", 5) + b"abcde" + b"\0"          # odd size: padded to even
    chunks += struct.pack("<4sI", b"data", len(data)) + data + (b"\0" if len(data) & 1 else b"")
    with open(path, "wb") as f:
        f.write(struct.pack("<4sI", b"RIFF", 4 + len(chunks)) + b"WAVE" + chunks)
    return data


def load_front_end(x: np.ndarray, sr: int, sample_rate: int = 22050, mono: bool = True) -> np.ndarray:
    """mono mix then resample, in the reference's order (audio_processing.py:32-39)."""
    if mono and x.shape[0] > 1:
        x = mono_mix(x)
    return resample(x, sr, sample_rate)
