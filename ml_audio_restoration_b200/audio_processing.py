"""Audio helpers on the restoration path (reference: src/utils/audio_processing.py).

Only what inference needs: `normalize_audio` (device kernel, no host syncs), `chunk_audio`
(the reference's chunk vocabulary), and small WAV load/save helpers so `restore_audio` works
without `soundfile` (not installed in this image).
"""
from __future__ import annotations

import ctypes as C
import wave

import numpy as np
import torch

from . import _lib


def normalize_audio(audio: torch.Tensor, target_db: float = -20.0) -> torch.Tensor:
    """RMS-normalise to `target_db`, then peak-limit to 1.0 (audio_processing.py:58-87).

    CUDA tensors only: two stream-ordered kernels (reduce, scale) with the gain computed on the
    device, instead of the reference's `if rms == 0` / `if max_val > 1.0` host round trips.
    Returns a new tensor; a silent input comes back unchanged.
    """
    if not audio.is_cuda:
        raise RuntimeError("normalize_audio: input must be a CUDA tensor -- this build has no CPU fallback")
    out = audio.to(torch.float32).contiguous().clone()
    if out.numel() == 0:
        raise RuntimeError("normalize_audio: empty audio")
    with torch.cuda.device(out.device):
        scratch = torch.empty(_lib.NORMALIZE_SCRATCH_BYTES, dtype=torch.uint8, device=out.device)
        _lib.check(_lib.lib().ar_normalize(out.data_ptr(), out.numel(), float(target_db), scratch.data_ptr(),
                                           torch.cuda.current_stream(out.device).cuda_stream))
    return out


def chunk_audio(audio: torch.Tensor, chunk_size: int, overlap: int = 0) -> list:
    """Views of `audio[..., N]` of length `chunk_size` every `chunk_size - overlap` samples.

    Mirrors the reference helper (audio_processing.py:229-253) including its tail rule: when
    N is not a multiple of the stride one extra chunk holding the LAST `chunk_size` samples is
    appended (so a few samples can be dropped when N % stride == 0 but (N - chunk_size) % stride != 0,
    and a single short chunk comes back when N < chunk_size).  The batched GPU pipeline does
    not use this function; it uses the gap-free plan of `inference.plan_chunks`.
    """
    n = audio.shape[-1]
    stride = chunk_size - overlap
    starts = list(range(0, n - chunk_size + 1, stride))
    pieces = [audio[..., s:s + chunk_size] for s in starts]
    if n % stride != 0:
        pieces.append(audio[..., -chunk_size:])
    return pieces


# ----------------------------------------------------------------------------- minimal WAV I/O
def _read_wav(path: str):
    try:
        import soundfile as sf  # optional
        data, sr = sf.read(path, always_2d=True, dtype="float32")
        return torch.from_numpy(np.ascontiguousarray(data.T)), sr
    except ImportError:
        pass
    try:
        with wave.open(path, "rb") as w:
            sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n)
    except wave.Error:
        return _read_float_wav(path)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        a = v.astype(np.float32) / 8388608.0
    elif width == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise RuntimeError(f"unsupported WAV sample width {width}")
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1, nch).T)), sr


def _read_float_wav(path: str):
    """32-bit IEEE-float WAV (format tag 3), which the stdlib `wave` module rejects."""
    import struct
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise RuntimeError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(blob):
        cid, size = struct.unpack_from("<4sI", blob, pos)
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", blob, pos + 8)
        elif cid == b"data":
            data = blob[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    if fmt is None or data is None or fmt[0] != 3 or fmt[5] != 32:
        raise RuntimeError(f"{path}: unsupported WAV encoding")
    a = np.frombuffer(data, dtype="<f4").reshape(-1, fmt[1])
    return torch.from_numpy(np.array(a.T, dtype=np.float32, order="C")), fmt[2]


def load_audio(file_path: str, sample_rate: int = 22050, mono: bool = True):
    """(audio [C,N] float32, sample_rate) -- audio_processing.py:10-42 (host tensors, as the reference returns)."""
    audio, sr = _read_wav(file_path)
    if mono and audio.shape[0] > 1:
        audio = audio.mean(dim=0, keepdim=True)
    if sr != sample_rate:
        import torchaudio.functional as AF
        audio = AF.resample(audio, sr, sample_rate)
    return audio, sample_rate


def resample_mono_cuda(audio: torch.Tensor, sr: int, sample_rate: int = 22050) -> torch.Tensor:
    """`[C,N]` float32 CUDA -> `[1, ceil(sample_rate*N/sr)]`: the mono mix (`torch.mean(dim=0)`, audio_processing.py:33)
    fused with `torchaudio.transforms.Resample(sr, sample_rate)` (:38) in one kernel (`ar_resample_mono`)."""
    if not audio.is_cuda or audio.dim() != 2:
        raise RuntimeError("resample_mono_cuda: expected a [C,N] CUDA tensor -- this build has no CPU fallback")
    x = audio.to(torch.float32).contiguous()
    L = _lib.lib()
    n_out = C.c_int64()
    _lib.check(L.ar_resample_length(x.shape[1], int(sr), int(sample_rate), C.byref(n_out)))
    y = torch.empty((1, n_out.value), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.ar_resample_mono(x.data_ptr(), x.shape[0], x.shape[1], int(sr), int(sample_rate), y.data_ptr(),
                                      n_out.value, torch.cuda.current_stream(x.device).cuda_stream))
    return y


def load_audio_cuda(file_path: str, sample_rate: int = 22050, device="cuda"):
    """`load_audio(..., mono=True)` with the arithmetic on the GPU: 16-bit PCM files are uploaded as raw
    int16 frames (half the H2D bytes of float32) from pinned memory, decoded, mixed to mono and resampled on the device.
    Returns (audio [1,N] float32 CUDA, sample_rate)."""
    dev = torch.device(device)
    with wave.open(file_path, "rb") as w:
        sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n) if width == 2 else None
    if raw is None:                                    # other encodings: host decode, device mix + resample
        audio, sr = _read_wav(file_path)
        return resample_mono_cuda(audio.to(dev), sr, sample_rate), sample_rate
    pcm = torch.frombuffer(bytearray(raw), dtype=torch.int16).pin_memory().to(dev, non_blocking=True)
    planar = torch.empty((nch, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ar_pcm16_to_float(pcm.data_ptr(), nch, n, planar.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream))
    return resample_mono_cuda(planar, sr, sample_rate), sample_rate


def save_audio(file_path: str, audio: torch.Tensor, sample_rate: int = 22050, encoding: str = "float32") -> None:
    """Write `[C,N]` float audio as WAV (audio_processing.py:45-55).  The reference's `torchaudio.save` of a float
    tensor writes 32-bit IEEE-float samples; `encoding="pcm16"` writes 16-bit PCM instead."""
    import os
    import struct
    a = audio.detach().to("cpu", torch.float32).numpy()
    if a.ndim == 1:
        a = a[None]
    parent = os.path.dirname(os.path.abspath(file_path))
    os.makedirs(parent, exist_ok=True)
    if encoding == "pcm16":
        pcm = np.round(np.clip(a, -1.0, 1.0).T * 32767.0).astype("<i2")
        with wave.open(file_path, "wb") as w:
            w.setnchannels(a.shape[0])
            w.setsampwidth(2)
            w.setframerate(int(sample_rate))
            w.writeframes(pcm.tobytes())
        return
    if encoding != "float32":
        raise ValueError(f"unknown WAV encoding {encoding!r}")
    data = np.ascontiguousarray(a.T).astype("<f4").tobytes()
    nch, sr = a.shape[0], int(sample_rate)
    fmt = struct.pack("<4sIHHIIHH", b"fmt ", 16, 3, nch, sr, sr * nch * 4, nch * 4, 32)      # format tag 3 = IEEE float
    fact = struct.pack("<4sII", b"fact", 4, a.shape[1])
    body = b"WAVE" + fmt + fact + struct.pack("<4sI", b"data", len(data)) + data
    with open(file_path, "wb") as f:
        f.write(struct.pack("<4sI", b"RIFF", len(body)) + body)
