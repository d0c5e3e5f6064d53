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
    with wave.open(path, "rb") as w:
        sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
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


def load_audio(file_path: str, sample_rate: int = 22050, mono: bool = True):
    """(audio [C,N] float32, sample_rate) -- audio_processing.py:10-42."""
    audio, sr = _read_wav(file_path)
    if mono and audio.shape[0] > 1:
        audio = audio.mean(dim=0, keepdim=True)
    if sr != sample_rate:
        import torchaudio.functional as AF
        audio = AF.resample(audio, sr, sample_rate)
    return audio, sample_rate


def save_audio(file_path: str, audio: torch.Tensor, sample_rate: int = 22050) -> None:
    """Write `[C,N]` float audio as 16-bit PCM WAV (audio_processing.py:45-55)."""
    a = audio.detach().to("cpu", torch.float32).clamp(-1.0, 1.0).numpy()
    if a.ndim == 1:
        a = a[None]
    pcm = np.round(a.T * 32767.0).astype("<i2")
    with wave.open(file_path, "wb") as w:
        w.setnchannels(a.shape[0])
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(pcm.tobytes())
