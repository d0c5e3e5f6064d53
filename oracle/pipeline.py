"""CPU restatement of the pipeline around the models (test infrastructure).

  * `normalize_audio`  -- audio_processing.py:58-87 (pinned by golden vectors).
  * `restore_whole`    -- the tensor part of inference.py:45-98 (whole file, batch 1; pinned
                          because it is just normalize + the three pinned forwards).
  * `plan_chunks` / `crossfade_window` / `restore_chunked` -- THIS repo's chunk -> batch ->
    overlap-add scheme.  PARITY UNPINNED by the reference (no such function exists there,
    SURVEY.md D3/D4).  Its vocabulary is `chunk_audio(audio, chunk_size, overlap)`
    (audio_processing.py:229-253) and its tail handling is the zero-pad of
    `Trainer.generate_test_output` (trainer.py:656-665); with overlap=0 it is exactly that
    loop (non-overlapping chunks, zero-padded tail, per-chunk LSTM reset, concat, trim).

Scheme (all chunks have the same length so the GPU can batch them):
    hop      = chunk_size - overlap                       (overlap <= chunk_size // 2)
    n_chunks = 1 if N <= chunk_size else ceil((N - overlap) / hop)
    chunk i  = x[i*hop : i*hop + chunk_size], zero-padded on the right
    y_i      = stereo(super_res(denoise(chunk i)))        (state reset per chunk)
    window   : w_i[j] = 1, except a linear ramp-up  (j + .5)/(r*overlap) over the first
               r*overlap output samples when i > 0 and the complementary ramp-down over the
               last r*overlap when i < n-1 (r = output/input rate, 2 with super-res)
    out[p]   = sum_i w_i[p - r*i*hop] * y_i[p - r*i*hop],  p < r*N
"""
from __future__ import annotations

import math

import torch

from .models import denoiser_forward, super_resolution_forward, stereo_forward, stereo_forward_window

DEFAULT_CHUNK = 44100      # 2.0 s @ 22.05 kHz (trainer.py:652)
DEFAULT_OVERLAP = 2052     # hop 42048 = 0 mod 8 (denoiser pools)


def normalize_audio(audio: torch.Tensor, target_db: float = -20.0) -> torch.Tensor:
    """RMS -> target dB over ALL elements, then peak-limit to 1.0 (audio_processing.py:58-87)."""
    rms = torch.sqrt(torch.mean(audio ** 2))
    if rms == 0:
        return audio
    gain = (10 ** (target_db / 20)) / rms
    out = audio * gain
    peak = out.abs().max()
    if peak > 1.0:
        out = out / peak
    return out


def chain_forward(sds, x, enable_super_resolution=True, dtype=torch.float32):
    """denoise -> (super-res) -> stereo on a batch of chunks [B,1,T] -> [B,2,rT] (inference.py:59-95)."""
    y = denoiser_forward(sds["denoiser"], x, dtype)
    if enable_super_resolution:
        y = super_resolution_forward(sds["super_resolution"], y, dtype)
    return stereo_forward(sds["stereo"], y, dtype)


def restore_whole(sds, audio_1n, enable_super_resolution=True, dtype=torch.float32):
    """inference.py:45-98 without file I/O: [1,N] mono -> [2,rN]."""
    a = normalize_audio(audio_1n.to(dtype))
    y = chain_forward(sds, a.unsqueeze(0), enable_super_resolution, dtype).squeeze(0)
    return normalize_audio(y)


def plan_chunks(num_samples: int, chunk_size: int = DEFAULT_CHUNK, overlap: int = DEFAULT_OVERLAP):
    """Chunk start offsets (input-rate samples).  See module docstring."""
    if not (0 <= overlap <= chunk_size // 2):
        raise ValueError("overlap must be in [0, chunk_size // 2]")
    if num_samples <= 0:
        raise ValueError("empty audio")
    hop = chunk_size - overlap
    n = 1 if num_samples <= chunk_size else math.ceil((num_samples - overlap) / hop)
    return [i * hop for i in range(n)]


def crossfade_window(i: int, n_chunks: int, chunk_size: int, overlap: int, rate: int) -> torch.Tensor:
    """fp32 window w_i over the r*chunk_size output samples of chunk i."""
    L, V = rate * chunk_size, rate * overlap
    w = torch.ones(L, dtype=torch.float32)
    if V > 0:
        inv = torch.tensor(1.0 / V, dtype=torch.float32)
        ramp = (torch.arange(V, dtype=torch.float32) + 0.5) * inv
        if i > 0:
            w[:V] = ramp
        if i < n_chunks - 1:
            w[L - V:] = 1.0 - ramp
    return w


def split_chunks(audio_1n: torch.Tensor, chunk_size: int, overlap: int) -> torch.Tensor:
    """[1,N] -> [n_chunks,1,chunk_size] with the tail zero-padded (trainer.py:656-665)."""
    N = audio_1n.shape[-1]
    starts = plan_chunks(N, chunk_size, overlap)
    out = audio_1n.new_zeros(len(starts), 1, chunk_size)
    for i, s in enumerate(starts):
        seg = audio_1n[0, s:s + chunk_size]
        out[i, 0, :seg.numel()] = seg
    return out


def stitch_chunks(y: torch.Tensor, num_samples: int, chunk_size: int, overlap: int, rate: int) -> torch.Tensor:
    """[n_chunks,C,r*chunk_size] -> [C, r*N] weighted overlap-add."""
    n, C, L = y.shape
    hop = chunk_size - overlap
    total = rate * ((n - 1) * hop + chunk_size)
    out = y.new_zeros(C, total)
    for i in range(n):
        w = crossfade_window(i, n, chunk_size, overlap, rate).to(y.dtype)
        out[:, rate * i * hop: rate * i * hop + L] += y[i] * w
    return out[:, :rate * num_samples]


def restore_chunked(sds, audio_1n, chunk_size=DEFAULT_CHUNK, overlap=DEFAULT_OVERLAP,
                    enable_super_resolution=True, dtype=torch.float32, batch=8, normalize=True):
    """Chunked chain with overlap-add stitching: [1,N] mono -> [2,rN]."""
    a = audio_1n.to(dtype)
    if normalize:
        a = normalize_audio(a)
    chunks = split_chunks(a, chunk_size, overlap)
    outs = [chain_forward(sds, chunks[b:b + batch], enable_super_resolution, dtype)
            for b in range(0, chunks.shape[0], batch)]
    rate = 2 if enable_super_resolution else 1
    y = stitch_chunks(torch.cat(outs, 0), a.shape[-1], chunk_size, overlap, rate)
    return normalize_audio(y) if normalize else y


# mode="exact": same constants as ml_audio_restoration_b200/inference.py (receptive fields: SURVEY.md App. E)
EXACT_HALO, EXACT_STEREO_HALO, EXACT_LSTM_LEAD = 96, 40, 16


def restore_exact(sds, audio_1n, chunk_size=DEFAULT_CHUNK, enable_super_resolution=True, dtype=torch.float32, normalize=True):
    """Whole-file-exact chunked restoration, restated from the oracle forwards: chunks whose starts are
    multiples of 8 with EXACT_HALO discarded input samples per interior edge for denoise -> super-res, then stereo
    segments with conv halos and the LSTM (h, c) carried from segment to segment.  Unlike `restore_chunked` this scheme IS
    pinned by the reference: its result must equal `restore_whole` (inference.py:59-95 on the whole file) to rounding."""
    a = audio_1n.to(dtype)
    if normalize:
        a = normalize_audio(a)
    N, H = a.shape[-1], EXACT_HALO
    r = 2 if enable_super_resolution else 1
    L = chunk_size - chunk_size % 8
    if L < 4 * H:
        raise ValueError("chunk_size too small for mode='exact'")
    hop = L - 2 * H
    n = 1 if N <= L else -(-(N - 2 * H) // hop)
    sig = a.new_zeros(1, r * N)
    for i in range(n):
        s0, s1 = i * hop, min(N, i * hop + L)
        z = denoiser_forward(sds["denoiser"], a[:, s0:s1].unsqueeze(0), dtype)
        if enable_super_resolution:
            z = super_resolution_forward(sds["super_resolution"], z, dtype)
        lo = 0 if i == 0 else s0 + H
        hi = N if i == n - 1 else s0 + hop + H
        sig[:, r * lo:r * hi] = z[0, :, r * (lo - s0):r * (hi - s0)]
    M, E, P = r * N, EXACT_STEREO_HALO, EXACT_LSTM_LEAD
    S2 = (r * L - 2 * E) // 16 * 16
    out = a.new_zeros(2, M)
    state = None
    n_seg = -(-M // S2)
    for j in range(n_seg):
        a0, a1 = j * S2, min(M, (j + 1) * S2)
        e0, e1 = max(0, a0 - E), min(M, a1 + E)
        y, state = stereo_forward_window(sds["stereo"], sig[:, e0:e1].unsqueeze(0), state,
                                         lstm_start=0 if j == 0 else a0 - P - e0,
                                         state_pos=None if j == n_seg - 1 else a1 - P - e0, dtype=dtype)
        out[:, a0:a1] = y[0, :, a0 - e0:a1 - e0]
    return normalize_audio(out) if normalize else out
