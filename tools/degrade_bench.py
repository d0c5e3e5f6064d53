"""Times the degradation generator on one GPU: CUDA events around `apply_vinyl_artifacts` and around
one `filtfilt`, for a batch of 2 s training chunks and for one 3-minute side; prints samples/s and the HBM rate on the
bytes the six filtfilt launches actually move (fwd state 4 B, fwd emit 4 + 8 B, bwd state 8 B, bwd emit 8 + 4 B = 36 B per
sample; the algorithmic minimum is 8 B).  Also times scipy's filtfilt on the host for the same row (1 core)."""
import json
import time

import numpy as np
import torch
from scipy import signal

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_audio_restoration_b200 import audio_processing as ap  # noqa: E402


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = []
    for rows, n, sr in [(256, 44100, 22050), (1, 3_969_000, 22050), (64, 3_969_000, 22050)]:
        x = 0.1 * torch.randn(rows, n, device="cuda")
        noise = [torch.randn_like(x) for _ in range(3)]
        np.random.seed(0)
        plan = ap.plan_vinyl_artifacts(n, sr)
        b, a = ap.butter(4, 100 / (sr / 2), "low")
        ms_f = timed(lambda: ap.filtfilt(b, a, x))
        ms_all = timed(lambda: ap.apply_vinyl_artifacts(x, sr, plan, *noise))
        t0 = time.perf_counter()
        signal.filtfilt(b, a, x[0].cpu().numpy())
        cpu_row_s = time.perf_counter() - t0
        out.append({"rows": rows, "n": n, "pops": int(len(plan["pops"])), "filtfilt_ms": ms_f,
                    "filtfilt_Msamples_per_s": rows * n / ms_f / 1e3, "filtfilt_GBps_moved": 36.0 * rows * n / ms_f / 1e6,
                    "generator_ms": ms_all, "generator_audio_s_per_s": rows * n / sr / (ms_all / 1e3),
                    "scipy_filtfilt_one_row_ms": cpu_row_s * 1e3,
                    "scipy_filtfilt_Msamples_per_s_1core": n / cpu_row_s / 1e6})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
