"""Golden vectors for the load_audio front end, generated from the INSTALLED torchaudio (the reference's dependency,
audio_processing.py:38 `torchaudio.transforms.Resample(sr, sample_rate)`) in the build container:

    python tests/golden/make_golden_io.py      ->  tests/golden/golden_io_v1.npz

Inputs are regenerated from the seeds stored in the file (numpy default_rng), outputs are torchaudio's."""
import os

import numpy as np
import torch
import torchaudio

CASES = [(44100, 22050, 1, 3001), (48000, 22050, 2, 2500), (16000, 22050, 1, 1500), (8000, 22050, 1, 700),
         (22050, 44100, 1, 900), (32000, 22050, 2, 1234), (44100, 22050, 2, 64)]


def make_input(seed, C, N):
    return (0.3 * np.random.default_rng(seed).standard_normal((C, N))).astype(np.float32)


def main():
    out = {"torchaudio_version": np.array(torchaudio.__version__)}
    for i, (orig, new, C, N) in enumerate(CASES):
        x = torch.from_numpy(make_input(100 + i, C, N))
        xm = x.mean(dim=0, keepdim=True) if C > 1 else x
        y = torchaudio.transforms.Resample(orig, new)(xm)
        out[f"case{i}_meta"] = np.array([orig, new, C, N, 100 + i], dtype=np.int64)
        out[f"case{i}_y"] = y.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_io_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
