import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def state_dicts():
    from oracle.weights import make_state_dict
    return {n: make_state_dict(n, 1234) for n in ("denoiser", "super_resolution", "stereo")}


def snr_db(ref, got):
    import torch
    ref = torch.as_tensor(ref).double()
    got = torch.as_tensor(got).double()
    num = (ref ** 2).sum()
    den = ((ref - got) ** 2).sum()
    if den == 0:
        return float("inf")
    return float(10 * torch.log10(num / den))
