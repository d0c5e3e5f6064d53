"""CPU: the oracle's restatement of `simulate_vinyl_artifacts` (audio_processing.py:122-226) against golden vectors
produced by the unmodified reference function (tests/golden/make_golden_degrade.py), and its restated scipy pieces
(`butter`, `lfilter_zi`, `filtfilt`) against the installed scipy."""
import os
import sys

import numpy as np
import pytest
import torch
from scipy import signal

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_degrade import CASES, make_input  # noqa: E402
from oracle import degrade  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_degrade_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_reference_golden(gold, i):
    seed, C, N, sr, kw = CASES[i]
    np.random.seed(seed)
    torch.manual_seed(seed)
    y = degrade.simulate_vinyl_artifacts(torch.from_numpy(make_input(seed, C, N)), sr, **kw).numpy()
    ref = gold[f"case{i}_y"]
    assert y.shape == ref.shape == (C, N) and y.dtype == np.float32
    # same draws, same float32 / float64 arithmetic; only butter's last-ulp coefficient differences remain
    assert np.abs(y - ref).max() <= 1e-7


@pytest.mark.parametrize("order,wn,btype", [(4, 2500 / 11025, "high"), (4, 100 / 11025, "low"), (3, 7000 / 11025, "low"),
                                            (3, 6000 / 22050, "low"), (4, 2500 / 22050, "high"), (2, 0.5, "low"),
                                            (1, 0.3, "high")])
def test_butter_matches_scipy(order, wn, btype):
    b, a = degrade.butter(order, wn, btype)
    bs, as_ = signal.butter(order, wn, btype=btype)
    np.testing.assert_allclose(b, bs, rtol=1e-11, atol=0)
    np.testing.assert_allclose(a, as_, rtol=1e-11, atol=0)
    np.testing.assert_allclose(degrade.lfilter_zi(bs, as_), signal.lfilter_zi(bs, as_), rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("n", [16, 17, 100, 4410])
def test_filtfilt_matches_scipy(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    for order, wn, btype in [(4, 2500 / 11025, "high"), (4, 100 / 11025, "low"), (3, 7000 / 11025, "low")]:
        if n <= 3 * (order + 1):
            continue
        b, a = signal.butter(order, wn, btype=btype)
        y = degrade.filtfilt(b, a, x)
        ref = signal.filtfilt(b, a, x)
        assert y.dtype == np.float64
        np.testing.assert_allclose(y, ref, rtol=1e-9, atol=1e-12)


def test_filtfilt_rejects_short_input():
    b, a = signal.butter(4, 0.2)
    with pytest.raises(ValueError):
        degrade.filtfilt(b, a, np.zeros(15, dtype=np.float32))


def test_plan_consumes_numpy_generator_like_the_reference():
    # pops near the end are truncated; pops of <= 10 samples draw no resonance frequency (audio_processing.py:171-182)
    np.random.seed(3)
    plan = degrade.draw_plan(500, 22050, impulse_rate=300.0)
    assert all(0 < p["length"] <= 500 - p["loc"] for p in plan["pops"])
    assert any(p["resonance_freq"] is None for p in plan["pops"]) or all(p["length"] > 10 for p in plan["pops"])
    assert 0.015 <= plan["surface_level"] <= 0.03 and 6000 <= plan["rolloff_hz"] <= 8000
