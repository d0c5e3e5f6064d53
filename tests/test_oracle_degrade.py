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


def test_lfilter_delegation_matches_a_plain_python_recurrence():
    """The oracle delegates the serial direct-form-II-transposed recurrence to scipy.signal.lfilter; a pure-Python loop of
    the same recurrence (small case) pins that delegation, state hand-over included."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal(60)
    for order, wn, btype in [(4, 100 / 11025, "low"), (3, 7000 / 11025, "low"), (4, 2500 / 11025, "high")]:
        b, a = degrade.butter(order, wn, btype)
        zi = degrade.lfilter_zi(b, a) * x[0]
        y_ref, z_ref = signal.lfilter(b, a, x, zi=zi)
        z = list(zi)
        y = []
        for xn in x:
            yn = z[0] + b[0] * xn
            for i in range(order - 1):
                z[i] = z[i + 1] + xn * b[i + 1] - yn * a[i + 1]
            z[order - 1] = xn * b[order] - yn * a[order]
            y.append(yn)
        np.testing.assert_allclose(np.array(y), y_ref, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(np.array(z), z_ref, rtol=1e-9, atol=1e-15)


def test_pop_impulse_shape():
    """One pop: exponential decay from amp*polarity with the 0.3*tau time constant, ringing only beyond 10 samples
    (audio_processing.py:171-186)."""
    pop = {"loc": 0, "amp": 0.4, "polarity": -1, "decay_time": 0.002, "length": 44, "resonance_freq": None}
    imp = degrade.pop_impulse(pop, 22050)
    assert imp[0] == pytest.approx(-0.4) and np.all(np.diff(imp) > 0) and len(imp) == 44
    assert imp[10] == pytest.approx(-0.4 * np.exp(-10 / (22050 * 0.002 * 0.3)))
    pop["resonance_freq"] = 5000.0
    ring = degrade.pop_impulse(pop, 22050) - imp
    k = np.arange(44)
    np.testing.assert_allclose(ring, 0.3 * np.sin(2 * np.pi * 5000.0 * k / 22050) * np.exp(-k / (22050 * 0.002 * 0.3)) * 0.4 * 0.2,
                               rtol=1e-12, atol=1e-18)
