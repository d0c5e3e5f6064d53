"""GPU: the degradation generator through the C-ABI against the CPU oracle and the golden vectors
produced by the unmodified reference function (audio_processing.py:122-226).

Tolerance: the reference's pipeline is float32 with float64 filter passes rounded back to float32; the GPU runs the same
operations in the same precision, re-associated only inside the block-parallel IIR scan (float64).  Asserted: max abs
error <= 1e-6 on signals of amplitude ~0.1-1 (a few float32 ulps)."""
import os
import sys

import numpy as np
import pytest
import torch
from scipy import signal

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_degrade import CASES, make_input  # noqa: E402
from oracle import degrade  # noqa: E402
from ml_audio_restoration_b200 import audio_processing as ap  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_degrade_v1.npz")
TOL = 1e-6
FILTERS = [(4, 2500 / 11025, "high"), (4, 100 / 11025, "low"), (3, 7000 / 11025, "low"), (1, 0.3, "high"), (2, 0.05, "low")]


def cpu_noise(seed, shape, add_rumble=True):
    """The three torch.randn_like draws the reference makes on a CPU tensor after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    z = torch.empty(shape)
    return torch.randn_like(z), torch.randn_like(z), (torch.randn_like(z) if add_rumble else None)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_matches_reference_golden(i):
    seed, C, N, sr, kw = CASES[i]
    x = torch.from_numpy(make_input(seed, C, N))
    surface, crackle, rumble = cpu_noise(seed, x.shape, kw.get("add_rumble", True))
    np.random.seed(seed)
    plan = ap.plan_vinyl_artifacts(N, sr, **kw)
    y = ap.apply_vinyl_artifacts(x.cuda(), sr, plan, surface.cuda(), crackle.cuda(), None if rumble is None else rumble.cuda())
    ref = np.load(GOLD)[f"case{i}_y"]
    assert y.shape == (C, N) and y.dtype == torch.float32
    assert np.abs(y.cpu().numpy() - ref).max() <= TOL


@pytest.mark.parametrize("n", [16, 17, 31, 33, 100, 4095, 4096, 4127, 44100, 1 << 20])
def test_filtfilt_matches_scipy(n):
    rng = np.random.default_rng(n)
    rows = 3 if n < 100000 else 1
    x = rng.standard_normal((rows, n)).astype(np.float32)
    xg = torch.from_numpy(x).cuda()
    for order, wn, btype in FILTERS:
        if n <= 3 * (order + 1):
            continue
        b, a = signal.butter(order, wn, btype=btype)
        y = ap.filtfilt(b, a, xg).cpu().numpy()
        ref = np.stack([signal.filtfilt(b, a, x[r]) for r in range(rows)])
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(y - ref).max() <= TOL * scale, (n, order, wn, btype)


def test_filtfilt_fused_inputs_and_errors():
    rng = np.random.default_rng(0)
    x, u, v = (torch.from_numpy(rng.standard_normal((2, 5000)).astype(np.float32)) for _ in range(3))
    b, a = signal.butter(3, 0.6)
    got = ap.filtfilt(b, a, x.cuda(), scale=0.25, add1=u.cuda(), add2=v.cuda()).cpu().numpy()
    s = ((x * np.float32(0.25)) + u + v).numpy()
    ref = np.stack([signal.filtfilt(b, a, s[r]) for r in range(2)])
    assert np.abs(got - ref).max() <= TOL
    with pytest.raises(ValueError):                       # scipy: "must be greater than padlen"
        ap.filtfilt(b, a, torch.zeros(1, 12, device="cuda"))
    with pytest.raises(RuntimeError):
        ap.filtfilt(b, a, x.cuda(), add1=u.cuda()[:, :100])


def test_filtfilt_many_short_rows():
    """More rows than one grid dimension holds (the launcher walks them in slabs of 65 535)."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((70000, 40)).astype(np.float32)
    b, a = signal.butter(4, 2500 / 11025, btype="high")
    y = ap.filtfilt(b, a, torch.from_numpy(x).cuda()).cpu().numpy()
    for r in (0, 1, 65534, 65535, 65536, 69999):
        assert np.abs(y[r] - signal.filtfilt(b, a, x[r])).max() <= TOL


def test_filtfilt_is_linear_and_zero_phase_at_full_size():
    """Size-independent properties on a 3-minute side (3 969 000 samples): linearity, DC gain 1 for the low-pass, and
    time-reversal symmetry filtfilt(rev(x)) == rev(filtfilt(x)) (zero phase) away from the two ends, where the
    forward-first / backward-first edge transients differ by construction."""
    n = 3_969_000
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((1, n), device="cuda", generator=g)
    w = torch.randn((1, n), device="cuda", generator=g)
    b, a = signal.butter(3, 7000 / 11025)
    fx, fw = ap.filtfilt(b, a, x), ap.filtfilt(b, a, w)
    fsum = ap.filtfilt(b, a, x, add1=w)
    assert (fsum - (fx + fw)).abs().max().item() <= 5e-6
    rev = ap.filtfilt(b, a, x.flip(-1).contiguous()).flip(-1)
    assert (rev - fx)[:, 5000:-5000].abs().max().item() <= 5e-6
    ones = torch.ones((1, n), device="cuda")
    assert (ap.filtfilt(b, a, ones) - 1).abs().max().item() <= 1e-6
    # a sampled window against scipy run on the same data around it (the filter's memory is a few hundred samples)
    lo, hi = 2_000_000, 2_050_000
    ref = signal.filtfilt(b, a, x[0, lo - 5000:hi + 5000].cpu().numpy())[5000:-5000]
    assert np.abs(fx[0, lo:hi].cpu().numpy() - ref).max() <= TOL


@pytest.mark.parametrize("seed,C,N,sr,kw", [(21, 2, 44100, 22050, {}), (22, 1, 88200, 44100, {"impulse_rate": 100.0}),
                                            (23, 4, 10000, 22050, {"impulse_rate": 2000.0, "add_rolloff": False}),
                                            (24, 1, 3 * 22050, 22050, {"impulse_rate": 0.0}),
                                            (25, 2, 2 * 22050, 22050, {"add_rumble": False, "add_rolloff": False})])
def test_matches_oracle(seed, C, N, sr, kw):
    """BASELINE chunk sizes, dense overlapping pops (2000 per second), no pops at all, filters switched off."""
    x = torch.from_numpy(make_input(seed, C, N))
    np.random.seed(seed)
    torch.manual_seed(seed)
    ref = degrade.simulate_vinyl_artifacts(x, sr, **kw)
    surface, crackle, rumble = cpu_noise(seed, x.shape, kw.get("add_rumble", True))
    np.random.seed(seed)
    plan = ap.plan_vinyl_artifacts(N, sr, **kw)
    y = ap.apply_vinyl_artifacts(x.cuda(), sr, plan, surface.cuda(), crackle.cuda(), None if rumble is None else rumble.cuda())
    assert (y.cpu() - ref).abs().max().item() <= TOL


def test_drop_in_signature_draws_on_device():
    """`simulate_vinyl_artifacts(audio, sample_rate, ...)` as the reference calls it (preprocessing.py / mixed_dataset.py):
    global generators, CUDA tensor in, same shape / dtype out; reproducible under the same seeds."""
    x = (0.1 * torch.randn(1, 44100)).cuda()
    outs = []
    for _ in range(2):
        np.random.seed(9)
        torch.manual_seed(9)
        outs.append(ap.simulate_vinyl_artifacts(x, 22050, impulse_rate=20.0))
    assert outs[0].shape == x.shape and outs[0].dtype == torch.float32 and outs[0].is_cuda
    assert torch.equal(outs[0], outs[1])
    assert 0.01 < (outs[0] - x).std().item() < 0.2


def test_add_noise_matches_reference_formula():
    x = (0.1 * torch.randn(2, 5000)).cuda()
    torch.manual_seed(3)
    y = ap.add_noise(x, 0.02)
    torch.manual_seed(3)
    ref = x + torch.randn_like(x) * 0.02           # audio_processing.py:118-119 on the same (CUDA) generator
    assert torch.equal(y, ref)


def test_generated_side_through_the_chain(state_dicts):
    """BASELINE config 4 recipe end to end on the GPU: a clean music proxy -> `simulate_vinyl_artifacts` -> the chunked
    restoration chain; the degraded input equals the oracle's degradation of the same draws and the restored output equals
    the oracle chain on that input."""
    from oracle import pipeline as opipe
    from ml_audio_restoration_b200 import RestorationPipeline
    sr, n = 22050, 3 * 4096 + 777
    t = torch.arange(n) / sr
    clean = (0.1 * torch.sin(2 * torch.pi * 220 * t) + 0.05 * torch.sin(2 * torch.pi * 554.4 * t))[None]
    surface, crackle, rumble = cpu_noise(31, clean.shape)
    np.random.seed(31)
    plan = ap.plan_vinyl_artifacts(n, sr, impulse_rate=40.0)
    worn = ap.apply_vinyl_artifacts(clean.cuda(), sr, plan, surface.cuda(), crackle.cuda(), rumble.cuda())
    np.random.seed(31)
    torch.manual_seed(31)
    worn_ref = degrade.simulate_vinyl_artifacts(clean, sr, impulse_rate=40.0)
    assert (worn.cpu() - worn_ref).abs().max().item() <= TOL
    pipe = RestorationPipeline.from_state_dicts(state_dicts["denoiser"], state_dicts["super_resolution"], state_dicts["stereo"], "cuda")
    out = pipe.restore(worn, mode="chunked", chunk_size=4096, overlap=256, return_device=True)
    ref = opipe.restore_chunked(state_dicts, worn_ref, chunk_size=4096, overlap=256)
    assert out.shape == ref.shape == (2, 2 * n)
    err = (out.cpu() - ref).abs().max().item()
    assert err <= 1e-3, err
