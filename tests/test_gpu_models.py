"""GPU: the three drop-in modules against the CPU oracle and the reference's golden vectors."""
import pytest
import torch

import oracle
from oracle.weights import make_input
from ml_audio_restoration_b200 import _lib
from gpu_util import make_model, assert_close

pytestmark = pytest.mark.gpu

FWD = {"denoiser": oracle.denoiser_forward, "super_resolution": oracle.super_resolution_forward,
       "stereo": oracle.stereo_forward}
GOLDEN = {"denoiser": (2, [8, 9, 101, 1037, 2048]), "super_resolution": (3, [4, 65, 1000]), "stereo": (2, [7, 130, 1500])}
ENGINES = {"simt": _lib.ENGINE_SIMT, "umma": _lib.ENGINE_UMMA}


@pytest.fixture(scope="module")
def models(state_dicts):
    cache = {}

    def get(name, engine):
        if (name, engine) not in cache:
            cache[(name, engine)] = make_model(name, state_dicts[name], ENGINES[engine])
        return cache[(name, engine)]
    return get


@pytest.mark.parametrize("engine", ["simt", "umma"])
@pytest.mark.parametrize("name", ["denoiser", "super_resolution", "stereo"])
def test_model_matches_reference_golden(models, golden, name, engine):
    B, lengths = GOLDEN[name]
    m = models(name, engine)
    for T in lengths:
        x = make_input(B, T)
        with torch.no_grad():
            y = m(x.cuda())
        assert_close(torch.from_numpy(golden[f"{name}_T{T}"]), y, f"{name}[{engine}] T={T} vs reference golden")


@pytest.mark.parametrize("engine", ["simt", "umma"])
@pytest.mark.parametrize("name,B,T", [("denoiser", 2, 44100), ("super_resolution", 16, 44100), ("stereo", 4, 44100)])
def test_model_baseline_config_vs_oracle(models, state_dicts, name, B, T, engine):
    """BASELINE.json configs 1-3 at full size."""
    if engine == "simt" and name != "denoiser":
        B = 2  # the CUDA-core engine is a cross-check, keep it short
    x = make_input(B, T)
    ref = FWD[name](state_dicts[name], x)
    with torch.no_grad():
        y = models(name, engine)(x.cuda())
    assert_close(ref, y, f"{name}[{engine}] B={B} T={T} vs oracle")


@pytest.mark.parametrize("T", [44096, 44101, 127, 128, 129, 255, 257])
def test_ragged_lengths(models, state_dicts, T):
    for name in ("denoiser", "super_resolution", "stereo"):
        x = make_input(1, T, seed=T)
        ref = FWD[name](state_dicts[name], x)
        with torch.no_grad():
            y = models(name, "umma")(x.cuda())
        assert_close(ref, y, f"{name} T={T}")


def test_impulse_and_silence_inputs(models, state_dicts):
    x = make_input(2, 4000)
    x[0, 0, 1000] = 0.9          # a pop (impulse-mask branch, denoiser.py:62-86)
    x[0, 0, 1001] = -0.8
    x[1] = 0.0                   # digital silence
    for name in ("denoiser", "super_resolution", "stereo"):
        ref = FWD[name](state_dicts[name], x)
        with torch.no_grad():
            y = models(name, "umma")(x.cuda())
        assert_close(ref, y, f"{name} impulse/silence", min_snr=None if name == "denoiser" else 60.0)


def test_noncontiguous_input_is_accepted(models, state_dicts):
    x = make_input(2, 2 * 1000)[:, :, ::2]   # stereo_separator.py:93 makes inputs contiguous
    ref = oracle.stereo_forward(state_dicts["stereo"], x.contiguous())
    with torch.no_grad():
        y = models("stereo", "umma")(x.cuda())
    assert_close(ref, y, "stereo non-contiguous")


def test_stereo_large_batch_tensor_core_lstm(models, state_dicts):
    """More than two sequences per SM switches the recurrence to the tensor-core (mma.sync fp16) kernel;
    ragged length (not a multiple of the 8-step block) and a batch that is not a multiple of 8."""
    m = models("stereo", "umma")
    x = make_input(301, 203, seed=11)
    ref, (hn, cn) = oracle.stereo_forward(state_dicts["stereo"], x, return_state=True)
    with torch.no_grad():
        y, st = m.forward_with_state(x.cuda())
    assert_close(ref, y, "stereo B=301 T=203 (tensor-core LSTM)")
    assert_close(hn[0], st[:, 0], "carried h (tensor-core LSTM)", max_abs=1e-3, min_snr=50.0)
    assert_close(cn[0], st[:, 1], "carried c (tensor-core LSTM)", max_abs=1e-3, min_snr=50.0)


@pytest.mark.parametrize("B,T", [(1, 100), (3, 129), (2, 1500), (5, 4099), (160, 1000)])
def test_fused_chains_match_layer_by_layer(state_dicts, B, T):
    """The fused dilated-block launches (conv k3 -> conv k1 [-> LSTM input projection], conv_chain.cu) compute the
    same fp16-rounded intermediates as the layer-by-layer launches: outputs agree far inside the tolerance, and
    both agree with the oracle.  B=160 gives every SM pair several tile pairs (steady-state pipeline)."""
    fused = make_model("stereo", state_dicts["stereo"], fusion=True)
    plain = make_model("stereo", state_dicts["stereo"], fusion=False)
    x = make_input(B, T, seed=B * 7 + T)
    with torch.no_grad():
        yf = fused(x.cuda())
        yp = plain(x.cuda())
    assert_close(yp, yf, f"stereo fused vs layer-by-layer B={B} T={T}", max_abs=2e-5, min_snr=90.0)
    if B <= 5:
        assert_close(oracle.stereo_forward(state_dicts["stereo"], x), yf, f"stereo fused vs oracle B={B} T={T}")


@pytest.mark.parametrize("name", ["denoiser", "super_resolution"])
@pytest.mark.parametrize("B,T", [(1, 8), (2, 125), (1, 126), (3, 127), (1, 252), (2, 253), (1, 1009), (5, 4099), (150, 1100)])
def test_fused_double_convs_match_layer_by_layer(state_dicts, name, B, T):
    """The fused k3 -> k3 launches (U-Net double convs with the max-pool copy, super-resolution residual blocks with the
    skip add; conv_chain.cu, tile stride 126) compute the same fp16-rounded intermediates as the layer-by-layer launches:
    outputs agree far inside the tolerance, and both agree with the oracle.  Lengths around the 126-row tile stride and its
    multiples (at every U-Net level: T, T/2, T/4), odd tile counts (idle peer CTA), and B=150 for the steady-state pipeline."""
    fused = make_model(name, state_dicts[name], fusion=2)      # every pair that fits, not only the measured-faster ones
    plain = make_model(name, state_dicts[name], fusion=0)
    x = make_input(B, T, seed=B * 11 + T)
    with torch.no_grad():
        yf = fused(x.cuda())
        yp = plain(x.cuda())
    assert_close(yp, yf, f"{name} fused vs layer-by-layer B={B} T={T}", max_abs=2e-5, min_snr=90.0)
    if B <= 5:
        assert_close(FWD[name](state_dicts[name], x), yf, f"{name} fused vs oracle B={B} T={T}")


def test_error_behaviour(models, state_dicts):
    den = models("denoiser", "umma")
    with pytest.raises(RuntimeError):          # reference: RuntimeError from max_pool1d for T < 8
        den(make_input(1, 7).cuda())
    with pytest.raises(RuntimeError):          # no CPU fallback
        den(make_input(1, 64))
    with pytest.raises(RuntimeError):
        den(make_input(1, 64).cuda().squeeze(1))
    for name, shape in (("denoiser", (0, 1, 64)), ("super_resolution", (0, 1, 128)), ("stereo", (0, 2, 64))):
        y = models(name, "umma")(torch.zeros(0, 1, 64, device="cuda"))      # reference: an empty batch gives an empty output
        assert tuple(y.shape) == shape and y.is_cuda and y.dtype == torch.float32
    den.train()
    with pytest.raises(NotImplementedError):
        den(make_input(1, 64).cuda())
    den.eval()
    from ml_audio_restoration_b200.models import AudioDenoiser
    bad = dict(state_dicts["denoiser"])
    bad.pop("final_conv.bias")
    with pytest.raises(RuntimeError):          # load_state_dict(strict) semantics
        AudioDenoiser().load_state_dict(bad)


def test_weight_update_repacks(state_dicts):
    m = make_model("super_resolution", state_dicts["super_resolution"])
    x = make_input(1, 500).cuda()
    with torch.no_grad():
        y0 = m(x)
        m.reconstruction.bias.add_(0.25)
        y1 = m(x)
    assert_close(y0.cpu() + 0.25, y1, "bias update visible after repack", max_abs=1e-6, min_snr=None)


@pytest.mark.parametrize("name,B,T", [("denoiser", 2, 44100), ("super_resolution", 3, 5000), ("stereo", 4, 4100)])
def test_cuda_graph_capture_replays_the_forward(models, name, B, T):
    """`module.capture(x)`: the forward's launch train captured into a CUDA graph (BASELINE configs 1-3 are launch-bound at
    their small batches); replays on new inputs give exactly what the eager forward gives."""
    m = models(name, "umma")
    xs = [make_input(B, T, seed=50 + i).cuda() for i in range(3)]
    with torch.no_grad():
        g = m.capture(xs[0])
        for x in xs:
            want = m(x)
            got = g(x).clone()
            assert torch.equal(want, got)
    with pytest.raises(RuntimeError):
        g(make_input(B, T + 8).cuda())
