"""GPU: the two conv engines (tcgen05 implicit GEMM, CUDA-core cross-check) against torch conv1d
on the CPU, layer shapes taken from the three models."""
import pytest
import torch
import torch.nn.functional as F

from ml_audio_restoration_b200 import _lib
from gpu_util import debug_conv, assert_close

pytestmark = pytest.mark.gpu


def operand_round(x):
    """Operand rounding of the engines: fp16 (same 11-bit significand as TF32), round to nearest."""
    return x.half().float()


SHAPES = [
    # B, Cin, Cout, T,    k, dil   (where it occurs)
    (2, 32, 32, 300, 3, 1),      # SR trunk / denoiser enc0
    (1, 32, 64, 1000, 3, 1),     # stereo enc1a
    (2, 64, 64, 257, 1, 1),      # stereo enc1b (k1)
    (1, 64, 128, 640, 3, 2),     # stereo enc2a (dilation 2)
    (1, 128, 128, 515, 3, 4),    # stereo enc3a
    (1, 128, 128, 515, 3, 8),    # stereo enc4a (max reach)
    (1, 128, 256, 129, 1, 1),    # LSTM input projection
    (1, 64, 256, 400, 7, 1),     # fused L+R decoder layer 0
    (1, 128, 64, 400, 7, 1),     # decoder layer 1
    (1, 256, 256, 100, 3, 1),    # denoiser bottleneck
    (3, 32, 32, 128, 5, 1),      # SR hf_emphasis, exactly one tile
    (1, 16, 32, 8, 3, 1),        # tiny
]


@pytest.mark.parametrize("engine", [_lib.ENGINE_SIMT, _lib.ENGINE_UMMA], ids=["simt", "umma"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_engine_matches_torch(engine, shape):
    B, Cin, Cout, T, k, d = shape
    g = torch.Generator().manual_seed(hash(shape) % (2 ** 31))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    y = debug_conv(x, w, b, dilation=d, lrelu=1, engine=engine)
    # both engines see fp16-rounded operands (activations are rounded when stored, weights when packed);
    # the debug hook also returns the result through the fp16 activation layout
    ref = F.leaky_relu(F.conv1d(operand_round(x).double(), operand_round(w).double(), b.double(), padding=d * (k - 1) // 2, dilation=d), 0.2).float()
    assert_close(ref.half().float(), y, f"conv {shape}", max_abs=1e-2, min_snr=66.0)  # 1 fp16 ulp at |y|~8 is 7.8e-3


def test_engines_agree_bitwise_close():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 1000, generator=g)
    w = torch.randn(128, 64, 3, generator=g) / 14
    b = torch.randn(128, generator=g)
    a = debug_conv(x, w, b, dilation=2, engine=_lib.ENGINE_SIMT)
    c = debug_conv(x, w, b, dilation=2, engine=_lib.ENGINE_UMMA)
    assert_close(a, c, "simt vs umma", max_abs=1e-2, min_snr=66.0)   # both outputs are fp16-rounded


def test_k7_heads_tensor_core_vs_cuda_core(state_dicts):
    """The two stereo output heads run as a tap-along-N tcgen05 GEMM with a shifted-sum epilogue (final_umma.cu); the
    CUDA-core cross-check engine keeps the scalar kernel.  Same model, both paths, ragged lengths around the 122-output
    tile stride."""
    import torch
    from oracle.weights import make_input
    from gpu_util import make_model, assert_close
    from ml_audio_restoration_b200 import _lib
    mt = make_model("stereo", state_dicts["stereo"])
    ms = make_model("stereo", state_dicts["stereo"], engine=_lib.ENGINE_SIMT)
    for T in (121, 122, 123, 244, 245, 1000):
        x = make_input(3, T, seed=T).cuda()
        with torch.no_grad():
            assert_close(ms(x), mt(x), f"stereo heads T={T}: tensor-core vs CUDA-core engine")
