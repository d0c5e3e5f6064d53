"""Drop-in `AudioDenoiser` (reference: src/models/denoiser.py:6-144) on the sm_100a kernels.

Same constructor signature, same parameter/buffer names (so reference checkpoints load with
strict=True), same `forward(x[B,1,T]) -> [B,1,T]`.  The forward body is one call into
libaudiorestore_sm100 (`ar_model_forward`): a 3-level 1-D U-Net whose k3 convs run as tcgen05
implicit GEMMs with BatchNorm folded, max-pool / skip-concat / transposed-conv interleave fused
into the epilogues, and the transient-mask branch, analytic impulse mask and final 1x1 conv
fused into one tail kernel.
"""
import torch.nn as nn

from .. import _lib
from ._native import NativeModule

_FEATURES = [32, 64, 128]


def _double_conv(cin, cout):
    # container layout of the reference `_conv_block` (denoiser.py:51-60): indices 0,1,3,4 hold parameters
    return nn.Sequential(
        nn.Conv1d(cin, cout, 3, padding=1), nn.BatchNorm1d(cout), nn.LeakyReLU(0.2, inplace=True),
        nn.Conv1d(cout, cout, 3, padding=1), nn.BatchNorm1d(cout), nn.LeakyReLU(0.2, inplace=True))


class AudioDenoiser(NativeModule):
    KIND = _lib.MODEL_DENOISER

    def __init__(self, in_channels=1, out_channels=1, features=[32, 64, 128]):
        super().__init__()
        if in_channels != 1 or out_channels != 1 or list(features) != _FEATURES:
            raise NotImplementedError(
                "the sm_100a kernels are specialised for AudioDenoiser(1, 1, [32, 64, 128]) "
                "(the configuration inference.py:51 constructs)")
        widths = [in_channels] + list(features)
        self.encoder = nn.ModuleList(_double_conv(widths[i], widths[i + 1]) for i in range(3))
        self.decoder = nn.ModuleList()
        self.pool = nn.MaxPool1d(2, 2)
        self.bottleneck = _double_conv(features[-1], 2 * features[-1])
        for f in reversed(features):
            self.decoder.append(nn.ConvTranspose1d(2 * f, f, kernel_size=2, stride=2))
            self.decoder.append(_double_conv(2 * f, f))
        f0 = features[0]
        self.transient_detector = nn.Sequential(
            nn.Conv1d(f0, f0 // 2, 3, padding=1), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv1d(f0 // 2, f0 // 4, 3, padding=1), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv1d(f0 // 4, 1, 3, padding=1), nn.Sigmoid())
        self.final_conv = nn.Conv1d(f0, out_channels, kernel_size=1)

    def _out_shape(self, B, T):
        return (B, 1, T)
