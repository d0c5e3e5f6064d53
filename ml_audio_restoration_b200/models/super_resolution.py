"""Drop-in `AudioSuperResolution` (reference: src/models/super_resolution.py:6-122).

`forward(x[B,1,T]) -> [B,1,2T]`.  Natively: CUDA-core k7 stem, nine 32->32 k3 tensor-core convs
with residual adds in the epilogue, the ConvTranspose1d(k4,s2,p1) as a 2-phase sub-pixel GEMM
whose epilogue interleaves the phases, the k5 conv at the doubled rate, and a tail kernel that
fuses the k7 reconstruction conv with the linear-interpolation residual.
"""
import torch.nn as nn

from .. import _lib
from ._native import NativeModule


class ResidualBlockEfficient(nn.Module):
    """Parameter container of the reference block (super_resolution.py:104-113)."""

    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv1d(channels, channels, 3, padding=1)
        self.bn1 = nn.BatchNorm1d(channels)
        self.relu = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = nn.Conv1d(channels, channels, 3, padding=1)
        self.bn2 = nn.BatchNorm1d(channels)


class AudioSuperResolution(NativeModule):
    KIND = _lib.MODEL_SUPER_RES

    def __init__(self, upscale_factor: int = 2, channels: int = 1, base_channels: int = 32,
                 num_residual_blocks: int = 4):
        super().__init__()
        if (upscale_factor, channels, base_channels, num_residual_blocks) != (2, 1, 32, 4):
            raise NotImplementedError(
                "the sm_100a kernels are specialised for AudioSuperResolution(upscale_factor=2, channels=1, "
                "base_channels=32, num_residual_blocks=4) (inference.py:66)")
        c = base_channels
        self.upscale_factor = upscale_factor
        self.initial = nn.Sequential(nn.Conv1d(channels, c, 7, padding=3), nn.LeakyReLU(0.2, inplace=True))
        self.residual_blocks = nn.ModuleList(ResidualBlockEfficient(c) for _ in range(num_residual_blocks))
        self.middle = nn.Sequential(nn.Conv1d(c, c, 3, padding=1), nn.BatchNorm1d(c))
        self.upsample_blocks = nn.ModuleList([nn.Sequential(
            nn.ConvTranspose1d(c, c, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, inplace=True))])
        self.hf_emphasis = nn.Sequential(nn.Conv1d(c, c, 5, padding=2), nn.LeakyReLU(0.2, inplace=True))
        self.reconstruction = nn.Conv1d(c, channels, 7, padding=3)

    def _out_shape(self, B, T):
        return (B, 1, 2 * T)
