"""SiamFC cross-correlation heads (interface of projects/siamfc-pytorch/siamfc/heads.py:7-58).
The biased 1x1 adapters of ``SiamConvFC`` run through the tcgen05 conv kernel; the correlation itself is a direct
CUDA kernel (csrc/xcorr.cu) fused with ``* out_scale``."""
import torch.nn as nn

from .. import ops

__all__ = ['SiamFC', 'SiamConvFC']


class SiamFC(nn.Module):

    def __init__(self, out_scale=0.001):
        super().__init__()
        self.out_scale = out_scale

    def forward(self, z, x):
        return ops.xcorr(z, x, self.out_scale)


class SiamConvFC(nn.Module):

    def __init__(self, in_channels, channels, num_convs=1, kernel_size=1, out_scale=0.001):
        super().__init__()
        self.out_scale = out_scale
        z_convs, x_convs = [], []
        last = in_channels
        for _ in range(num_convs):
            z_convs.append(nn.Conv2d(last, channels, kernel_size))
            x_convs.append(nn.Conv2d(last, channels, kernel_size))
            last = channels
        self.z_convs = nn.Sequential(*z_convs)
        self.x_convs = nn.Sequential(*x_convs)

    def forward(self, z, x):
        z = ops.conv_stack_nhwc(z, self.z_convs)
        x = ops.conv_stack_nhwc(x, self.x_convs)
        return ops.xcorr_nhwc(z, x, self.out_scale)
