"""SiamFC cross-correlation heads with the reference's constructor / call signatures
(projects/siamfc-pytorch/siamfc/heads.py:7-58).  ``z`` is the exemplar feature map [nz,C,hz,wz], ``x`` the search
feature maps [nx,C,h,w]; search item i is correlated with exemplar i % nz and the response is multiplied by
``out_scale``.  The correlation is a direct CUDA kernel over NHWC operands (csrc/xcorr.cu); the biased 1x1 adapter
convs of ``SiamConvFC`` go through the tcgen05 conv kernel."""
from torch import nn

from .. import ops

__all__ = ['SiamFC', 'SiamConvFC']


class _XCorrHead(nn.Module):
    """Shared behaviour: optional per-branch adapter stacks, then the scaled cross-correlation."""

    def __init__(self, out_scale):
        super().__init__()
        self.out_scale = out_scale

    def _adapt(self, z, x):
        return ops.nchw_to_nhwc(z), ops.nchw_to_nhwc(x)

    def forward(self, z, x):
        z_nhwc, x_nhwc = self._adapt(z, x)
        return ops.xcorr_nhwc(z_nhwc, x_nhwc, self.out_scale)


class SiamFC(_XCorrHead):

    def __init__(self, out_scale=0.001):
        super().__init__(out_scale)


class SiamConvFC(_XCorrHead):

    def __init__(self, in_channels, channels, num_convs=1, kernel_size=1, out_scale=0.001):
        super().__init__(out_scale)
        widths = [in_channels] + [channels] * num_convs
        self.z_convs = nn.Sequential(*(nn.Conv2d(a, b, kernel_size) for a, b in zip(widths, widths[1:])))
        self.x_convs = nn.Sequential(*(nn.Conv2d(a, b, kernel_size) for a, b in zip(widths, widths[1:])))

    def _adapt(self, z, x):
        return ops.conv_stack_nhwc(z, self.z_convs), ops.conv_stack_nhwc(x, self.x_convs)
