"""ORACLE -- test infrastructure only (see oracle/__init__.py).

    SiamFC._fast_xcorr / SiamConvFC.forward   projects/siamfc-pytorch/siamfc/heads.py:16-23, 46-58
"""
import torch.nn.functional as F


def xcorr(z, x, out_scale=0.001):
    """Per-pair valid cross-correlation via grouped conv (heads.py:16-23): z [nz,c,hz,wz], x [nx,c,h,w]."""
    nz = z.size(0)
    nx, c, h, w = x.size()
    out = F.conv2d(x.reshape(-1, nz * c, h, w), z, groups=nz)
    return out.reshape(nx, -1, out.size(-2), out.size(-1)) * out_scale


def siam_conv_fc(sd, z, x, out_scale=0.001, num_convs=1):
    """1x1 (biased) conv adapters on both branches then xcorr (heads.py:46-49)."""
    for i in range(num_convs):
        z = F.conv2d(z, sd[f'z_convs.{i}.weight'], sd[f'z_convs.{i}.bias'])
        x = F.conv2d(x, sd[f'x_convs.{i}.weight'], sd[f'x_convs.{i}.bias'])
    return xcorr(z, x, out_scale)
