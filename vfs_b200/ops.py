"""torch-tensor front end of the C ABI (device memory + stream plumbing only; all arithmetic is in
libvfs_b200.so).  "split" tensors are bf16 [2, N, H, W, C] (hi/lo planes, NHWC) -- see include/vfs_b200.h."""
import ctypes

import torch

from . import _native as nat
from ._native import VfsConvDesc, check, current_stream, ptr


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: the B200 path has no CPU fallback')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')


def conv_out_hw(H, W, ksize, stride, dilation):
    dil = 1 if ksize == 1 else dilation
    pad = 0 if ksize == 1 else dil
    Ho = (H + 2 * pad - dil * (ksize - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (ksize - 1) - 1) // stride + 1
    return Ho, Wo


def to_split(x):
    """NCHW fp32 -> split NHWC bf16 [2,N,H,W,C]."""
    _require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.ndim == 4
    N, C, H, W = x.shape
    out = torch.empty((2, N, H, W, C), dtype=torch.bfloat16, device=x.device)
    check(nat.lib().vfs_nchw_f32_to_split(ptr(x), ptr(out), N, C, H, W, current_stream()), 'nchw_f32_to_split')
    return out


def from_split(xs):
    """split NHWC [2,N,H,W,C] -> NCHW fp32."""
    _require_cuda(xs, 'xs')
    assert xs.dtype == torch.bfloat16 and xs.ndim == 5 and xs.shape[0] == 2
    _, N, H, W, C = xs.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=xs.device)
    check(nat.lib().vfs_split_to_nchw_f32(ptr(xs), ptr(out), N, C, H, W, current_stream()), 'split_to_nchw_f32')
    return out


def pack_conv_weight(w):
    """OIHW fp32 -> split [2, Cout, k*k*Cin] with K index (r*k+s)*Cin + ci."""
    _require_cuda(w, 'w')
    assert w.dtype == torch.float32 and w.ndim == 4 and w.shape[2] == w.shape[3]
    Cout, Cin, k, _ = w.shape
    out = torch.empty((2, Cout, k * k * Cin), dtype=torch.bfloat16, device=w.device)
    check(nat.lib().vfs_pack_conv_weight(ptr(w), ptr(out), Cout, Cin, k, current_stream()), 'pack_conv_weight')
    return out


def _desc(xs, Cout, ksize, stride, dilation, relu):
    _, N, H, W, Cin = xs.shape
    return VfsConvDesc(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=ksize, stride=stride, dilation=dilation,
                       relu=int(bool(relu)))


def conv_bn_act(xs, w_split, scale, shift, ksize, stride=1, dilation=1, relu=True, residual=None,
                want_split=True, want_f32=False):
    """tcgen05 implicit-GEMM conv + folded BN + (residual) + (ReLU).  Returns (out_split|None, out_f32_nhwc|None)."""
    for t, n in ((xs, 'xs'), (w_split, 'w_split'), (scale, 'scale'), (shift, 'shift')):
        _require_cuda(t, n)
    Cout = w_split.shape[1]
    _, N, H, W, Cin = xs.shape
    assert w_split.shape[2] == ksize * ksize * Cin, (w_split.shape, ksize, Cin)
    Ho, Wo = conv_out_hw(H, W, ksize, stride, dilation)
    out = torch.empty((2, N, Ho, Wo, Cout), dtype=torch.bfloat16, device=xs.device) if want_split else None
    out32 = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=xs.device) if want_f32 else None
    if residual is not None:
        _require_cuda(residual, 'residual')
        assert tuple(residual.shape) == (2, N, Ho, Wo, Cout), (residual.shape, (2, N, Ho, Wo, Cout))
    d = _desc(xs, Cout, ksize, stride, dilation, relu)
    check(nat.lib().vfs_conv_bn_act(ctypes.byref(d), ptr(xs), ptr(w_split), ptr(scale), ptr(shift), ptr(residual),
                                    ptr(out), ptr(out32), current_stream()), 'conv_bn_act')
    return out, out32


def debug_conv_bn_act_simt(xs, w_split, scale, shift, ksize, stride=1, dilation=1, relu=True, residual=None):
    """fp32 SIMT evaluation of the same contract (test instrument).  Returns fp32 NHWC."""
    Cout = w_split.shape[1]
    _, N, H, W, Cin = xs.shape
    Ho, Wo = conv_out_hw(H, W, ksize, stride, dilation)
    out32 = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=xs.device)
    d = _desc(xs, Cout, ksize, stride, dilation, relu)
    check(nat.lib().vfs_debug_conv_bn_act_simt(ctypes.byref(d), ptr(xs), ptr(w_split), ptr(scale), ptr(shift),
                                               ptr(residual), ptr(out32), current_stream()), 'debug_conv_simt')
    return out32


def stem_forward(x, weight, scale, shift):
    """conv7x7/s2 + BN + ReLU + maxpool3x3/s2 on NCHW fp32 input -> split NHWC [2,N,Hp,Wp,64]."""
    for t, n in ((x, 'x'), (weight, 'weight'), (scale, 'scale'), (shift, 'shift')):
        _require_cuda(t, n)
    N, C, H, W = x.shape
    assert C == 3 and tuple(weight.shape) == (64, 3, 7, 7)
    Hc, Wc = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Hp, Wp = (Hc + 2 - 3) // 2 + 1, (Wc + 2 - 3) // 2 + 1
    ws_bytes = nat.lib().vfs_stem_workspace_bytes(N, H, W)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=x.device)
    out = torch.empty((2, N, Hp, Wp, 64), dtype=torch.bfloat16, device=x.device)
    check(nat.lib().vfs_stem_forward(ptr(x), ptr(weight), ptr(scale), ptr(shift), ptr(out), ptr(ws), N, H, W,
                                     current_stream()), 'stem_forward')
    return out
