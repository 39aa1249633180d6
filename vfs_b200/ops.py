"""torch-tensor front end of the C ABI (device memory + stream plumbing only; all arithmetic is in
libvfs_b200.so).  "split" tensors are fp16 [2, N, H, W, C] (hi/lo planes, NHWC) -- see include/vfs_b200.h."""
import ctypes

import torch

from . import _native as nat
from ._native import VfsConvDesc, current_stream, ptr

LAUNCHES = [0]  # kernels of libvfs_b200.so launched through this module (bench.py reports the count)
# parameter data_ptr -> gradient tensor the native backward kernels ACCUMULATE into (dp.FlatTrainState registers the
# views of its flat gradient buffer here; autograd then receives None for those parameters)
GRAD_SINKS = {}
GRAD_SINK_OWNER = {}   # parameter data_ptr -> the dp.FlatTrainState owning its gradient view
# bumped whenever native kernels rewrote parameters / BN buffers without touching tensor versions (flat SGD step,
# CUDA-graph replays of a training step): part of every engine's plan / graph validity stamp
WEIGHT_EPOCH = [0]


def grad_sink(p):
    return GRAD_SINKS.get(p.data_ptr()) if (p is not None and GRAD_SINKS) else None
_KERNELS_PER_CALL = {'comm_allreduce_f32': 3, 'stem_forward': 2, 'masked_attention': 2, 'features_to_split_norm': 2, 'conv_dgrad': 1, 'conv_wgrad': 2, 'seg_postprocess': 3, 'siamfc_response_peak': 3}


def check(rc, what=''):
    """Raise on a non-zero return code, otherwise account the kernels the entry point launched."""
    nat.check(rc, what)
    LAUNCHES[0] += _KERNELS_PER_CALL.get(what, 1)


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
    """NCHW fp32 -> split NHWC fp16 [2,N,H,W,C]."""
    _require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.ndim == 4
    N, C, H, W = x.shape
    out = torch.empty((2, N, H, W, C), dtype=torch.float16, device=x.device)
    check(nat.lib().vfs_nchw_f32_to_split(ptr(x), ptr(out), N, C, H, W, current_stream()), 'nchw_f32_to_split')
    return out


def to_split_scaled(x, scale):
    """NCHW fp32 -> split NHWC of ``scale * x`` (gradient entry of the backward pass)."""
    _require_cuda(x, 'x')
    N, C, H, W = x.shape
    out = torch.empty((2, N, H, W, C), dtype=torch.float16, device=x.device)
    check(nat.lib().vfs_nchw_f32_to_split_scaled(ptr(x), ptr(out), N, C, H, W, float(scale), current_stream()),
          'nchw_f32_to_split_scaled')
    return out


def from_split(xs):
    """split NHWC [2,N,H,W,C] -> NCHW fp32."""
    _require_cuda(xs, 'xs')
    assert xs.dtype == torch.float16 and xs.ndim == 5 and xs.shape[0] == 2
    _, N, H, W, C = xs.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=xs.device)
    check(nat.lib().vfs_split_to_nchw_f32(ptr(xs), ptr(out), N, C, H, W, current_stream()), 'split_to_nchw_f32')
    return out


# Power-of-two factor the engine multiplies conv weights by before the split (1 / WEIGHT_SCALE goes into the epilogue's
# scale vector): keeps both fp16 planes of kaiming-sized weights (~0.02) in the normal range -> 22 instead of ~17 bits.
WEIGHT_SCALE = 256.0
WEIGHT_SCALE_LOG2 = 8


def pack_conv_weight(w, wscale=1.0):
    """OIHW fp32 -> split [2, Cout, k*k*Cin] of ``wscale * w`` with K index (r*k+s)*Cin + ci."""
    _require_cuda(w, 'w')
    assert w.dtype == torch.float32 and w.ndim == 4 and w.shape[2] == w.shape[3]
    Cout, Cin, k, _ = w.shape
    out = torch.empty((2, Cout, k * k * Cin), dtype=torch.float16, device=w.device)
    check(nat.lib().vfs_pack_conv_weight_scaled(ptr(w), ptr(out), Cout, Cin, k, float(wscale), current_stream()),
          'pack_conv_weight')
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
    out = torch.empty((2, N, Ho, Wo, Cout), dtype=torch.float16, device=xs.device) if want_split else None
    out32 = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=xs.device) if want_f32 else None
    if residual is not None:
        _require_cuda(residual, 'residual')
        assert tuple(residual.shape) == (2, N, Ho, Wo, Cout), (residual.shape, (2, N, Ho, Wo, Cout))
    d = _desc(xs, Cout, ksize, stride, dilation, relu)
    check(nat.lib().vfs_conv_bn_act(ctypes.byref(d), ptr(xs), ptr(w_split), ptr(scale), ptr(shift), ptr(residual),
                                    ptr(out), ptr(out32), current_stream()), 'conv_bn_act')
    return out, out32


def conv_set_pair_policy(mode=2, min_pair_tiles=48):
    """0 = 1-CTA tiles only, 1 = CTA pairs whenever possible, 2 = pairs for launches with >= min_pair_tiles."""
    check(nat.lib().vfs_conv_set_pair_policy(int(mode), int(min_pair_tiles)), 'conv_set_pair_policy')


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


def stem_pack_weight(w):
    """OIHW fp32 [64,3,7,7] -> split [2,64,192] operand of the tensor-core stem."""
    _require_cuda(w, 'w')
    assert tuple(w.shape) == (64, 3, 7, 7) and w.dtype == torch.float32
    out = torch.empty((2, 64, 192), dtype=torch.float16, device=w.device)
    check(nat.lib().vfs_stem_pack_weight(ptr(w.contiguous()), ptr(out), current_stream()), 'stem_pack_weight')
    return out


def stem_forward(x, weight, scale, shift):
    """conv7x7/s2 + BN + ReLU + maxpool3x3/s2 on NCHW fp32 input -> split NHWC [2,N,Hp,Wp,64].
    ``weight``: packed by stem_pack_weight (an OIHW fp32 tensor is packed on the fly)."""
    for t, n in ((x, 'x'), (weight, 'weight'), (scale, 'scale'), (shift, 'shift')):
        _require_cuda(t, n)
    N, C, H, W = x.shape
    assert C == 3
    if weight.dtype == torch.float32:
        weight = stem_pack_weight(weight)
    assert tuple(weight.shape) == (2, 64, 192)
    Hc, Wc = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Hp, Wp = (Hc + 2 - 3) // 2 + 1, (Wc + 2 - 3) // 2 + 1
    ws_bytes = nat.lib().vfs_stem_workspace_bytes(N, H, W)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=x.device)
    out = torch.empty((2, N, Hp, Wp, 64), dtype=torch.float16, device=x.device)
    check(nat.lib().vfs_stem_forward(ptr(x), ptr(weight), ptr(scale), ptr(shift), ptr(out), ptr(ws), N, H, W,
                                     current_stream()), 'stem_forward')
    return out


# ---------------------------------------------------------------------------------------------------------
# train-mode BatchNorm around the conv kernel (batch statistics; SyncBN exchange = one all-reduce of the stats)
# ---------------------------------------------------------------------------------------------------------
_CONST_CACHE = {}


def _const_vec(value, n, device):
    key = (float(value), int(n), str(device))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.full((n, ), float(value), dtype=torch.float32, device=device)
        _CONST_CACHE[key] = t
    return t


def conv_stats(xs, w_split, ksize, stride=1, dilation=1, stats=None):
    """Raw conv output (fp32 NHWC) + per-channel [sum | sum of squares] (fp64 [2*Cout]) in one kernel.  ``stats``: a
    ZEROED fp64 [2*Cout] vector to accumulate into (the engine hands out slices of one arena per pass)."""
    Cout = w_split.shape[1]
    _, N, H, W, Cin = xs.shape
    Ho, Wo = conv_out_hw(H, W, ksize, stride, dilation)
    z = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=xs.device)
    if stats is None:
        stats = torch.zeros((2 * Cout, ), dtype=torch.float64, device=xs.device)
    d = _desc(xs, Cout, ksize, stride, dilation, False)
    check(nat.lib().vfs_conv_stats(ctypes.byref(d), ptr(xs), ptr(w_split), ptr(_const_vec(1, Cout, xs.device)),
                                   ptr(_const_vec(0, Cout, xs.device)), ptr(z), ptr(stats), current_stream()),
          'conv_stats')
    return z, stats


def channel_stats(z, stats=None):
    """fp32 [..., C] -> fp64 [2*C] per-channel sum | sum of squares (``stats``: zeroed vector to accumulate into)."""
    C = z.shape[-1]
    if stats is None:
        stats = torch.zeros((2 * C, ), dtype=torch.float64, device=z.device)
    check(nat.lib().vfs_channel_stats_f32(ptr(z), ptr(stats), z.numel() // C, C, current_stream()), 'channel_stats')
    return stats


def cross_rank_sum_(t):
    """In-place sum of a small CUDA statistics tensor over the ranks of the default group.  With a peer communicator
    installed (vfs_b200.peer) this is one kernel over NVLink peer memory on the current stream (no NCCL call, no host
    round trip, graph-capturable); otherwise torch.distributed.all_reduce.  Returns the world size."""
    import torch.distributed as dist
    from . import peer
    comm = peer.active()
    if comm is not None and comm.world > 1:
        comm.allreduce_small_(t)
        return comm.world
    dist.all_reduce(t)
    return dist.get_world_size()


def bn_finalize(stats, count, bn, nbt_list=None):
    """Batch statistics -> (scale, shift, save_mean, save_invstd); updates bn.running_* like torch.  With
    SyncBatchNorm in an initialised multi-rank process group the statistics are all-reduced first (that IS the
    SyncBN exchange: [sum, sum of squares] is equivalent to torch's gather of mean/invstd/count)."""
    C = stats.numel() // 2
    if _is_cross_rank_syncbn(bn):
        count = count * cross_rank_sum_(stats)
    dev = stats.device
    scale = torch.empty((C, ), dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale)
    invstd = torch.empty_like(scale)
    track = bn.track_running_stats and bn.running_mean is not None
    momentum = 0.1 if bn.momentum is None else bn.momentum
    check(nat.lib().vfs_bn_finalize(ptr(stats), float(count), ptr(bn.weight.detach()) if bn.affine else None,
                                    ptr(bn.bias.detach()) if bn.affine else None,
                                    ptr(bn.running_mean) if track else None, ptr(bn.running_var) if track else None,
                                    float(momentum), float(bn.eps), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), C,
                                    current_stream()), 'bn_finalize')
    if track and bn.num_batches_tracked is not None:
        if nbt_list is not None:
            nbt_list.append(bn.num_batches_tracked)   # bumped by ONE multi-tensor launch at the end of the pass
        else:
            bn.num_batches_tracked += 1   # also bumps a tensor version so cached eval-mode folds are refreshed
    return scale, shift, mean, invstd


def _z_args(z):
    """(fp32 pointer, split pointer, (N, H, W, C)) of a raw conv output kept as fp32 NHWC or as a split tensor."""
    if z.dtype == torch.float16:
        assert z.ndim == 5 and z.shape[0] == 2
        return None, ptr(z), tuple(z.shape[1:])
    return ptr(z), None, tuple(z.shape)


def bn_apply(z, scale, shift, residual=None, relu=True):
    """Raw conv output z (fp32 NHWC, or split NHWC from conv_stats_split) -> split NHWC relu?(z*scale + shift
    (+residual))."""
    zf, zs, (N, H, W, C) = _z_args(z)
    out = torch.empty((2, N, H, W, C), dtype=torch.float16, device=z.device)
    check(nat.lib().vfs_bn_apply(zf, zs, ptr(scale), ptr(shift), ptr(residual), ptr(out), N * H * W, C, int(relu),
                                 current_stream()), 'bn_apply')
    return out


def conv_stats_split(xs, w_split, ksize, stride=1, dilation=1, stats=None, wscale=1.0):
    """Train-mode forward conv: raw output as a SPLIT tensor [2,N,Ho,Wo,Cout] (TMA epilogue) + per-channel
    [sum | sum of squares] accumulated into ``stats`` (zeroed fp64 [2*Cout]) by the epilogue's math warps."""
    Cout = w_split.shape[1]
    _, N, H, W, Cin = xs.shape
    Ho, Wo = conv_out_hw(H, W, ksize, stride, dilation)
    z = torch.empty((2, N, Ho, Wo, Cout), dtype=torch.float16, device=xs.device)
    if stats is None:
        stats = torch.zeros((2 * Cout, ), dtype=torch.float64, device=xs.device)
    d = _desc(xs, Cout, ksize, stride, dilation, False)
    check(nat.lib().vfs_conv_stats_split(ctypes.byref(d), ptr(xs), ptr(w_split),
                                         ptr(_const_vec(1.0 / wscale, Cout, xs.device)),
                                         ptr(_const_vec(0, Cout, xs.device)), ptr(z), ptr(stats), current_stream()),
          'conv_stats')
    return z, stats


def stem_conv_raw(x, weight):
    N, C, H, W = x.shape
    if weight.dtype == torch.float32:
        weight = stem_pack_weight(weight)
    Hc, Wc = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    z = torch.empty((N, Hc, Wc, 64), dtype=torch.float32, device=x.device)
    check(nat.lib().vfs_stem_conv_raw(ptr(x), ptr(weight), ptr(z), N, H, W, current_stream()), 'stem_conv_raw')
    return z


def stem_bn_relu_pool(z, scale, shift, in_hw):
    N, Hc, Wc, _ = z.shape
    H, W = in_hw
    Hp, Wp = (Hc + 2 - 3) // 2 + 1, (Wc + 2 - 3) // 2 + 1
    out = torch.empty((2, N, Hp, Wp, 64), dtype=torch.float16, device=z.device)
    check(nat.lib().vfs_stem_bn_relu_pool(ptr(z), ptr(scale), ptr(shift), ptr(out), N, H, W, current_stream()),
          'stem_bn_relu_pool')
    return out


def pack_conv_weight_dgrad(w, wscale=1.0):
    """OIHW fp32 -> split [2, Cin, k*k*Cout] of ``wscale * w`` (flipped kernel, roles of Cin/Cout swapped) for
    conv_dgrad."""
    _require_cuda(w, 'w')
    Cout, Cin, k, _ = w.shape
    out = torch.empty((2, Cin, k * k * Cout), dtype=torch.float16, device=w.device)
    check(nat.lib().vfs_pack_conv_weight_dgrad_scaled(ptr(w), ptr(out), Cout, Cin, k, float(wscale),
                                                      current_stream()), 'pack_conv_weight_dgrad')
    return out


def conv_dgrad(dz, wt_split, in_hw, ksize, stride=1, dilation=1, add=None, wscale=1.0):
    """dX = conv_transpose(dZ, W) (+ add): ``dz`` split [2,N,Ho,Wo,Cout], ``wt_split`` from pack_conv_weight_dgrad,
    ``in_hw`` the forward input (H, W); returns split [2,N,H,W,Cin]."""
    _, N, Ho, Wo, Cout = dz.shape
    Cin = wt_split.shape[1]
    H, W = in_hw
    assert conv_out_hw(H, W, ksize, stride, dilation) == (Ho, Wo), (in_hw, dz.shape)
    dx = torch.empty((2, N, H, W, Cin), dtype=torch.float16, device=dz.device)
    if add is not None:
        assert tuple(add.shape) == tuple(dx.shape)
    d = VfsConvDesc(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=ksize, stride=stride, dilation=dilation, relu=0)
    check(nat.lib().vfs_conv_dgrad(ctypes.byref(d), ptr(dz), ptr(wt_split), ptr(_const_vec(1.0 / wscale, Cin, dz.device)),
                                   ptr(_const_vec(0, Cin, dz.device)), ptr(add), ptr(dx), current_stream()),
          'conv_dgrad')
    return dx


def conv_wgrad(xs, dz, ksize, stride=1, dilation=1, out=None, accumulate=False, out_scale=1.0):
    """dW (OIHW fp32) of the conv with forward input ``xs`` split [2,N,H,W,Cin] and output gradient ``dz`` split
    [2,N,Ho,Wo,Cout].  ``out``: existing gradient tensor to overwrite / accumulate into."""
    _, N, H, W, Cin = xs.shape
    Cout = dz.shape[-1]
    assert conv_out_hw(H, W, ksize, stride, dilation) == tuple(dz.shape[2:4]), (xs.shape, dz.shape)
    if out is None:
        out = torch.empty((Cout, Cin, ksize, ksize), dtype=torch.float32, device=xs.device)
        accumulate = False
    ws = torch.empty((nat.lib().vfs_conv_wgrad_workspace_bytes(Cout, Cin, ksize) // 4, ), dtype=torch.float32,
                     device=xs.device)
    d = VfsConvDesc(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=ksize, stride=stride, dilation=dilation, relu=0)
    check(nat.lib().vfs_conv_wgrad(ctypes.byref(d), ptr(xs), ptr(dz), ptr(ws), ptr(out), int(accumulate),
                                   float(out_scale), current_stream()), 'conv_wgrad')
    return out


# ---------------------------------------------------------------------------------------------------------
# SimSiam head / loss
# ---------------------------------------------------------------------------------------------------------
def _f32c(t, name):
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        raise RuntimeError(f'{name} must be float32')
    return t


def global_avg_pool(x):
    """[B,C,h,w] fp32 NCHW -> [B,C] (differentiable through the native backward kernel)."""
    if torch.is_grad_enabled() and x.requires_grad:
        from .autograd import AvgPoolFunction
        return AvgPoolFunction.apply(x)
    return global_avg_pool_raw(x)


def global_avg_pool_raw(x):
    x = _f32c(x.contiguous(), 'x')
    B, C = x.shape[:2]
    HW = x[0, 0].numel()
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    check(nat.lib().vfs_global_avg_pool(ptr(x), ptr(out), B, C, HW, current_stream()), 'global_avg_pool')
    return out


def linear_forward(x, weight, bias):
    """y = x @ W^T + b (fp32 CUDA tensors)."""
    x = _f32c(x.contiguous(), 'x')
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    check(nat.lib().vfs_linear(ptr(x), ptr(_f32c(weight.contiguous(), 'weight')), ptr(bias), ptr(y), M, N, K,
                               current_stream()), 'linear')
    return y


CROSS_RANK_SYNCBN = [True]   # test switch: False makes SyncBatchNorm use this rank's statistics only


def _is_cross_rank_syncbn(bn):
    import torch.distributed as dist
    return CROSS_RANK_SYNCBN[0] and isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() \
        and dist.get_world_size() > 1


def bn1d_forward_(y, bn, relu):
    """In-place BatchNorm1d(+ReLU) over y [M,N] with torch's running-stat bookkeeping; returns (mean, invstd,
    training) for the backward pass.  SyncBatchNorm in a multi-rank group takes the two-phase path: per-feature
    sums -> all-reduce -> finalise (running stats) -> affine+ReLU."""
    M, N = y.shape
    training = bn.training or (bn.running_mean is None)
    if training and _is_cross_rank_syncbn(bn):
        stats = channel_stats(y)
        scale, shift, mean, invstd = bn_finalize(stats, M, bn)    # all-reduces the sums, count = M * world
        check(nat.lib().vfs_affine_act_f32(ptr(y), ptr(scale), ptr(shift), M, N, int(relu), current_stream()),
              'affine_act')
        return mean, invstd, True
    momentum = 0.1 if bn.momentum is None else bn.momentum
    mean = torch.empty((N, ), dtype=torch.float32, device=y.device)
    invstd = torch.empty_like(mean)
    check(nat.lib().vfs_bn1d_act(ptr(y), M, N, ptr(bn.weight.detach()) if bn.affine else None,
                                 ptr(bn.bias.detach()) if bn.affine else None, ptr(bn.running_mean),
                                 ptr(bn.running_var), float(bn.eps), float(momentum), int(training), int(relu),
                                 ptr(mean), ptr(invstd), current_stream()), 'bn1d_act')
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return mean, invstd, training


def relu_(y):
    check(nat.lib().vfs_relu(ptr(y), y.numel(), current_stream()), 'relu')
    return y


def linear_bn_act(x, lin, bn=None, relu=False):
    """y = relu?(bn?(x @ W^T + b)) with ``lin`` an nn.Linear and ``bn`` a BatchNorm1d/SyncBatchNorm module.  When
    gradients are enabled the call goes through the native autograd Function (vfs_b200/autograd.py)."""
    if torch.is_grad_enabled() and (x.requires_grad or lin.weight.requires_grad):
        from .autograd import LinearBnActFunction
        gamma = bn.weight if (bn is not None and bn.affine) else None
        beta = bn.bias if (bn is not None and bn.affine) else None
        return LinearBnActFunction.apply(x, lin.weight, lin.bias, gamma, beta, bn, bool(relu))
    y = linear_forward(x, lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None)
    if bn is not None:
        bn1d_forward_(y, bn, relu)
    elif relu:
        relu_(y)
    return y


# ---- backward pieces of the head -------------------------------------------------------------------------
def linear_backward(dy, x, weight, need_dx=True, dW_out=None, db_out=None):
    """Returns (dx | None, dW, db) for y = x W^T + b.  ``dW_out`` / ``db_out``: existing gradient tensors to
    ACCUMULATE into (flat gradient views); dW / db are then returned as None."""
    M, N = dy.shape
    K = x.shape[1]
    dx = torch.empty((M, K), dtype=torch.float32, device=dy.device) if need_dx else None
    sunk = dW_out is not None
    dW = dW_out if sunk else torch.empty((N, K), dtype=torch.float32, device=dy.device)
    db = db_out if sunk else torch.empty((N, ), dtype=torch.float32, device=dy.device)
    check(nat.lib().vfs_linear_backward(ptr(dy), ptr(x), ptr(weight), ptr(dx), ptr(dW), ptr(db), M, N, K, int(sunk),
                                        current_stream()), 'linear_backward')
    LAUNCHES[0] += 1 if need_dx else 0
    return (dx, None, None) if sunk else (dx, dW, db)


def _bn_param_sinks(bn):
    """(dgamma, dbeta) accumulation targets registered for an affine norm layer, or (None, None)."""
    if bn is None or not getattr(bn, 'affine', False):
        return None, None
    dg, db = grad_sink(bn.weight), grad_sink(bn.bias)
    return (dg, db) if (dg is not None and db is not None) else (None, None)


def bn1d_backward(dy, pre, out, gamma, mean, invstd, training, relu, bn=None):
    """Returns (dpre, dgamma, dbeta).  With a cross-rank SyncBatchNorm the two per-feature sums are all-reduced
    (two-phase kernels shared with the backbone's BN backward)."""
    M, N = dy.shape
    if training and bn is not None and _is_cross_rank_syncbn(bn):
        sums = torch.zeros((2 * N, ), dtype=torch.float64, device=dy.device)
        yf = out if relu else None
        check(nat.lib().vfs_bn_bwd_reduce(None, ptr(dy), None, ptr(yf), ptr(pre), None, ptr(mean), ptr(invstd), ptr(sums), M,
                                          N, current_stream()), 'bn_bwd_reduce')
        world = _sync_sums(sums, bn)
        count = M * world
        dpre = torch.empty_like(dy)
        dg, db = _bn_param_sinks(bn)
        sunk = dg is not None
        if not sunk:
            dg = torch.empty((N, ), dtype=torch.float32, device=dy.device)
            db = torch.empty_like(dg)
        check(nat.lib().vfs_bn_bwd_apply(None, ptr(dy), None, ptr(yf), ptr(pre), None, ptr(mean), ptr(invstd), ptr(gamma),
                                         ptr(sums), float(count), None, ptr(dpre), None, ptr(dg), ptr(db), int(sunk),
                                         1.0 / world, M, N, current_stream()), 'bn_bwd_apply')
        return (dpre, None, None) if sunk else (dpre, dg, db)
    dpre = torch.empty_like(dy)
    dg, db = _bn_param_sinks(bn)
    sunk = dg is not None
    if not sunk:
        dg = torch.empty((N, ), dtype=torch.float32, device=dy.device)
        db = torch.empty_like(dg)
    check(nat.lib().vfs_bn1d_backward(ptr(dy), ptr(pre), ptr(out), ptr(dpre), M, N, ptr(gamma), ptr(mean),
                                      ptr(invstd), int(training), int(relu), ptr(dg), ptr(db), int(sunk),
                                      current_stream()), 'bn1d_backward')
    return (dpre, None, None) if sunk else (dpre, dg, db)


def relu_backward(dy, out):
    dx = torch.empty_like(dy)
    check(nat.lib().vfs_relu_backward(ptr(dy), ptr(out), ptr(dx), dy.numel(), current_stream()), 'relu_backward')
    return dx


def avgpool_backward(dy, hw_shape):
    """dY [B,C] -> dX NCHW [B,C,h,w] = dY / (h*w)."""
    B, C = dy.shape
    h, w = hw_shape
    dx = torch.empty((B, C, h, w), dtype=torch.float32, device=dy.device)
    check(nat.lib().vfs_avgpool_backward(ptr(dy), ptr(dx), B, C, h * w, current_stream()), 'avgpool_backward')
    return dx


def cosine_loss_backward(p, z, gout, with_norm=True, negative=False):
    B, D = p.shape
    dp = torch.empty_like(p)
    check(nat.lib().vfs_cosine_loss_backward(ptr(p), ptr(z), ptr(gout), ptr(dp), B, D, int(with_norm), int(negative),
                                             current_stream()), 'cosine_loss_backward')
    return dp


# ---- backward pieces of the backbone ---------------------------------------------------------------------
def _sync_sums(sums, bn):
    if _is_cross_rank_syncbn(bn):
        return cross_rank_sum_(sums)
    return 1


def bn_backward(dy, y_for_relu, z, mean, invstd, bn, want_g=False, dy_is_f32=False, want_f32=False,
                param_scale=1.0, sums=None, eval_mode=False):
    """BatchNorm(+ReLU) backward.  ``dy``: split [2,N,H,W,C] (or fp32 NHWC when dy_is_f32), ``y_for_relu``: forward
    output (split) or None, ``z`` raw conv output fp32 NHWC.  Returns (dz, g|None, dgamma, dbeta); dz is split (or
    fp32 when want_f32).  SyncBN: the two per-channel sums are all-reduced across ranks."""
    zf, zs, (N, H, W, C) = _z_args(z)
    M = N * H * W
    if sums is None:
        sums = torch.zeros((2 * C, ), dtype=torch.float64, device=z.device)
    dys, dyf = (None, dy) if dy_is_f32 else (dy, None)
    need_param = bn.affine and bn.weight.requires_grad
    if eval_mode:
        # BN on running statistics (norm_eval / frozen-BN fine-tuning): no batch-statistic terms, nothing to exchange
        # between ranks; the two sums are only needed for dgamma / dbeta
        if need_param:
            check(nat.lib().vfs_bn_bwd_reduce(ptr(dys), ptr(dyf), ptr(y_for_relu), None, zf, zs, ptr(mean), ptr(invstd),
                                              ptr(sums), M, C, current_stream()), 'bn_bwd_reduce')
        world, count = 1, 0.0
    else:
        check(nat.lib().vfs_bn_bwd_reduce(ptr(dys), ptr(dyf), ptr(y_for_relu), None, zf, zs, ptr(mean), ptr(invstd),
                                          ptr(sums), M, C, current_stream()), 'bn_bwd_reduce')
        world = _sync_sums(sums, bn)
        count = M * world
    # dgamma/dbeta come out of the all-reduced sums, i.e. already summed over ranks; the data-parallel gradient
    # all-reduce averages parameter gradients afterwards, so hand it this rank's 1/world share.
    param_scale = param_scale / world
    dz = torch.empty((N, H, W, C), dtype=torch.float32, device=z.device) if want_f32 else \
        torch.empty((2, N, H, W, C), dtype=torch.float16, device=z.device)
    g = torch.empty((2, N, H, W, C), dtype=torch.float16, device=z.device) if want_g else None
    dg, db, acc = (grad_sink(bn.weight), grad_sink(bn.bias), 1) if bn.affine else (None, None, 0)
    sunk = dg is not None and db is not None
    if eval_mode and not need_param:
        dg = db = None
        sunk, acc = True, 0
    elif not sunk:
        dg = torch.empty((C, ), dtype=torch.float32, device=z.device)
        db = torch.empty_like(dg)
        acc = 0
    check(nat.lib().vfs_bn_bwd_apply(ptr(dys), ptr(dyf), ptr(y_for_relu), None, zf, zs, ptr(mean), ptr(invstd),
                                     ptr(bn.weight.detach()) if bn.affine else None,
                                     ptr(sums) if (not eval_mode or need_param) else None, float(count),
                                     None if want_f32 else ptr(dz), ptr(dz) if want_f32 else None, ptr(g), ptr(dg),
                                     ptr(db), acc, float(param_scale), M, C, current_stream()), 'bn_bwd_apply')
    if sunk:
        return dz, g, None, None     # accumulated into the registered gradient views
    return dz, g, dg, db


def stem_pool_relu_backward(dpool, z, scale, shift, in_hw):
    """dPool split [2,N,Hp,Wp,64] -> gradient w.r.t. the BN output on the conv grid, fp32 [N,Hc,Wc,64]."""
    N, Hc, Wc, _ = z.shape
    g = torch.empty_like(z)
    check(nat.lib().vfs_stem_pool_relu_bwd(ptr(dpool), ptr(z), ptr(scale), ptr(shift), ptr(g), N, in_hw[0], in_hw[1],
                                           current_stream()), 'stem_pool_relu_bwd')
    return g


def stem_wgrad(x, dz, out_scale=1.0, out=None):
    """7x7 stem weight gradient; ``out``: existing gradient tensor to ACCUMULATE into."""
    N, _, H, W = x.shape
    dw = torch.empty((64, 3, 7, 7), dtype=torch.float32, device=x.device) if out is None else out
    check(nat.lib().vfs_stem_wgrad(ptr(x), ptr(dz), ptr(dw), int(out is not None), float(out_scale), N, H, W,
                                   current_stream()), 'stem_wgrad')
    return dw


def sgd_momentum_step_(p, grad, buf, lr, momentum, weight_decay, first, grad_scale=1.0):
    check(nat.lib().vfs_sgd_momentum_step(ptr(p), ptr(grad), ptr(buf), p.numel(), float(lr), float(momentum),
                                          float(weight_decay), int(first), float(grad_scale), current_stream()),
          'sgd_momentum_step')


def overflow_count(reset=True):
    """Values that left the fp16 range while being split since the last reset (synchronises the device)."""
    return int(nat.lib().vfs_overflow_count(int(reset)))


def cosine_sim_loss(p, z, with_norm=True, negative=False):
    """Per-sample 2 - 2*cos(p, z) (or -cos) for [B, D] inputs (differentiable w.r.t. ``p``)."""
    if torch.is_grad_enabled() and (p.requires_grad or z.requires_grad):
        from .autograd import CosineLossFunction
        return CosineLossFunction.apply(p, z, bool(with_norm), bool(negative))
    return cosine_sim_loss_raw(p, z, with_norm, negative)


def cosine_sim_loss_raw(p, z, with_norm=True, negative=False):
    p = _f32c(p.contiguous(), 'cls_score')
    z = _f32c(z.contiguous(), 'label')
    assert p.shape == z.shape and p.ndim == 2, (p.shape, z.shape)
    B, D = p.shape
    loss = torch.empty((B, ), dtype=torch.float32, device=p.device)
    check(nat.lib().vfs_cosine_sim_loss(ptr(p), ptr(z), ptr(loss), B, D, int(with_norm), int(negative),
                                        current_stream()), 'cosine_sim_loss')
    return loss


# ---------------------------------------------------------------------------------------------------------
# SiamFC
# ---------------------------------------------------------------------------------------------------------
def nchw_to_nhwc(x):
    x = _f32c(x.contiguous(), 'x')
    N, C, H, W = x.shape
    out = torch.empty((N, H, W, C), dtype=torch.float32, device=x.device)
    check(nat.lib().vfs_nchw_to_nhwc_f32(ptr(x), ptr(out), N, C, H, W, current_stream()), 'nchw_to_nhwc_f32')
    return out


def xcorr_nhwc(z, x, out_scale):
    """z [nz,hz,wz,C], x [nx,h,w,C] fp32 NHWC -> [nx,1,ho,wo]."""
    _f32c(z, 'z')
    _f32c(x, 'x')
    nz, hz, wz, C = z.shape
    nx, h, w, C2 = x.shape
    assert C == C2
    out = torch.empty((nx, 1, h - hz + 1, w - wz + 1), dtype=torch.float32, device=x.device)
    check(nat.lib().vfs_xcorr_nhwc(ptr(z), ptr(x), ptr(out), nz, nx, C, hz, wz, h, w, float(out_scale),
                                   current_stream()), 'xcorr_nhwc')
    return out


def xcorr(z, x, out_scale):
    """NCHW inputs (reference SiamFC.forward signature)."""
    return xcorr_nhwc(nchw_to_nhwc(z), nchw_to_nhwc(x), out_scale)


def conv_stack_nhwc(x, convs):
    """Apply a Sequential of biased 1x1 nn.Conv2d (SiamConvFC adapters, heads.py:40-48) through the tcgen05 conv
    kernel; NCHW fp32 in, NHWC fp32 out."""
    mods = list(convs)
    if not mods:
        return nchw_to_nhwc(x)
    xs = to_split(_f32c(x.contiguous(), 'x'))
    out32 = None
    for i, m in enumerate(mods):
        assert isinstance(m, torch.nn.Conv2d) and m.kernel_size == (1, 1) and m.stride == (1, 1), \
            'vfs_b200 SiamConvFC supports 1x1 adapters (the reference default)'
        w = pack_conv_weight(m.weight.detach().float().contiguous(), WEIGHT_SCALE)
        scale = torch.full((m.out_channels, ), 1.0 / WEIGHT_SCALE, dtype=torch.float32, device=x.device)
        shift = (m.bias.detach().float() if m.bias is not None else torch.zeros_like(scale)).contiguous()
        last = i == len(mods) - 1
        xs, out32 = conv_bn_act(xs, w, scale, shift, 1, 1, 1, relu=False, want_split=not last, want_f32=last)
    return out32


# ---------------------------------------------------------------------------------------------------------
# restricted attention (DAVIS label propagation)
# ---------------------------------------------------------------------------------------------------------
def features_to_split(x, normalize=True):
    """NCHW fp32 [N,C,H,W] -> (optionally L2-normalised over C) split NHWC [2,N,H,W,C]."""
    x = _f32c(x.contiguous(), 'x')
    N, C, H, W = x.shape
    out = torch.empty((2, N, H, W, C), dtype=torch.float16, device=x.device)
    ws = torch.empty((N * H * W, ), dtype=torch.float32, device=x.device) if normalize else None
    check(nat.lib().vfs_features_to_split(ptr(x), ptr(out), ptr(ws), N, C, H, W, int(normalize), current_stream()),
          'features_to_split_norm' if normalize else 'features_to_split')
    return out


def normalize_split(xs, out=None):
    """split NHWC [2,N,H,W,C] -> L2-normalised over C (into ``out`` [2,N',H,W,C] slice-compatible buffer)."""
    _require_cuda(xs, 'xs')
    _, N, H, W, C = xs.shape
    if out is None:
        out = torch.empty_like(xs)
    check(nat.lib().vfs_normalize_split(ptr(xs), ptr(out), N * H * W, C, xs.stride(0), out.stride(0),
                                        current_stream()), 'normalize_split')
    return out


_MASK_MODES = {None: 0, 'circle': 1, 'square': 2}


def attention_set_wide(mode=1):
    """1 = two key tiles per step in the scores kernel (default), 0 = one."""
    check(nat.lib().vfs_attention_set_wide(int(mode)), 'attention_set_wide')


def attention_bank_batched(q_bank, q_ids, k_bank, key_ids, values, val_ids, v_batch_stride, v_frame_stride,
                           v_chan_stride, Cv, mask, temperature, topk, non_mask_len=0, mode='softmax',
                           return_topk=False):
    """P independent (query frame, key set) problems in one launch.  ``q_bank`` [2,Fq,H,W,C] / ``k_bank`` [2,Fk,H,W,C]
    normalised split banks; ``q_ids`` [P], ``key_ids`` [P][T] bank frames; ``val_ids`` [P][T] frame indices used to
    address ``values`` (see include/vfs_b200.h).  Returns out [P,Cv,H*W] (+ top-k values / indices [P,topk,H*W])."""
    from ._native import VfsAttnDesc
    _require_cuda(k_bank, 'k_bank')
    assert q_bank.dtype == torch.float16 and k_bank.dtype == torch.float16
    _, Fk, H, W, C = k_bank.shape
    Fq = q_bank.shape[1]
    assert q_bank.shape[2:] == (H, W, C), (q_bank.shape, k_bank.shape)
    assert q_bank[0].is_contiguous() and k_bank[0].is_contiguous()
    P = len(q_ids)
    T = len(key_ids[0])
    assert len(key_ids) == P and len(val_ids) == P and all(len(k) == T for k in key_ids)
    d = VfsAttnDesc(H=H, W=W, C=C, Cv=Cv, T=T, topk=topk,
                    mask_mode=_MASK_MODES[mask.mode if mask is not None else None],
                    radius_y=mask.radius_y if mask is not None else 0,
                    radius_x=mask.radius_x if mask is not None else 0, non_mask_len=non_mask_len,
                    mode=0 if mode == 'softmax' else 1, temperature=float(temperature))
    ws_bytes = nat.lib().vfs_attention_workspace_bytes(ctypes.byref(d), P)
    ws = torch.empty((ws_bytes, ), dtype=torch.uint8, device=k_bank.device)
    out = torch.empty((P, Cv, H * W), dtype=torch.float32, device=k_bank.device)
    tv = torch.empty((P, topk, H * W), dtype=torch.float32, device=k_bank.device) if return_topk else None
    ti = torch.empty((P, topk, H * W), dtype=torch.int32, device=k_bank.device) if return_topk else None
    qarr = (ctypes.c_int32 * P)(*[int(i) for i in q_ids])
    karr = (ctypes.c_int32 * (P * T))(*[int(i) for row in key_ids for i in row])
    varr = (ctypes.c_int32 * (P * T))(*[int(i) for row in val_ids for i in row])
    check(nat.lib().vfs_masked_attention_batched(
        ctypes.byref(d), P, ptr(q_bank), q_bank.stride(0), Fq, qarr, ptr(k_bank), k_bank.stride(0), Fk, karr,
        ptr(values), varr, int(v_batch_stride), int(v_frame_stride), int(v_chan_stride), ptr(out), ptr(tv), ptr(ti),
        ptr(ws), ws_bytes, current_stream()), 'masked_attention')
    if return_topk:
        return out, tv, ti
    return out


def attention_bank(q_split, k_bank, key_frame_ids, values, v_frame_stride, v_chan_stride, Cv, mask, temperature,
                   topk, non_mask_len=0, mode='softmax', return_topk=False):
    """One query frame: ``q_split`` [2,1,H,W,C] view (may be a slice of the bank), ``k_bank`` [2,F,H,W,C] normalised
    split bank, ``key_frame_ids`` python list of bank frames, ``values`` fp32 tensor addressed by the given strides.
    Returns out [Cv, H*W] (and top-k values / indices [topk, H*W])."""
    ids = [list(key_frame_ids)]
    r = attention_bank_batched(q_split, [0], k_bank, ids, values, ids, 0, v_frame_stride, v_chan_stride, Cv, mask,
                               temperature, topk, non_mask_len, mode, return_topk)
    if return_topk:
        return r[0][0], r[1][0], r[2][0]
    return r[0]


def masked_attention(query, key, value, mask, temperature, topk, normalize=True, non_mask_len=0, mode='softmax',
                     return_topk=False):
    """Reference-shaped call: query [N,C,H,W], key [N,C,T,H,W], value [N,Cv,T,H,W] fp32 CUDA -> [N,Cv,H,W].
    All N batch items run in one launch (chunks of 32 items / 256 key frames)."""
    for t, n in ((query, 'query'), (key, 'key'), (value, 'value')):
        _require_cuda(t, n)
    N, C, H, W = query.shape
    T, Cv = key.shape[2], value.shape[1]
    assert key.shape[3:] == (H, W), 'vfs_b200: query and key feature maps must have the same size'
    HW = H * W
    qs = features_to_split(query.float(), normalize)                                              # [2,N,H,W,C]
    ks = features_to_split(key.float().transpose(1, 2).reshape(N * T, C, H, W).contiguous(), normalize)
    v = value.float().contiguous()                                                                # [N,Cv,T,H,W]
    per = max(1, min(32, 256 // T))
    outs, tvs, tis = [], [], []
    for b0 in range(0, N, per):
        bs = list(range(b0, min(N, b0 + per)))
        # NOTE reference quirk kept for parity: local_attention.py:320-326 flattens value_vec over the batch and
        # gathers with per-item indices that are NOT offset by the batch, so every batch item reads the value
        # maps of batch item 0 (invisible in VFS, which always calls with N == 1) -> v_batch_stride = 0.
        r = attention_bank_batched(qs, bs, ks, [[b * T + t for t in range(T)] for b in bs], v,
                                   [list(range(T)) for _ in bs], 0, HW, T * HW, Cv, mask, temperature, topk,
                                   non_mask_len, mode, return_topk)
        if return_topk:
            outs.append(r[0]); tvs.append(r[1]); tis.append(r[2])
        else:
            outs.append(r)
    out = torch.cat(outs).reshape(N, Cv, H, W)
    if return_topk:
        return out, torch.cat(tvs), torch.cat(tis)
    return out


GENERIC_ATTN_DEBUG = None


def masked_attention_generic(query, key, value, mask, temperature, topk, normalize=True, non_mask_len=0,
                             mode='softmax'):
    """masked_attention_efficient for an arbitrary boolean ``mask`` tensor ([HWk,HWq], or [N,HWk,HWq] with T == 1)
    and / or ``topk=None``: the [T*HWk, HWq] affinity of each batch item is materialised by the tcgen05 conv kernel
    (key pixels as the image, query pixels as 1x1 filters, scale = 1/temperature) and finished by
    vfs_generic_attention (csrc/dense.cu).  O(T*HW^2) memory -- the fused window kernel is the production path."""
    for t, n in ((query, 'query'), (key, 'key'), (value, 'value')):
        _require_cuda(t, n)
    N, C, Hq, Wq = query.shape
    T, Hk, Wk = key.shape[2:]
    Cv = value.shape[1]
    HWq, HWk = Hq * Wq, Hk * Wk
    rows = T * HWk
    Cp, HWqp = -(-C // 64) * 64, -(-HWq // 64) * 64
    if rows * HWqp * 4 > 16 * 2**30:
        raise NotImplementedError(f'vfs_b200: the general attention path would materialise {rows}x{HWqp} affinities')
    if topk is not None and not 1 <= topk <= 16:
        raise NotImplementedError('vfs_b200.masked_attention_efficient: topk must be in [1, 16] or None')
    dev = query.device
    m8 = None
    if mask is not None:
        m8 = (mask != 0).to(device=dev, dtype=torch.uint8).contiguous()
        assert m8.shape[-2:] == (HWk, HWq), (m8.shape, HWk, HWq)
    out = torch.empty((N, Cv, HWq), dtype=torch.float32, device=dev)
    scale = torch.full((HWqp, ), 1.0 / float(temperature), dtype=torch.float32, device=dev)
    shift = _const_vec(0, HWqp, dev)
    inv_ws = torch.empty((max(rows, HWq), ), dtype=torch.float32, device=dev)
    for b in range(N):
        a_split = torch.zeros((2, T, Hk, Wk, Cp), dtype=torch.float16, device=dev)
        w_split = torch.zeros((2, HWqp, Cp), dtype=torch.float16, device=dev)
        kb = key[b].float().transpose(0, 1).contiguous()                                # [T,C,Hk,Wk]
        what = 'features_to_split_norm' if normalize else 'features_to_split'
        check(nat.lib().vfs_features_to_split_ex(ptr(kb), ptr(a_split), ptr(inv_ws), T, C, Hk, Wk, int(normalize), Cp,
                                                 a_split.stride(0), current_stream()), what)
        check(nat.lib().vfs_features_to_split_ex(ptr(query[b:b + 1].float().contiguous()), ptr(w_split), ptr(inv_ws),
                                                 1, C, Hq, Wq, int(normalize), Cp, w_split.stride(0),
                                                 current_stream()), what)
        _, aff = conv_bn_act(a_split, w_split, scale, shift, 1, 1, 1, relu=False, want_split=False, want_f32=True)
        if GENERIC_ATTN_DEBUG is not None:      # tools/dense_flake_probe.py: intermediates of the last call
            GENERIC_ATTN_DEBUG.update(a_split=a_split, w_split=w_split, aff=aff, scale=scale, shift=shift)
        # reference quirk (local_attention.py:320-326): with top-k every batch item gathers the values of item 0
        vals = value[0 if topk is not None else b].float().reshape(Cv, rows).contiguous()
        mb = None if m8 is None else (m8 if m8.ndim == 2 else m8[b])
        check(nat.lib().vfs_generic_attention(ptr(aff), rows, HWqp, HWk, HWq, ptr(mb), int(non_mask_len), ptr(vals),
                                              Cv, int(topk or 0), 0 if mode == 'softmax' else 1, ptr(out[b]),
                                              current_stream()), 'generic_attention')
        if GENERIC_ATTN_DEBUG is not None:
            GENERIC_ATTN_DEBUG.update(mask=mb, vals=vals, out=out[b])
    return out.reshape(N, Cv, Hq, Wq)


def seg_postprocess(seg_logit, fh, fw, out_hw, out=None):
    """Propagated logits fp32 [Cv, fh*fw] (or a batch [P, Cv, fh*fw]) -> uint8 label map [H, W] ([P, H, W]):
    bilinear upsample + per-channel min-max + argmax."""
    _require_cuda(seg_logit, 'seg_logit')
    assert seg_logit.dtype == torch.float32 and seg_logit.is_contiguous()
    batched = seg_logit.ndim == 3
    P = seg_logit.shape[0] if batched else 1
    Cv = seg_logit.shape[-2]
    H, W = out_hw
    if out is None:
        out = torch.empty((P, H, W) if batched else (H, W), dtype=torch.uint8, device=seg_logit.device)
    ws = torch.empty((P * 2 * Cv, ), dtype=torch.int32, device=seg_logit.device)
    check(nat.lib().vfs_seg_postprocess_batched(ptr(seg_logit), ptr(out), ptr(ws), P, Cv, fh, fw, H, W,
                                                current_stream()), 'seg_postprocess')
    return out


def siamfc_response_peak(responses, hann_window, upscaled_size, scale_penalty, window_influence):
    """responses fp32 [S,R,R] (CUDA), hann_window fp64 [U,U] (CUDA, normalised) -> int32[3] device tensor
    {scale id, peak row, peak column} of the upsampled, penalised, window-blended response (csrc/post.cu)."""
    _require_cuda(responses, 'responses')
    _require_cuda(hann_window, 'hann_window')
    assert responses.dtype == torch.float32 and responses.ndim == 3 and responses.shape[1] == responses.shape[2]
    assert hann_window.dtype == torch.float64 and tuple(hann_window.shape) == (upscaled_size, upscaled_size)
    S, R = responses.shape[0], responses.shape[1]
    responses, hann_window = responses.contiguous(), hann_window.contiguous()
    ws = torch.empty((nat.lib().vfs_siamfc_peak_workspace_bytes(S, upscaled_size), ), dtype=torch.uint8,
                     device=responses.device)
    out = torch.empty((3, ), dtype=torch.int32, device=responses.device)
    check(nat.lib().vfs_siamfc_response_peak(ptr(responses), S, R, int(upscaled_size), ptr(hann_window),
                                             float(scale_penalty), float(window_influence), ptr(ws), ptr(out),
                                             current_stream()), 'siamfc_response_peak')
    return out


def dense_affinity(src_img, dst_img, temperature=1., normalize=True, softmax_dim=None, mask=None):
    """compute_affinity (affinity_utils.py:6-30): [B,C,h,w] x [B,C,h',w'] -> [B, hw, h'w'] fp32.  The GEMM runs on
    the tcgen05 conv kernel (src pixels = image, dst pixels = 1x1 filter bank, scale = 1/temperature); the mask fill
    and the softmax along ``softmax_dim`` are one more kernel."""
    for t, n in ((src_img, 'src_img'), (dst_img, 'dst_img')):
        _require_cuda(t, n)
    B, C = src_img.shape[:2]
    hs, ws_ = src_img.shape[2:]
    hd, wd = dst_img.shape[2:]
    HWs, HWd = hs * ws_, hd * wd
    if mask is not None:
        from .common.affinity_utils import NeighborMask
        if not isinstance(mask, NeighborMask):
            raise NotImplementedError('vfs_b200.compute_affinity: only masks built by spatial_neighbor() are supported')
        assert HWs == HWd and (mask.height, mask.width) == (hd, wd)
    assert softmax_dim in (None, 1, 2)
    Cp, HWp = -(-C // 64) * 64, -(-HWd // 64) * 64
    dev = src_img.device
    out = torch.empty((B, HWs, HWd), dtype=torch.float32, device=dev)
    scale = torch.full((HWp, ), 1.0 / float(temperature), dtype=torch.float32, device=dev)
    shift = _const_vec(0, HWp, dev)
    inv_ws = torch.empty((max(HWs, HWd), ), dtype=torch.float32, device=dev)
    for b in range(B):
        a_split = torch.zeros((2, 1, hs, ws_, Cp), dtype=torch.float16, device=dev)
        w_split = torch.zeros((2, HWp, Cp), dtype=torch.float16, device=dev)
        check(nat.lib().vfs_features_to_split_ex(ptr(src_img[b:b + 1].float().contiguous()), ptr(a_split), ptr(inv_ws),
                                                 1, C, hs, ws_, int(normalize), Cp, a_split.stride(0),
                                                 current_stream()), 'features_to_split_norm' if normalize else
              'features_to_split')
        check(nat.lib().vfs_features_to_split_ex(ptr(dst_img[b:b + 1].float().contiguous()), ptr(w_split), ptr(inv_ws),
                                                 1, C, hd, wd, int(normalize), Cp, w_split.stride(0),
                                                 current_stream()), 'features_to_split_norm' if normalize else
              'features_to_split')
        _, aff = conv_bn_act(a_split, w_split, scale, shift, 1, 1, 1, relu=False, want_split=False, want_f32=True)
        check(nat.lib().vfs_masked_softmax(ptr(aff), ptr(out[b]), 1, HWs, HWd, HWp, int(softmax_dim or 0),
                                           _MASK_MODES[mask.mode if mask is not None else None],
                                           mask.radius_y if mask is not None else 0,
                                           mask.radius_x if mask is not None else 0, wd, int(mask is not None),
                                           current_stream()), 'masked_softmax')
    return out


def propagate_dense(img, affinity, topk=None):
    """propagate (affinity_utils.py:33-50): img [B,Cv,h,w], affinity [B,hw,hw] -> [B,Cv,h,w]."""
    for t, n in ((img, 'img'), (affinity, 'affinity')):
        _require_cuda(t, n)
    B, Cv, h, w = img.shape
    assert tuple(affinity.shape) == (B, h * w, h * w), (affinity.shape, img.shape)
    out = torch.empty((B, Cv, h, w), dtype=torch.float32, device=img.device)
    check(nat.lib().vfs_propagate_dense(ptr(img.float().contiguous()), ptr(affinity.float().contiguous()), ptr(out), B,
                                        Cv, h * w, int(topk or 0), current_stream()), 'propagate_dense')
    return out
