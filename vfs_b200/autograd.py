"""torch.autograd glue of the native training path.

The reference trains through stock autograd (``loss.backward()`` in mmcv's OptimizerHook after
``BaseTracker.train_step``, trackers/base.py:119-156).  To stay drop-in for that runner contract the native forward
pass is exposed as ``torch.autograd.Function``s whose backward runs the hand-written kernels:

  BackboneFunction      stem + residual stages (train-mode BN) ; backward = BN/ReLU backward kernels, tcgen05
                        dgrad (conv_tc.cu) and wgrad (wgrad_tc.cu) per ConvModule, SyncBN sums all-reduced
  LinearBnActFunction   Linear (+BatchNorm1d) (+ReLU) of the SimSiam head
  AvgPoolFunction       global average pool
  CosineLossFunction    2 - 2 cos(p, z) per sample (z is a constant, as in SimSiam's stop-gradient)

Nothing here computes in torch; tensors are device buffers handed to libvfs_b200.so.
"""
import torch

from . import ops

# Gradients travel through the tensor cores as split-fp16 values; they are multiplied by this power of two on entry
# to the backbone's backward pass (fp16 keeps 11+11 bits only above ~6e-5) and parameter gradients are divided by
# it on the way out.  Exact (power of two) unless a gradient overflows 65504/GRAD_SCALE, which the overflow
# counter reports.
GRAD_SCALE = 4096.0


class BackboneFunction(torch.autograd.Function):
    """forward(x, engine, out_indices, *params) -> tuple of NCHW fp32 stage outputs."""

    @staticmethod
    def forward(ctx, x, engine, out_indices, *params):
        tape = []
        engine.tape = tape
        try:
            outs, out_ids = engine.forward_taped(x, out_indices)
        finally:
            engine.tape = None
        ctx.engine, ctx.tape, ctx.out_ids = engine, tape, out_ids
        ctx.params = params
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        grads = ctx.engine.backward(ctx.tape, ctx.out_ids, grad_outs)
        ctx.tape = None  # free the saved activations
        return (None, None, None) + tuple(grads.get(id(p)) for p in ctx.params)


class LinearBnActFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn, relu):
        x = x.contiguous()
        y = ops.linear_forward(x, weight.detach(), bias.detach() if bias is not None else None)
        pre = mean = invstd = None
        training = False
        if bn is not None:
            pre = y.clone()
            mean, invstd, training = ops.bn1d_forward_(y, bn, relu)
        elif relu:
            ops.relu_(y)
        ctx.save_for_backward(x, weight, pre, y, mean, invstd, gamma)
        ctx.has_bn, ctx.relu, ctx.training, ctx.has_bias = bn is not None, relu, training, bias is not None
        ctx.bn = bn
        ctx.bias = bias
        ctx.need_dx = ctx.needs_input_grad[0]   # (x.requires_grad is False inside forward for a .contiguous() copy)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, pre, y, mean, invstd, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        dgamma = dbeta = None
        if ctx.has_bn:
            dpre, dgamma, dbeta = ops.bn1d_backward(dy, pre, y, gamma.detach() if gamma is not None else None, mean,
                                                    invstd, ctx.training, ctx.relu, bn=ctx.bn)
            if gamma is None:
                dgamma = dbeta = None
        elif ctx.relu:
            dpre = ops.relu_backward(dy, y)
        else:
            dpre = dy
        dW_sink = ops.grad_sink(weight)
        db_sink = ops.grad_sink(ctx.bias) if ctx.has_bias else None
        if dW_sink is not None and (db_sink is not None or not ctx.has_bias):
            dx, _, _ = ops.linear_backward(dpre, x, weight.detach(), need_dx=ctx.need_dx, dW_out=dW_sink,
                                           db_out=db_sink)
            return dx, None, None, dgamma, dbeta, None, None
        dx, dW, db = ops.linear_backward(dpre, x, weight.detach(), need_dx=ctx.need_dx)
        return dx, dW, (db if ctx.has_bias else None), dgamma, dbeta, None, None


class AvgPoolFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x):
        ctx.hw = tuple(x.shape[2:])
        return ops.global_avg_pool_raw(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.avgpool_backward(dy.contiguous(), ctx.hw)


class CosineLossFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, p, z, with_norm, negative):
        p, z = p.contiguous(), z.contiguous()
        ctx.save_for_backward(p, z)
        ctx.with_norm, ctx.negative = with_norm, negative
        return ops.cosine_sim_loss_raw(p, z, with_norm, negative)

    @staticmethod
    def backward(ctx, gout):
        p, z = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError('vfs_b200 CosineSimLoss: gradient w.r.t. the label is not implemented '
                                      '(SimSiam detaches it, sim_siam_head.py:171-173)')
        return ops.cosine_loss_backward(p, z, gout.contiguous(), ctx.with_norm, ctx.negative), None, None, None
