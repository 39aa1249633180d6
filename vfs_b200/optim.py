"""Optimiser and data-parallel gradient plumbing of the training step.

``SGD`` has torch.optim.SGD's constructor and update rule (momentum, weight decay, dampening 0, no Nesterov --
the reference's ``optimizer = dict(type='SGD', lr=0.05, momentum=0.9, weight_decay=0.0001)``, configs/*:134) with the
update done by one fused CUDA kernel per parameter (csrc/train.cu), and ``allreduce_grads`` is the DDP-equivalent
gradient average over NCCL (one flat bucket; the 1/world factor is folded into the SGD kernel's grad_scale)."""
import torch
import torch.distributed as dist

from . import ops


class SGD(torch.optim.Optimizer):

    def __init__(self, params, lr=0.05, momentum=0.0, dampening=0, weight_decay=0.0, nesterov=False):
        if dampening != 0 or nesterov:
            raise NotImplementedError('vfs_b200.optim.SGD: dampening / nesterov are not used by the VFS configs')
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.grad_scale = 1.0  # set to 1/world_size when gradients were summed (not averaged) across ranks
        self.flat = None       # dp.FlatTrainState once attach_flat() was called: one fused launch per step

    def attach_flat(self, comm=None):
        """Move parameters, gradients and momentum into the flat buffers of ``vfs_b200.dp.FlatTrainState`` (``p.data``
        / ``p.grad`` become views): the native backward accumulates straight into the flat gradient, ``step`` is one
        kernel launch whose hyper-parameters live in device memory, ``zero_grad`` one memset.  With a peer
        communicator the gradient buffer lives in its symmetric region (``allreduce_grads`` = peer-memory kernel)."""
        from .dp import FlatTrainState
        if len(self.param_groups) != 1:
            raise NotImplementedError('vfs_b200.optim.SGD.attach_flat: one parameter group (every VFS config)')
        g = self.param_groups[0]
        params = [p for p in g['params'] if p.requires_grad]
        flat = FlatTrainState(params, comm, g['lr'], g['momentum'], g['weight_decay'])
        for p, off in zip(flat.params, flat.offsets):
            view = flat.flat_momentum[off:off + p.numel()].view(p.shape)
            old = self.state[p].get('momentum_buffer')
            if old is not None:
                view.copy_(old)
            self.state[p]['momentum_buffer'] = view
        self.flat = flat
        return flat

    def zero_grad(self, set_to_none=True):
        if self.flat is not None:
            self.flat.zero_grad()      # the gradient views stay in place
            return
        super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self.flat is not None:
            g = self.param_groups[0]
            self.flat.set_hyper(g['lr'], g['momentum'], g['weight_decay'], self.grad_scale)
            self.flat.sgd_step()
            return loss
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError('vfs_b200.optim.SGD needs CUDA parameters (no CPU fallback)')
                if not p.is_contiguous() or p.dtype != torch.float32:
                    raise RuntimeError('vfs_b200.optim.SGD: parameters must be contiguous float32 tensors')
                state = self.state[p]
                first = 'momentum_buffer' not in state
                if first:
                    state['momentum_buffer'] = torch.empty_like(p, memory_format=torch.contiguous_format)
                ops.sgd_momentum_step_(p.data, p.grad.contiguous(), state['momentum_buffer'], group['lr'],
                                       group['momentum'], group['weight_decay'], first, self.grad_scale)
                p.add_(0)  # bump the tensor version: cached packed weights are refreshed on the next forward
        return loss


def build_optimizer(model, cfg):
    """``cfg = dict(type='SGD', lr=..., momentum=..., weight_decay=...)`` like mmcv's build_optimizer."""
    cfg = dict(cfg)
    if cfg.pop('type') != 'SGD':
        raise KeyError('vfs_b200.optim: only SGD is provided natively (the optimiser of every VFS config)')
    return SGD([p for p in model.parameters() if p.requires_grad], **cfg)


def allreduce_grads(params, average=True):
    """Sum (average) the gradients of ``params`` over the default process group with ONE all-reduce of a flat
    bucket (NCCL over NVLink on the box; gloo in the CPU tests).  No-op without an initialised group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    from . import ops
    owners = {id(o): o for o in (ops.GRAD_SINK_OWNER.get(p.data_ptr()) for p in params)}
    if len(owners) == 1 and None not in [o for o in owners.values()]:
        # gradients are views of one flat buffer (dp.FlatTrainState): reduce it in place, peer-memory kernel or NCCL
        next(iter(owners.values())).allreduce_grads(average=average)
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
