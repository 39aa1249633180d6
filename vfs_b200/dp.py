"""Flat parameter / gradient / momentum buffers of the data-parallel training step.

The reference wraps the model in MMDistributedDataParallel (mmaction/apis/train.py:58-66): gradients are copied into
25 MB buckets, all-reduced over NCCL while backward runs, copied back, and torch.optim.SGD then walks the parameters
one by one.  Here the three per-parameter tensor families live in three flat buffers (every parameter at a 16-byte
aligned offset):

  * ``flat_grad``   inside the peer communicator's symmetric data region when one is given -- the native backward
                    kernels (tcgen05 wgrad, BN backward, head backward) ACCUMULATE straight into it (``p.grad`` is a
                    view, like DDP's gradient_as_bucket_view), one memset per step replaces the per-parameter zeroing
                    and the view-2 accumulation kernels, and the gradient all-reduce is one two-shot kernel over NVLink
                    peer memory (csrc/comm.cu) on that buffer in place;
  * ``flat_param``  ``p.data`` are views, so the SGD update of all parameters is ONE kernel launch reading its
                    hyper-parameters from device memory (LR schedules work under CUDA-graph replay);
  * ``flat_momentum``  zero-initialised (torch's first-step rule buf = g is momentum*0 + g).
"""
import torch

from . import ops
from ._native import current_stream, ptr
from . import _native as nat


def _layout(params, align=4):
    offsets, n = [], 0
    for p in params:
        offsets.append(n)
        n += -(-p.numel() // align) * align
    return offsets, n


class FlatTrainState:

    def __init__(self, params, comm=None, lr=0.05, momentum=0.9, weight_decay=1e-4):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('FlatTrainState: no trainable parameters')
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('vfs_b200.dp.FlatTrainState needs CUDA parameters (no CPU fallback)')
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError('FlatTrainState: parameters must be float32 on one device')
        self.offsets, self.numel = _layout(self.params)
        self.comm = comm
        self.world = comm.world if comm is not None else 1
        if comm is not None and self.world > 1:
            if comm.data_bytes < self.numel * 4:
                raise RuntimeError(f'peer communicator data region ({comm.data_bytes} B) is smaller than the flat '
                                   f'gradient buffer ({self.numel * 4} B)')
            self.flat_grad = comm.data()[:self.numel * 4].view(torch.float32)
            self.flat_grad.zero_()
        else:
            self.flat_grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.flat_param = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.flat_momentum = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.hyper = torch.tensor([lr, momentum, weight_decay, 1.0], dtype=torch.float32, device=dev)
        self._hyper_host = torch.tensor([lr, momentum, weight_decay, 1.0], dtype=torch.float32).pin_memory()
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + n].view(p.shape)
                p.grad = self.flat_grad[off:off + n].view(p.shape)
                ops.GRAD_SINKS[p.data_ptr()] = p.grad
                ops.GRAD_SINK_OWNER[p.data_ptr()] = self
        ops.WEIGHT_EPOCH[0] += 1

    def close(self):
        for p in self.params:
            ops.GRAD_SINKS.pop(p.data_ptr(), None)
            ops.GRAD_SINK_OWNER.pop(p.data_ptr(), None)

    # ------------------------------------------------------------------ per step
    def zero_grad(self):
        self.flat_grad.zero_()

    def set_hyper(self, lr, momentum, weight_decay, grad_scale=1.0):
        """Write the optimiser hyper-parameters into the device vector the (possibly captured) SGD launch reads."""
        new = torch.tensor([float(lr), float(momentum), float(weight_decay), float(grad_scale)], dtype=torch.float32)
        if not torch.equal(self._hyper_host, new):
            self._hyper_host.copy_(new)
            self.hyper.copy_(self._hyper_host, non_blocking=True)

    def allreduce_grads(self, average=True):
        """Sum (average) the flat gradient over the ranks: peer-memory two-shot kernel, or torch.distributed."""
        import torch.distributed as dist
        if self.comm is not None and self.world > 1:
            self.comm.allreduce_(self.flat_grad, 1.0 / self.world if average else 1.0)
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_grad)
            if average:
                self.flat_grad /= dist.get_world_size()

    def sgd_step(self):
        ops.check(nat.lib().vfs_sgd_momentum_step_dev(ptr(self.flat_param), ptr(self.flat_grad),
                                                      ptr(self.flat_momentum), self.numel, ptr(self.hyper),
                                                      current_stream()), 'sgd_momentum_step_dev')
        ops.WEIGHT_EPOCH[0] += 1      # cached packed weights / folded BN of every engine are stale now
