"""Whole-step CUDA graph of the SimSiam training step, single- and multi-rank.

The eager step issues ~1400 kernels from Python (forward of two views, native backward, optimiser); at tens of
microseconds of interpreter + ctypes time per launch the host, not the B200, sets the step time.
``GraphedTrainStep`` captures forward + loss + backward + gradient all-reduce + optimiser step once over static
buffers (torch's "whole network capture" recipe) and replays it with one launch per step -- CUDA streams and graphs
instead of a tracing compiler.  The runner contract stays the reference's: ``step(data_batch)`` returns
``dict(loss, log_vars, num_samples)`` like ``BaseTracker.train_step`` (mmaction/models/trackers/base.py:113-156), with
the parameter update already applied (what mmcv's ``OptimizerHook`` does after it).

Multi-rank (the reference's MMDistributedDataParallel + SyncBN setting, mmaction/apis/train.py:58-66, configs/*:9,15):
every collective of the step is a kernel over NVLink peer memory (vfs_b200.peer / csrc/comm.cu) -- the per-layer SyncBN
statistic exchanges, the logged-scalar average and the two-shot gradient all-reduce on the flat gradient buffer -- so
the captured graph contains no NCCL call and the ranks run in lock-step without the host.

Parameters, gradients and momentum live in the flat buffers of ``vfs_b200.dp.FlatTrainState``; the SGD update is one
launch that reads {lr, momentum, weight_decay} from device memory, so a per-iteration LR schedule (the configs'
``lr_config = dict(policy='CosineAnnealing', by_epoch=False)``) is followed by writing ``optimizer.param_groups`` as
usual: ``__call__`` copies the four floats before the replay, nothing is re-captured.

Restrictions of graph capture: fixed input shapes (one graph per shape), no host-side control flow that depends on
device values."""
from collections import OrderedDict

import torch
import torch.distributed as dist

from . import ops, peer
from .dp import _layout
from .trackers.base import _batch_size, _reduce_entry


class GraphedTrainStep:

    def __init__(self, model, optimizer, data_batch, warmup=2, comm=None):
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200.GraphedTrainStep needs a CUDA device')
        if not hasattr(optimizer, 'attach_flat'):
            raise TypeError('vfs_b200.GraphedTrainStep needs a vfs_b200.optim.SGD optimizer')
        self.model, self.optimizer = model, optimizer
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in data_batch.items()}
        self.num_samples = _batch_size(data_batch)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.comm = None
        if self.world > 1:
            self.comm = comm or peer.active()
            if self.comm is None:
                group_params = [p for p in optimizer.param_groups[0]['params'] if p.requires_grad]
                self.comm = peer.install(peer.PeerComm(data_bytes=_layout(group_params)[1] * 4))
            elif peer.active() is None:
                peer.install(self.comm)
        self.flat = optimizer.flat if optimizer.flat is not None else optimizer.attach_flat(self.comm)
        self._keys = None
        self._capture(warmup)

    # one eager step on the current stream; returns (total loss, packed logged scalars)
    def _step(self):
        model = self.model
        model.iteration += 1
        self.flat.zero_grad()
        losses = model(**self.static)
        reduced = OrderedDict((name, _reduce_entry(name, value)) for name, value in losses.items())
        total = sum(v for name, v in reduced.items() if 'loss' in name)
        reduced['loss'] = total
        self._keys = list(reduced.keys())
        total.backward()
        packed = torch.stack([v.detach().float().reshape(()) for v in reduced.values()]).contiguous()
        if self.world > 1:
            self.flat.allreduce_grads(average=True)
            packed = (packed / self.world).contiguous()
            ops.cross_rank_sum_(packed)
        self.flat.sgd_step()
        return total.detach(), packed

    def _sync_hyper(self):
        g = self.optimizer.param_groups[0]
        self.flat.set_hyper(g['lr'], g['momentum'], g['weight_decay'], getattr(self.optimizer, 'grad_scale', 1.0))

    def _capture(self, warmup):
        # Warm-up steps (they build plans and kernel attributes) are real optimizer steps: snapshot the model and undo
        # them, so that constructing the graph does not train the model.
        self._sync_hyper()
        snapshot = None
        if warmup > 0:
            tensors = list(self.model.parameters()) + list(self.model.buffers())
            snapshot = [(t, t.detach().clone()) for t in tensors]
            momentum = self.flat.flat_momentum.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snapshot is not None:
            with torch.no_grad():
                for t, saved in snapshot:
                    t.copy_(saved)
                self.flat.flat_momentum.copy_(momentum)
            ops.WEIGHT_EPOCH[0] += 1
        engine = getattr(getattr(self.model, 'backbone', None), 'engine', None)
        if engine is not None:
            engine.invalidate()              # every weight pack is (re)issued inside the captured step
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.packed = self._step()

    def set_lr(self, lrs):
        """Convenience for schedules: same effect as writing ``optimizer.param_groups[i]['lr']``."""
        lrs = list(lrs) if isinstance(lrs, (list, tuple)) else [lrs] * len(self.optimizer.param_groups)
        for g, lr in zip(self.optimizer.param_groups, lrs):
            g['lr'] = lr

    def __call__(self, data_batch, log=True):
        for k, v in data_batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=True)
        self._sync_hyper()                   # LR / momentum / weight-decay hooks write param_groups; 16-byte H2D
        self.graph.replay()
        ops.WEIGHT_EPOCH[0] += 1             # parameters and BN buffers changed without tensor-version bumps
        log_vars = None
        if log:
            log_vars = OrderedDict(zip(self._keys, self.packed.tolist()))                # the step's only D2H
            if self.comm is not None:
                self.comm.check()
        return dict(loss=self.loss, log_vars=log_vars, num_samples=self.num_samples)
