"""Whole-step CUDA graph of the SimSiam training step.

The eager step issues ~1400 kernels from Python (forward of two views, native backward, fused SGD per parameter); at
~45 us of interpreter + ctypes time per launch the host, not the B200, sets the step time (profiles/
r01_comparators_v11.json: 68.6 ms per step).  ``GraphedTrainStep`` captures forward + loss + backward + gradient
all-reduce + optimizer step once over static buffers (torch's "whole network capture" recipe) and replays it with one
launch per step -- CUDA streams and graphs instead of a tracing compiler.  The runner contract stays the reference's:
``step(data_batch)`` returns ``dict(loss, log_vars, num_samples)`` like ``BaseTracker.train_step`` (trackers/base.py
:113-156 in the reference), with the parameter update already applied (what ``OptimizerHook`` does after it).

Restrictions of graph capture: fixed input shapes (one graph per shape), no host-side control flow that depends on
device values, parameters / optimizer hyper-parameters changed from the host (e.g. an LR schedule) must be written
into the tensors the graph reads -- ``set_lr`` re-captures when the learning rate changes."""
from collections import OrderedDict

import torch
import torch.distributed as dist

from .optim import allreduce_grads
from .trackers.base import _batch_size, _reduce_entry


class GraphedTrainStep:

    def __init__(self, model, optimizer, data_batch, warmup=2):
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200.GraphedTrainStep needs a CUDA device')
        self.model, self.optimizer = model, optimizer
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in data_batch.items()}
        self.num_samples = _batch_size(data_batch)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1:
            # measured on 2 x B200 (round 1): capturing the ~330 SyncBN / gradient NCCL collectives of the step
            # dead-locks inside the capture; multi-rank training runs the eager step until the collectives are fused
            # into the BN kernels
            raise NotImplementedError('vfs_b200.GraphedTrainStep: multi-rank capture is not supported yet, '
                                      'use model.train_step eagerly')
        self._keys = None
        self._capture(warmup)

    # one eager step on the current stream; returns (total loss, packed logged scalars)
    def _step(self):
        model = self.model
        model.iteration += 1
        losses = model(**self.static)
        reduced = OrderedDict((name, _reduce_entry(name, value)) for name, value in losses.items())
        total = sum(v for name, v in reduced.items() if 'loss' in name)
        reduced['loss'] = total
        self._keys = list(reduced.keys())
        self.optimizer.zero_grad(set_to_none=True)
        total.backward()
        packed = torch.stack([v.detach().float().reshape(()) for v in reduced.values()])
        if self.world > 1:
            allreduce_grads(self.params, average=True)
            packed = packed / self.world
            dist.all_reduce(packed)
        self.optimizer.step()
        return total.detach(), packed

    def _capture(self, warmup):
        # Warm-up steps (they build plans, optimizer state and kernel attributes) are real optimizer steps: snapshot
        # the model and undo them, so that constructing the graph does not train the model.  Zeroed momentum buffers
        # are equivalent to absent ones (buf = momentum * 0 + grad on the first real step, dampening is 0).
        snapshot = None
        if warmup > 0:
            tensors = list(self.model.parameters()) + list(self.model.buffers())
            snapshot = [(t, t.detach().clone()) for t in tensors]
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
                for state in self.optimizer.state.values():
                    if 'momentum_buffer' in state:
                        state['momentum_buffer'].zero_()
        engine = getattr(getattr(self.model, 'backbone', None), 'engine', None)
        if engine is not None:
            engine.invalidate()              # every weight pack is (re)issued inside the captured step
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.packed = self._step()
        self._lr = [g['lr'] for g in self.optimizer.param_groups]

    def set_lr(self, lrs):
        """Learning rates are kernel arguments baked into the graph: re-capture when a schedule changes them."""
        lrs = list(lrs) if isinstance(lrs, (list, tuple)) else [lrs] * len(self.optimizer.param_groups)
        if lrs != self._lr:
            for g, lr in zip(self.optimizer.param_groups, lrs):
                g['lr'] = lr
            self._capture(0)

    def __call__(self, data_batch, log=True):
        for k, v in data_batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        log_vars = OrderedDict(zip(self._keys, self.packed.tolist())) if log else None   # the step's only D2H
        return dict(loss=self.loss, log_vars=log_vars, num_samples=self.num_samples)
