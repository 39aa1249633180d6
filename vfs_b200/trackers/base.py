"""Tracker base class with the reference's interface (mmaction/models/trackers/base.py:12-178): constructor
arguments, the ``forward(imgs, return_loss=...)`` switch, and ``train_step`` / ``val_step`` returning
``dict(loss, log_vars, num_samples)`` for mmcv's runner."""
import abc
from collections import OrderedDict

import torch
import torch.distributed as dist
from torch import nn

from .. import builder


def _batch_size(data_batch):
    return len(next(iter(data_batch.values())))


def _reduce_entry(name, value):
    if torch.is_tensor(value):
        return value.mean()
    if isinstance(value, list):
        return sum(item.mean() for item in value)
    raise TypeError(f'{name} is not a tensor or list of tensors')


class BaseTracker(nn.Module, abc.ABC):
    """Holds ``backbone`` (built from its config) and optionally ``cls_head``; the buffer ``iteration`` counts
    ``train_step`` calls and is part of the checkpoint, like in the reference."""

    def __init__(self, backbone, cls_head=None, train_cfg=None, test_cfg=None):
        nn.Module.__init__(self)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.fp16_enabled = False
        self.backbone = builder.build_backbone(backbone)
        if cls_head is not None:
            self.cls_head = builder.build_head(cls_head)
        self.init_weights()
        self.register_buffer('iteration', torch.tensor(0, dtype=torch.float))

    # -- to be provided by the concrete tracker ---------------------------------------------------------
    @abc.abstractmethod
    def forward_train(self, imgs, labels):
        ...

    @abc.abstractmethod
    def forward_test(self, imgs, **kwargs):
        ...

    # -- shared behaviour -----------------------------------------------------------------------------------
    @property
    def with_cls_head(self):
        return getattr(self, 'cls_head', None) is not None

    def init_weights(self):
        for part in (self.backbone, self.cls_head if self.with_cls_head else None):
            if part is not None:
                part.init_weights()

    def extract_feat(self, imgs):
        return self.backbone(imgs)

    def forward(self, imgs, return_loss=True, **kwargs):
        step = self.forward_train if return_loss else self.forward_test
        return step(imgs, **kwargs)

    @staticmethod
    def _parse_losses(losses):
        """Loss dict -> (total loss tensor, OrderedDict of python floats).  Every entry is averaged; the total is the
        sum of the entries whose key contains 'loss'; logged values are averaged over ranks (reference
        base.py:76-110).  The reference does one all-reduce and one ``.item()`` per key; here all keys are packed
        into one vector: ONE all-reduce and ONE device->host copy per step, same logged values."""
        reduced = OrderedDict((name, _reduce_entry(name, value)) for name, value in losses.items())
        total = sum(v for name, v in reduced.items() if 'loss' in name)
        reduced['loss'] = total
        packed = torch.stack([v.detach().float().reshape(()) for v in reduced.values()])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .. import ops
            packed /= dist.get_world_size()
            ops.cross_rank_sum_(packed)     # peer-memory kernel when a communicator is installed, else all_reduce
        log_vars = OrderedDict(zip(reduced.keys(), packed.tolist()))
        return total, log_vars

    def _step_outputs(self, losses, data_batch):
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=_batch_size(data_batch))

    def train_step(self, data_batch, optimizer, **kwargs):
        self.iteration += 1
        return self._step_outputs(self(**data_batch), data_batch)

    def val_step(self, data_batch, optimizer, **kwargs):
        losses = self(data_batch['imgs'], data_batch['ref_seg_map'], data_batch['img_meta'])
        return self._step_outputs(losses, data_batch)
