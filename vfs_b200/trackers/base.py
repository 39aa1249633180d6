"""Tracker base class -- interface of mmaction/models/trackers/base.py:12-178."""
from abc import ABCMeta, abstractmethod
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import builder


class BaseTracker(nn.Module, metaclass=ABCMeta):
    """Owns ``backbone`` (and optionally ``cls_head``); subclasses define ``forward_train`` / ``forward_test``."""

    def __init__(self, backbone, cls_head=None, train_cfg=None, test_cfg=None):
        super().__init__()
        self.backbone = builder.build_backbone(backbone)
        if cls_head is not None:
            self.cls_head = builder.build_head(cls_head)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.init_weights()
        self.fp16_enabled = False
        self.register_buffer('iteration', torch.tensor(0, dtype=torch.float))

    @property
    def with_cls_head(self):
        return hasattr(self, 'cls_head') and self.cls_head is not None

    def init_weights(self):
        self.backbone.init_weights()
        if self.with_cls_head:
            self.cls_head.init_weights()

    def extract_feat(self, imgs):
        return self.backbone(imgs)

    @abstractmethod
    def forward_train(self, imgs, labels):
        pass

    @abstractmethod
    def forward_test(self, imgs, **kwargs):
        pass

    @staticmethod
    def _parse_losses(losses):
        """mean every entry, sum the ones whose key contains 'loss', average logged scalars across ranks
        (reference base.py:76-110).  The per-key all-reduce + ``.item()`` of the reference is batched into ONE
        all-reduce and ONE device->host copy per step (same logged values, one host sync instead of 2+)."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        keys = list(log_vars)
        packed = torch.stack([log_vars[k].detach().float().reshape(()) for k in keys])
        if dist.is_available() and dist.is_initialized():
            packed = packed / dist.get_world_size()
            dist.all_reduce(packed)
        for k, v in zip(keys, packed.tolist()):
            log_vars[k] = v
        return loss, log_vars

    def forward(self, imgs, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(imgs, **kwargs)
        return self.forward_test(imgs, **kwargs)

    def train_step(self, data_batch, optimizer, **kwargs):
        """Returns ``dict(loss, log_vars, num_samples)`` (reference base.py:119-156)."""
        self.iteration += 1
        losses = self(**data_batch)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(next(iter(data_batch.values()))))

    def val_step(self, data_batch, optimizer, **kwargs):
        losses = self(data_batch['imgs'], data_batch['ref_seg_map'], data_batch['img_meta'])
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(next(iter(data_batch.values()))))
