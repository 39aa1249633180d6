"""Weighted-loss base class (interface of mmaction/models/losses/base.py:6-37)."""
from abc import ABCMeta, abstractmethod

import torch.nn as nn


class BaseWeightedLoss(nn.Module, metaclass=ABCMeta):
    """Subclasses implement ``_forward``; ``forward`` multiplies by ``loss_weight``."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    @abstractmethod
    def _forward(self, *args, **kwargs):
        pass

    def forward(self, *args, **kwargs):
        return self._forward(*args, **kwargs) * self.loss_weight
