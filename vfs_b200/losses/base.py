"""Loss base class with the reference's contract (mmaction/models/losses/base.py): subclasses provide ``_forward``;
calling the module returns ``loss_weight * _forward(...)``."""
import abc

from torch import nn


class BaseWeightedLoss(nn.Module, abc.ABC):

    def __init__(self, loss_weight=1.0):
        nn.Module.__init__(self)
        self.loss_weight = loss_weight

    @abc.abstractmethod
    def _forward(self, *args, **kwargs):
        """Unweighted loss."""

    def forward(self, *args, **kwargs):
        unweighted = self._forward(*args, **kwargs)
        return unweighted * self.loss_weight
