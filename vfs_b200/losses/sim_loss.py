"""Cosine-similarity loss of the SimSiam head, computed by one fused CUDA kernel per call
(csrc/head.cu: L2-normalise both inputs, dot, 2 - 2*cos) -- interface of
mmaction/models/losses/sim_loss.py:25-63."""
from .. import ops
from ..registry import LOSSES
from .base import BaseWeightedLoss


@LOSSES.register_module()
class CosineSimLoss(BaseWeightedLoss):
    """``2 - 2 * cos(cls_score, label)`` per sample (``-cos`` if ``negative``)."""

    def __init__(self, with_norm=True, negative=False, pairwise=False, **kwargs):
        super().__init__(**kwargs)
        self.with_norm = with_norm
        self.negative = negative
        self.pairwise = pairwise

    def _forward(self, cls_score, label, mask=None, **kwargs):
        if mask is not None:
            assert self.pairwise
        if self.pairwise:
            raise NotImplementedError('vfs_b200: pairwise CosineSimLoss is not used by any VFS config and has no '
                                      'native kernel')
        return ops.cosine_sim_loss(cls_score, label, with_norm=self.with_norm, negative=self.negative)
