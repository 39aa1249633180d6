from .base import BaseWeightedLoss
from .sim_loss import CosineSimLoss

__all__ = ['BaseWeightedLoss', 'CosineSimLoss']
