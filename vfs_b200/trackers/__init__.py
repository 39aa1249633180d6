from .base import BaseTracker
from .sim_siam_base_tracker import SimSiamBaseTracker
from .vanilla_tracker import VanillaTracker

__all__ = ['BaseTracker', 'VanillaTracker', 'SimSiamBaseTracker']
