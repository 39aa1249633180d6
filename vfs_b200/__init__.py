"""vfs_b200 -- B200-native (sm_100a) implementation of the xvjiarui/VFS hot path behind the reference's
mmaction2/mmcv registry + config API.  Importing the package registers ``ResNet``, ``SimSiamHead``,
``CosineSimLoss``, ``SimSiamBaseTracker`` and ``VanillaTracker`` in the package-owned registries
(``vfs_b200.registry``) so the reference's ``configs/*.py`` build unchanged through ``build_model``."""
from . import backbones, heads, losses, trackers  # noqa: F401  (registration side effects)
from .graphs import GraphedTrainStep  # noqa: F401
from .pipelines import DeviceNormalizeFormat, DeviceTrainAugment, PinnedRing  # noqa: F401
from .builder import (build_backbone, build_head, build_loss, build_model, build_tracker)  # noqa: F401
from .mmcv_lite import Config, ConfigDict  # noqa: F401
from .registry import BACKBONES, HEADS, LOSSES, TRACKERS  # noqa: F401

__version__ = '0.1.0'
