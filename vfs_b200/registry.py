"""Registries of the drop-in boundary -- same names as mmaction/models/registry.py:3-9 so the reference's
``configs/*.py`` (``type='ResNet'``, ``'SimSiamHead'``, ``'CosineSimLoss'``, ``'SimSiamBaseTracker'``,
``'VanillaTracker'``) resolve to the B200-native classes.  The package owns its registries (mmcv raises on
duplicate registration)."""
from .mmcv_lite import Registry

BACKBONES = Registry('backbone')
HEADS = Registry('head')
RECOGNIZERS = Registry('recognizer')
LOSSES = Registry('loss')
LOCALIZERS = Registry('localizer')
TRACKERS = Registry('tracker')
DROP_LAYERS = Registry('drop_layer')
