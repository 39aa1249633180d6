"""Registries of the drop-in boundary.  The registry *names* are the reference's (mmaction/models/registry.py:3-9) so
that its ``configs/*.py`` -- ``type='ResNet'``, ``'SimSiamHead'``, ``'CosineSimLoss'``, ``'SimSiamBaseTracker'``,
``'VanillaTracker'`` -- resolve to the B200-native classes.  This package owns its own registry objects: mmcv raises
on duplicate registration, so sharing the reference's would need ``force=True``."""
from .mmcv_lite import Registry

_KINDS = ('backbone', 'head', 'recognizer', 'loss', 'localizer', 'tracker', 'drop_layer')
(BACKBONES, HEADS, RECOGNIZERS, LOSSES, LOCALIZERS, TRACKERS, DROP_LAYERS) = (Registry(kind) for kind in _KINDS)

__all__ = ['BACKBONES', 'HEADS', 'RECOGNIZERS', 'LOSSES', 'LOCALIZERS', 'TRACKERS', 'DROP_LAYERS']
