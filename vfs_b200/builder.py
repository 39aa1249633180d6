"""Factory functions of the drop-in boundary.  Same public names and call signatures as the reference's
``mmaction.models.builder`` (``build_backbone(cfg)``, ``build_model(cfg, train_cfg, test_cfg)`` ...), generated from
one table: each entry names the registry it draws from and whether the train/test configs are injected as default
constructor arguments (the reference does that for recognizers and trackers only)."""
import torch.nn as nn

from . import registry as _reg
from .mmcv_lite import build_from_cfg

_FAMILIES = {
    # public suffix: (registry, receives train_cfg/test_cfg)
    'backbone': (_reg.BACKBONES, False),
    'head': (_reg.HEADS, False),
    'drop_layer': (_reg.DROP_LAYERS, False),
    'loss': (_reg.LOSSES, False),
    'localizer': (_reg.LOCALIZERS, False),
    'recognizer': (_reg.RECOGNIZERS, True),
    'tracker': (_reg.TRACKERS, True),
}
# order in which build_model looks a type name up (reference builder.py:68-78)
_MODEL_FAMILIES = ('localizer', 'recognizer', 'tracker')


def build(cfg, registry, default_args=None):
    """One config -> one module; a list of configs -> ``nn.Sequential`` of them."""
    if isinstance(cfg, list):
        return nn.Sequential(*(build_from_cfg(item, registry, default_args) for item in cfg))
    return build_from_cfg(cfg, registry, default_args)


def _make_builder(family):
    registry, wants_cfgs = _FAMILIES[family]
    if wants_cfgs:
        def builder(cfg, train_cfg=None, test_cfg=None):
            return build(cfg, registry, dict(train_cfg=train_cfg, test_cfg=test_cfg))
    else:
        def builder(cfg):
            return build(cfg, registry)
    builder.__name__ = f'build_{family}'
    builder.__doc__ = f'Instantiate a {family} from its config dict (type name looked up in {registry.name!r}).'
    return builder


build_backbone = _make_builder('backbone')
build_head = _make_builder('head')
build_drop_layer = _make_builder('drop_layer')
build_loss = _make_builder('loss')
build_localizer = _make_builder('localizer')
build_recognizer = _make_builder('recognizer')
build_tracker = _make_builder('tracker')


def build_model(cfg, train_cfg=None, test_cfg=None):
    """Top-level entry used by tools/train.py and tools/test.py: dispatch on which registry knows ``cfg['type']``."""
    type_name = dict(cfg)['type']
    for family in _MODEL_FAMILIES:
        registry, wants_cfgs = _FAMILIES[family]
        if type_name in registry:
            maker = globals()[f'build_{family}']
            return maker(cfg, train_cfg, test_cfg) if wants_cfgs else maker(cfg)
    raise KeyError(f'{type_name} not in any registry')
