"""``build_*`` helpers with the call signatures of mmaction/models/builder.py:8-78."""
import torch.nn as nn

from .mmcv_lite import build_from_cfg
from .registry import BACKBONES, DROP_LAYERS, HEADS, LOCALIZERS, LOSSES, RECOGNIZERS, TRACKERS


def build(cfg, registry, default_args=None):
    """A list of configs becomes an ``nn.Sequential`` (builder.py:24-29)."""
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_head(cfg):
    return build(cfg, HEADS)


def build_drop_layer(cfg):
    return build(cfg, DROP_LAYERS)


def build_loss(cfg):
    return build(cfg, LOSSES)


def build_recognizer(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, RECOGNIZERS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_tracker(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, TRACKERS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_localizer(cfg):
    return build(cfg, LOCALIZERS)


def build_model(cfg, train_cfg=None, test_cfg=None):
    """Dispatch on registry membership of ``cfg['type']`` (builder.py:68-78)."""
    obj_type = dict(cfg)['type']
    if obj_type in LOCALIZERS:
        return build_localizer(cfg)
    if obj_type in RECOGNIZERS:
        return build_recognizer(cfg, train_cfg, test_cfg)
    if obj_type in TRACKERS:
        return build_tracker(cfg, train_cfg, test_cfg)
    raise KeyError(f'{obj_type} not in any registry')
