"""Checkpoint interop: tracker checkpoint -> torchvision-style backbone checkpoint, the job of the
reference's tools/convert_weights/convert_to_pretrained.py:7-58 (its inverse is ``ResNet._load_torchvision_checkpoint``).

    python -m vfs_b200.convert  work_dirs/.../latest.pth  r50_vfs_torchvision.pth
"""
import re
import sys
from collections import OrderedDict

import torch

# (ConvModule-style key pattern, torchvision replacement), applied to keys with the ``backbone.`` prefix removed
_RULES = (
    (re.compile(r'^conv1\.conv\.(.+)$'), r'conv1.\1'),
    (re.compile(r'^conv1\.(bn|gn)\.(.+)$'), r'\g<1>1.\2'),
    (re.compile(r'^(layer\d+\.\d+)\.downsample\.conv\.(.+)$'), r'\1.downsample.0.\2'),
    (re.compile(r'^(layer\d+\.\d+)\.downsample\.(?:bn|gn)\.(.+)$'), r'\1.downsample.1.\2'),
    (re.compile(r'^(layer\d+\.\d+)\.conv(\d)\.conv\.(.+)$'), r'\1.conv\2.\3'),
    (re.compile(r'^(layer\d+\.\d+)\.conv(\d)\.(bn|gn)\.(.+)$'), r'\1.\g<3>\2.\4'),
)


def backbone_to_torchvision(state_dict):
    """``{'backbone.layer1.0.conv1.bn.weight': t, 'img_head...': ...}`` -> ``{'layer1.0.bn1.weight': t, ...}``; keys
    outside ``backbone.`` are dropped, an unknown backbone key raises like the reference script."""
    state_dict = state_dict.get('state_dict', state_dict)
    out = OrderedDict()
    for key, value in state_dict.items():
        if not key.startswith('backbone'):
            continue
        inner = key.replace('backbone.', '')
        for pattern, repl in _RULES:
            if pattern.match(inner):
                out[pattern.sub(repl, inner)] = value
                break
        else:
            raise RuntimeError(inner)
    return out


def convert(src, dst):
    """File-to-file form with the reference script's output layout: ``dict(state_dict=..., meta={})``."""
    torch.save(dict(state_dict=backbone_to_torchvision(torch.load(src, map_location='cpu')), meta=dict()), dst)

