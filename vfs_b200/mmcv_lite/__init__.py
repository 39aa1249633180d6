"""Self-contained stand-in for the few mmcv symbols the VFS hot path needs (mmcv is not installed here and
the reference pins mmcv-full 1.2.1, docker/Dockerfile:80)."""
from .cnn import ConvModule, build_norm_layer, constant_init, kaiming_init, normal_init
from .config import Config, ConfigDict
from .registry import Registry, build_from_cfg

__all__ = ['Registry', 'build_from_cfg', 'Config', 'ConfigDict', 'ConvModule', 'build_norm_layer',
           'kaiming_init', 'constant_init', 'normal_init']
