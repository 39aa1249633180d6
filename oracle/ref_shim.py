"""ORACLE -- test infrastructure only.

Loads the *unmodified* reference hot-path modules from /root/reference (when present, i.e. only in the authoring
container) so that tests/golden/make_golden.py can pin the oracle against the reference itself.  mmcv is not
installed and there is no network, so a stand-in ``mmcv`` exposing exactly the symbols those files import is
placed in ``sys.modules`` (semantics of mmcv-full 1.2.1); heavyweight ``__init__`` files of
mmaction are bypassed by registering stub parent packages whose ``__path__`` points into the reference tree.
Nothing from the reference is copied into this repository.
"""
import importlib
import importlib.util
import logging
import os
import sys
import types

import torch.nn as nn

REF_ROOT = os.environ.get('VFS_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'mmaction', 'models'))


def _make_mmcv():
    from vfs_b200.mmcv_lite import Registry, build_from_cfg, ConfigDict
    from vfs_b200.mmcv_lite import cnn as lite

    class RefConvModule(nn.Module):
        """conv -> norm -> act with mmcv naming (``conv``, ``bn``, ``activate``) and a REAL forward."""

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                     bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, **kw):
            super().__init__()
            self.with_norm = norm_cfg is not None
            self.with_activation = act_cfg is not None
            if bias == 'auto':
                bias = not self.with_norm
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
            if self.with_norm:
                self.norm_name, norm = lite.build_norm_layer(norm_cfg, out_channels)
                self.add_module(self.norm_name, norm)
            if self.with_activation:
                self.activate = nn.ReLU(inplace=act_cfg.get('inplace', inplace))
            lite.kaiming_init(self.conv, a=0, nonlinearity='relu')
            if self.with_norm:
                lite.constant_init(self.norm, 1, bias=0)

        @property
        def norm(self):
            return getattr(self, self.norm_name)

        def forward(self, x):
            x = self.conv(x)
            if self.with_norm:
                x = self.norm(x)
            if self.with_activation:
                x = self.activate(x)
            return x

    def auto_fp16(*a, **k):
        return lambda f: f

    def imresize(img, size, interpolation='nearest', backend='pillow', **kw):
        from PIL import Image
        import numpy as np
        assert backend == 'pillow' and interpolation == 'nearest'
        return np.array(Image.fromarray(img).resize(size, Image.NEAREST))

    mmcv = types.ModuleType('mmcv')
    mmcv.__path__ = []
    utils = types.ModuleType('mmcv.utils')
    utils.Registry, utils.build_from_cfg = Registry, build_from_cfg
    utils.get_logger = lambda name, log_file=None, log_level=logging.INFO: logging.getLogger(name)
    utils.print_log = lambda *a, **k: None
    utils.collect_env = lambda: {}
    utils._BatchNorm = nn.modules.batchnorm._BatchNorm
    utils.SyncBatchNorm = nn.SyncBatchNorm
    utils._ConvNd = nn.modules.conv._ConvNd
    cnn = types.ModuleType('mmcv.cnn')
    cnn.ConvModule = RefConvModule
    cnn.build_norm_layer = lite.build_norm_layer
    cnn.kaiming_init, cnn.constant_init, cnn.normal_init = lite.kaiming_init, lite.constant_init, lite.normal_init
    cnn.CONV_LAYERS = Registry('conv layer')
    cnn.build_plugin_layer = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    cnn.build_activation_layer = lambda cfg: nn.ReLU(inplace=cfg.get('inplace', False))
    cnn.NonLocal3d = type('NonLocal3d', (nn.Module, ), {})
    runner = types.ModuleType('mmcv.runner')
    runner.auto_fp16 = auto_fp16
    runner._load_checkpoint = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    runner.load_checkpoint = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    runner.get_dist_info = lambda: (0, 1)
    parallel = types.ModuleType('mmcv.parallel')
    parallel.collate = lambda *a, **k: None
    mmcv.utils, mmcv.cnn, mmcv.runner, mmcv.parallel = utils, cnn, runner, parallel
    mmcv.ConfigDict = ConfigDict
    mmcv.mkdir_or_exist = lambda d, mode=0o777: os.makedirs(d, mode=mode, exist_ok=True)
    mmcv.imresize = imresize
    mmcv.BaseStorageBackend = object

    class FileClient:
        @staticmethod
        def register_backend(*a, **k):
            return lambda cls: cls

    mmcv.FileClient = FileClient
    for m in (mmcv, utils, cnn, runner, parallel):
        sys.modules[m.__name__] = m


def _stub_pkg(name, path):
    pkg = types.ModuleType(name)
    pkg.__path__ = [path]
    sys.modules[name] = pkg
    return pkg


_loaded = None


def load_reference():
    """Returns a namespace with the reference's ResNet, SimSiamHead, CosineSimLoss, trackers, common ops,
    build_model and the SiamFC heads -- all imported from the files under /root/reference unchanged."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    if 'mmcv' in sys.modules and not hasattr(sys.modules['mmcv'], '__vfs_shim__'):
        pass  # a real mmcv would also do
    _make_mmcv()
    sys.modules['mmcv'].__vfs_shim__ = True
    mm = os.path.join(REF_ROOT, 'mmaction')
    _stub_pkg('mmaction', mm)
    importlib.import_module('mmaction.utils')
    models = _stub_pkg('mmaction.models', os.path.join(mm, 'models'))
    importlib.import_module('mmaction.models.registry')
    importlib.import_module('mmaction.models.builder')
    common = importlib.import_module('mmaction.models.common')
    backbones = _stub_pkg('mmaction.models.backbones', os.path.join(mm, 'models', 'backbones'))
    resnet = importlib.import_module('mmaction.models.backbones.resnet')
    backbones.ResNet = resnet.ResNet
    _stub_pkg('mmaction.models.losses', os.path.join(mm, 'models', 'losses'))
    importlib.import_module('mmaction.models.losses.base')
    sim_loss = importlib.import_module('mmaction.models.losses.sim_loss')
    _stub_pkg('mmaction.models.heads', os.path.join(mm, 'models', 'heads'))
    head = importlib.import_module('mmaction.models.heads.sim_siam_head')
    trackers = importlib.import_module('mmaction.models.trackers')
    builder = sys.modules['mmaction.models.builder']
    spec = importlib.util.spec_from_file_location(
        'ref_siamfc_heads', os.path.join(REF_ROOT, 'projects', 'siamfc-pytorch', 'siamfc', 'heads.py'))
    siamfc_heads = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(siamfc_heads)
    ns = types.SimpleNamespace(
        ResNet=resnet.ResNet, SimSiamHead=head.SimSiamHead, CosineSimLoss=sim_loss.CosineSimLoss,
        SimSiamBaseTracker=trackers.SimSiamBaseTracker, VanillaTracker=trackers.VanillaTracker,
        build_model=builder.build_model, common=common, siamfc_heads=siamfc_heads, models=models,
        masked_attention_efficient=common.masked_attention_efficient, spatial_neighbor=common.spatial_neighbor,
        compute_affinity=common.compute_affinity, propagate=common.propagate)
    _loaded = ns
    return ns


def load_reference_siamfc_ops():
    """projects/siamfc-pytorch/siamfc/ops.py (+ bbox_utils.py, image_utils.py: plain cv2 / numpy) imported unchanged
    through a stub parent package, for pinning the crop helper of the SiamFC tracker."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    name = 'ref_siamfc_pkg'
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF_ROOT, 'projects', 'siamfc-pytorch', 'siamfc')]
        sys.modules[name] = pkg
    return importlib.import_module(name + '.ops')


def load_reference_siamfc_tracker():
    """projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py imported UNCHANGED (TrackerSiamFC.init / update / track,
    Net, _convert_batchnorm).  Its only missing imports are ``got10k.trackers.Tracker`` (a trivial base class holding
    ``name`` / ``is_deterministic``), a few mmcv helpers the inference path never calls, and sibling modules of the
    project (datasets / transforms / losses: plain torch / cv2 / numpy, loaded from the reference tree as they are)."""
    ref = load_reference()
    if 'got10k' not in sys.modules:
        got10k = types.ModuleType('got10k')
        got10k.__path__ = []
        trackers = types.ModuleType('got10k.trackers')

        class Tracker:                                   # got10k/trackers/__init__.py: name + determinism flag
            def __init__(self, name, is_deterministic=False):
                self.name, self.is_deterministic = name, is_deterministic

        trackers.Tracker = Tracker
        got10k.trackers = trackers
        sys.modules['got10k'], sys.modules['got10k.trackers'] = got10k, trackers
    mmcv = sys.modules['mmcv']
    mmcv.parallel.is_module_wrapper = lambda m: False
    mmcv.runner.save_checkpoint = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    if not hasattr(mmcv, 'ProgressBar'):
        mmcv.ProgressBar = type('ProgressBar', (), {'__init__': lambda self, n: None, 'update': lambda self: None})
    models = sys.modules['mmaction.models']
    models.ResNet = ref.ResNet
    models.build_backbone = sys.modules['mmaction.models.builder'].build_backbone
    name = 'ref_siamfc_pkg'
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF_ROOT, 'projects', 'siamfc-pytorch', 'siamfc')]
        sys.modules[name] = pkg
    return importlib.import_module(name + '.siamfc_tracker_base')
