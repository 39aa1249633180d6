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


def load_reference_pipelines():
    """mmaction/datasets/pipelines/{augmentations,formating}.py imported UNCHANGED (RandomResizedCrop, Resize, Flip,
    Normalize, FormatShape ...).  The three mmcv image helpers they call are restated with the cv2 calls mmcv-full 1.2.1
    makes (mmcv/image/geometric.py: imresize -> cv2.resize, imflip_ -> cv2.flip in place; photometric.py: imnormalize_
    -> cv2.subtract / cv2.multiply with float64 row vectors); skimage (absent, used only by an unrelated transform) is
    stubbed."""
    import cv2
    import numpy as np
    load_reference()
    mmcv = sys.modules['mmcv']
    codes = {'nearest': cv2.INTER_NEAREST, 'bilinear': cv2.INTER_LINEAR, 'bicubic': cv2.INTER_CUBIC,
             'area': cv2.INTER_AREA, 'lanczos': cv2.INTER_LANCZOS4}
    pillow_imresize = mmcv.imresize

    def imresize(img, size, return_scale=False, interpolation='bilinear', out=None, backend=None):
        if backend == 'pillow':
            return pillow_imresize(img, size, interpolation=interpolation, backend='pillow')
        h, w = img.shape[:2]
        resized = cv2.resize(img, size, dst=out, interpolation=codes[interpolation])
        if not return_scale:
            return resized
        return resized, size[0] / w, size[1] / h

    def imflip_(img, direction='horizontal'):
        assert direction in ['horizontal', 'vertical']
        return cv2.flip(img, 1 if direction == 'horizontal' else 0, img)

    def imnormalize_(img, mean, std, to_rgb=True):
        assert img.dtype != np.uint8
        mean = np.float64(mean.reshape(1, -1))
        stdinv = 1 / np.float64(std.reshape(1, -1))
        if to_rgb:
            cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
        cv2.subtract(img, mean, img)
        cv2.multiply(img, stdinv, img)
        return img

    def is_tuple_of(seq, expected_type):
        return isinstance(seq, tuple) and all(isinstance(x, expected_type) for x in seq)

    mmcv.imresize, mmcv.imflip_, mmcv.imnormalize_, mmcv.is_tuple_of = imresize, imflip_, imnormalize_, is_tuple_of
    mmcv.imflip = lambda img, direction='horizontal': np.flip(img, axis=1 if direction == 'horizontal' else 0)
    mmcv.rescale_size = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    mmcv.parallel.DataContainer = type('DataContainer', (), {'__init__': lambda self, data, **k: setattr(self, 'data', data)})
    mmcv.is_str = lambda x: isinstance(x, str)
    if 'skimage' not in sys.modules:
        sk = types.ModuleType('skimage')
        sk.__path__ = []
        sku = types.ModuleType('skimage.util')
        sku.view_as_windows = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
        sk.util = sku
        sys.modules['skimage'], sys.modules['skimage.util'] = sk, sku
    ds = os.path.join(REF_ROOT, 'mmaction', 'datasets')
    _stub_pkg('mmaction.datasets', ds)
    importlib.import_module('mmaction.datasets.registry')
    _stub_pkg('mmaction.datasets.pipelines', os.path.join(ds, 'pipelines'))
    aug = importlib.import_module('mmaction.datasets.pipelines.augmentations')
    fmt = importlib.import_module('mmaction.datasets.pipelines.formating')
    return types.SimpleNamespace(augmentations=aug, formating=fmt)
