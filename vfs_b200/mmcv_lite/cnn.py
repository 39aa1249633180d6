"""Stand-in for the ``mmcv.cnn`` symbols used on the VFS hot path (reference call sites:
mmaction/models/backbones/resnet.py:4, heads/sim_siam_head.py:2).

``ConvModule`` here is a *parameter container* with mmcv-full 1.2.1's attribute names (``conv``, ``bn``,
``activate``) so checkpoints keep their keys (``layer1.0.conv1.conv.weight``, ``...bn.running_mean``).
It holds no arithmetic: the conv -> BN -> ReLU(+residual) chain is executed as ONE fused tcgen05 kernel by
``vfs_b200.engine`` (csrc/conv_tc.cu).  Calling ``forward`` directly is an error, not a fallback.
"""
import torch.nn as nn

_NORM_TYPES = {
    'BN': nn.BatchNorm2d,
    'BN1d': nn.BatchNorm1d,
    'BN2d': nn.BatchNorm2d,
    'SyncBN': nn.SyncBatchNorm,
}


def build_norm_layer(cfg, num_features, postfix=''):
    """Returns ``(name, layer)`` like mmcv: name is ``'bn' + postfix`` for every batch-norm flavour."""
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict')
    if 'type' not in cfg:
        raise KeyError('the cfg dict must contain the key "type"')
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    if layer_type not in _NORM_TYPES:
        raise KeyError(f'Unrecognized norm type {layer_type}')
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    layer = _NORM_TYPES[layer_type](num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return 'bn' + str(postfix), layer


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    assert distribution in ('uniform', 'normal')
    if getattr(module, 'weight', None) is not None:
        if distribution == 'uniform':
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.normal_(module.weight, mean, std)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


class ConvModule(nn.Module):
    """conv -> norm -> act parameter bundle (mmcv ``ConvModule`` naming and defaults)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True):
        super().__init__()
        assert conv_cfg is None or isinstance(conv_cfg, dict)
        assert norm_cfg is None or isinstance(norm_cfg, dict)
        assert act_cfg is None or isinstance(act_cfg, dict)
        if conv_cfg is not None and conv_cfg.get('type', 'Conv2d') not in ('Conv', 'Conv2d'):
            raise KeyError(f"Unrecognized conv type {conv_cfg['type']}")
        if act_cfg is not None and act_cfg.get('type') != 'ReLU':
            raise KeyError(f"Unrecognized activation type {act_cfg.get('type')}")
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            self.activate = nn.ReLU(inplace=act_cfg.get('inplace', inplace))
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name)

    # geometry is read from the conv at execution time so change_stride() keeps working
    @property
    def kernel_size(self):
        return self.conv.kernel_size

    @property
    def stride(self):
        return self.conv.stride

    @property
    def dilation(self):
        return self.conv.dilation

    @property
    def padding(self):
        return self.conv.padding

    def init_weights(self):
        kaiming_init(self.conv, a=0, nonlinearity='relu')
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def forward(self, x):
        raise RuntimeError('vfs_b200 ConvModule is a parameter container: conv+BN+ReLU run fused inside the '
                           'tcgen05 engine (vfs_b200.engine); there is no eager/CPU fallback.')
