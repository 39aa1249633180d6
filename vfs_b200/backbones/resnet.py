"""B200-native 2-D ResNet backbone behind the reference's registry/config API.

Mirrors the *interface* of mmaction/models/backbones/resnet.py (class names, constructor kwargs and their
validation :346-378, state-dict keys, ``forward``/``forward_block``/``switch_strides``/
``switch_out_indices``/``output_stride``/``train``/``init_weights``), while the arithmetic runs in
``vfs_b200.engine`` : fp32 SIMT stem (csrc/stem.cu) and one fused tcgen05 implicit-GEMM kernel per
conv->BN->ReLU(+residual) (csrc/conv_tc.cu).  The nn.Modules below only own parameters.
"""
from functools import partial

import numpy as np
import torch
from torch import nn
from torch.nn.modules.batchnorm import _BatchNorm

from ..common.utils import change_stride
from ..mmcv_lite import ConvModule, constant_init, kaiming_init
from ..registry import BACKBONES


class BasicBlock(nn.Module):
    """3x3(stride, dilation) -> 3x3 residual block (reference resnet.py:15-113)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 conv_cfg=dict(type='Conv'), norm_cfg=dict(type='BN', requires_grad=True),
                 act_cfg=dict(type='ReLU', inplace=True), with_cp=False):
        super().__init__()
        assert style in ['pytorch', 'caffe']
        common = dict(bias=False, conv_cfg=conv_cfg, norm_cfg=norm_cfg)
        # only the first conv carries stride and dilation (:51-61); the second is always d=1 (:63-73)
        self.conv1 = ConvModule(inplanes, planes, kernel_size=3, stride=stride, padding=dilation,
                                dilation=dilation, act_cfg=act_cfg, **common)
        self.conv2 = ConvModule(planes, planes, kernel_size=3, stride=1, padding=1, dilation=1, act_cfg=None,
                                **common)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.style, self.stride, self.dilation = style, stride, dilation
        self.norm_cfg, self.with_cp = norm_cfg, with_cp

    def native_forward(self, engine, xs):
        identity = xs if self.downsample is None else engine.conv(self.downsample, xs, relu=False)
        y = engine.conv(self.conv1, xs, relu=True)
        return engine.conv(self.conv2, y, relu=True, residual=identity)


class Bottleneck(nn.Module):
    """1x1 -> 3x3(stride in 'pytorch' style) -> 1x1(x4) residual block (reference resnet.py:116-232)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 conv_cfg=dict(type='Conv'), norm_cfg=dict(type='BN', requires_grad=True),
                 act_cfg=dict(type='ReLU', inplace=True), with_cp=False):
        super().__init__()
        assert style in ['pytorch', 'caffe']
        self.inplanes, self.planes = inplanes, planes
        self.conv1_stride, self.conv2_stride = (1, stride) if style == 'pytorch' else (stride, 1)
        common = dict(bias=False, conv_cfg=conv_cfg, norm_cfg=norm_cfg)
        self.conv1 = ConvModule(inplanes, planes, kernel_size=1, stride=self.conv1_stride, act_cfg=act_cfg, **common)
        self.conv2 = ConvModule(planes, planes, kernel_size=3, stride=self.conv2_stride, padding=dilation,
                                dilation=dilation, act_cfg=act_cfg, **common)
        self.conv3 = ConvModule(planes, planes * self.expansion, kernel_size=1, act_cfg=None, **common)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride, self.dilation = stride, dilation
        self.norm_cfg, self.with_cp = norm_cfg, with_cp

    def native_forward(self, engine, xs):
        identity = xs if self.downsample is None else engine.conv(self.downsample, xs, relu=False)
        y = engine.conv(self.conv1, xs, relu=True)
        y = engine.conv(self.conv2, y, relu=True)
        return engine.conv(self.conv3, y, relu=True, residual=identity)


def make_res_layer(block, inplanes, planes, blocks, stride=1, dilation=1, style='pytorch', conv_cfg=None,
                   norm_cfg=None, act_cfg=None, with_cp=False):
    """One ResNet stage (reference resnet.py:235-306): a projection shortcut when shape changes; the first
    block of a dilated stage uses ``dilation // 2`` (:285)."""
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = ConvModule(inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False,
                                conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=None)
    kw = dict(style=style, conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg, with_cp=with_cp)
    first_dilation = dilation if dilation == 1 else dilation // 2
    layers = [block(inplanes, planes, stride, first_dilation, downsample, **kw)]
    inplanes = planes * block.expansion
    layers += [block(inplanes, planes, 1, dilation, **kw) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


@BACKBONES.register_module()
class ResNet(nn.Module):
    """ResNet-{18,34,50,101,152} feature extractor; kwargs as in the reference (resnet.py:346-363)."""

    arch_settings = {
        18: (BasicBlock, (2, 2, 2, 2)),
        34: (BasicBlock, (3, 4, 6, 3)),
        50: (Bottleneck, (3, 4, 6, 3)),
        101: (Bottleneck, (3, 4, 23, 3)),
        152: (Bottleneck, (3, 8, 36, 3)),
    }

    def __init__(self, depth, pretrained=None, torchvision_pretrain=True, in_channels=3, num_stages=4,
                 strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(3, ), style='pytorch',
                 frozen_stages=-1, conv_cfg=dict(type='Conv'), norm_cfg=dict(type='BN2d', requires_grad=True),
                 act_cfg=dict(type='ReLU', inplace=True), norm_eval=False, partial_bn=False, with_cp=False,
                 zero_init_residual=True):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError(f'invalid depth {depth} for resnet')
        assert 1 <= num_stages <= 4
        assert len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages
        self.depth, self.in_channels = depth, in_channels
        self.pretrained, self.torchvision_pretrain = pretrained, torchvision_pretrain
        self.num_stages, self.strides, self.dilations = num_stages, strides, dilations
        self.out_indices = self.original_out_indices = out_indices
        self.style, self.frozen_stages = style, frozen_stages
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.norm_eval, self.partial_bn, self.with_cp = norm_eval, partial_bn, with_cp
        self.zero_init_residual = zero_init_residual

        self.block, stage_blocks = self.arch_settings[depth]
        self.stage_blocks = stage_blocks[:num_stages]
        self.inplanes = 64
        self._make_stem_layer()
        self.res_layers = []
        for i, num_blocks in enumerate(self.stage_blocks):
            planes = 64 * 2**i
            layer = make_res_layer(self.block, self.inplanes, planes, num_blocks, stride=strides[i],
                                   dilation=dilations[i], style=style, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                   act_cfg=act_cfg, with_cp=with_cp)
            self.inplanes = planes * self.block.expansion
            name = f'layer{i + 1}'
            self.add_module(name, layer)
            self.res_layers.append(name)
        self._freeze_stages()
        self.feat_dim = self.block.expansion * 64 * 2**(len(self.stage_blocks) - 1)
        self._engine = None

    def _make_stem_layer(self):
        self.conv1 = ConvModule(self.in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False,
                                conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

    # ------------------------------------------------------------------ checkpoints / init
    def _load_torchvision_checkpoint(self, state_dict_tv, logger=None):
        """torchvision key names -> ConvModule names (inverse of tools/convert_weights/convert_to_pretrained.py;
        reference resnet.py:488-523)."""
        if 'state_dict' in state_dict_tv:
            state_dict_tv = state_dict_tv['state_dict']
        loaded = []
        for name, module in self.named_modules():
            if not isinstance(module, ConvModule):
                continue
            if 'downsample' in name:
                conv_name, bn_name = name + '.0', name + '.1'
            else:
                conv_name, bn_name = name, name.replace('conv', 'bn')
            module.conv.weight.data.copy_(state_dict_tv[conv_name + '.weight'])
            loaded.append(conv_name + '.weight')
            if module.conv.bias is not None:
                module.conv.bias.data.copy_(state_dict_tv[conv_name + '.bias'])
                loaded.append(conv_name + '.bias')
            for pname, p in list(module.bn.named_parameters()) + list(module.bn.named_buffers()):
                key = f'{bn_name}.{pname}'
                if key in state_dict_tv:
                    p.data.copy_(state_dict_tv[key])
                    loaded.append(key)
        return sorted(set(state_dict_tv.keys()) - set(loaded))

    def init_weights(self):
        if isinstance(self.pretrained, str):
            ckpt = torch.load(self.pretrained, map_location='cpu')
            if self.torchvision_pretrain:
                self._load_torchvision_checkpoint(ckpt)
            else:
                sd = ckpt.get('state_dict', ckpt)
                sd = {k[len('backbone.'):] if k.startswith('backbone.') else k: v for k, v in sd.items()}
                self.load_state_dict(sd, strict=False)
        elif self.pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    kaiming_init(m)
                elif isinstance(m, nn.BatchNorm2d):
                    constant_init(m, 1)
            if self.zero_init_residual:
                for m in self.modules():
                    if isinstance(m, Bottleneck):
                        constant_init(m.conv3.norm, 0)
                    elif isinstance(m, BasicBlock):
                        constant_init(m.conv2.norm, 0)
        else:
            raise TypeError('pretrained must be a str or None')

    # ------------------------------------------------------------------ execution
    @property
    def engine(self):
        if self._engine is None:
            from ..engine import BackboneEngine
            self._engine = BackboneEngine(self)
        return self._engine

    def _blocks(self):
        for i, name in enumerate(self.res_layers):
            for block in getattr(self, name):
                yield i, block

    def forward(self, x):
        """NCHW fp32 CUDA tensor -> stage outputs listed in ``out_indices`` (tensor if one, else tuple).  With
        gradients enabled and trainable parameters the call is recorded for the native backward pass
        (vfs_b200/autograd.py: BN/ReLU backward, tcgen05 dgrad + wgrad)."""
        out_indices = tuple(self.out_indices)
        params = [p for p in self.parameters() if p.requires_grad]
        if torch.is_grad_enabled() and params and self.training:
            # (a module in eval mode is treated as inference: its outputs carry no autograd history)
            self._check_trainable()
            from ..autograd import BackboneFunction
            outs = BackboneFunction.apply(x, self.engine, out_indices, *params)
        else:
            outs = self.engine.forward(x, out_indices=out_indices)
        return outs[0] if len(outs) == 1 else tuple(outs)

    def _check_trainable(self):
        """The native backward pass covers batch-statistics BN (the configs' pre-training, configs/*:9-11) and
        eval-mode BN inside the residual stages (norm_eval / frozen-BN fine-tuning).  The 7x7 stem with an eval-mode BN
        and trainable parameters has no backward kernel (the reference freezes the stem whenever it freezes anything,
        resnet.py:593-609): fail loudly."""
        cm = self.conv1
        if cm.conv.weight.requires_grad and cm.with_norm and not cm.norm.training:
            raise NotImplementedError('vfs_b200: backward through the stem with an eval-mode BatchNorm is not '
                                      'implemented; freeze the stem (frozen_stages >= 0) or keep its BN in train mode')

    def forward_block(self, x, index):
        return self.engine.forward(x, block_index=index)[0]

    @property
    def output_stride(self):
        return np.prod(self.strides[:self.num_stages]) * 4

    # ------------------------------------------------------------------ mode / stride switches
    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.conv1.eval()
            for p in self.conv1.parameters():
                p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f'layer{i}')
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def _partial_bn(self):
        seen = 0
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                seen += 1
                if seen >= 2:
                    m.eval()
                    m.weight.requires_grad = False
                    m.bias.requires_grad = False

    def switch_strides(self, strides=None):
        """Change the stride of the first block of every stage in place (reference resnet.py:624-637)."""
        for i, name in enumerate(self.res_layers):
            for m in getattr(self, name).modules():
                if isinstance(m, (BasicBlock, Bottleneck)) and m.downsample is not None:
                    stride = self.strides[i] if strides is None else strides[i]
                    m.downsample.apply(partial(change_stride, stride=stride))
                    strided = m.conv1 if (self.depth in [18, 34] or self.style != 'pytorch') else m.conv2
                    strided.apply(partial(change_stride, stride=stride))

    def switch_out_indices(self, out_indices=None):
        self.out_indices = self.original_out_indices if out_indices is None else out_indices

    def train(self, mode=True):
        """NOTE: returns None like the reference (resnet.py:645-654 has no ``return self``)."""
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, _BatchNorm):
                    m.eval()
        if mode and self.partial_bn:
            self._partial_bn()
