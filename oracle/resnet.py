"""ORACLE -- test infrastructure only (see oracle/__init__.py).

CPU restatement of the reference 2-D ResNet forward as pure functions over a state dict.  The reference's
arithmetic lives in torch ATen (conv2d / batch_norm / relu / max_pool2d called through mmcv ConvModule), so the
restatement calls the same ATen ops in the same order:

    ResNet.forward          mmaction/models/backbones/resnet.py:555-575
    _make_stem_layer        :422-435   conv7x7/s2/p3 -> BN -> ReLU ; MaxPool2d(3, 2, 1)
    Bottleneck.forward      :200-232   1x1 -> 3x3(stride, 'pytorch' style) -> 1x1 ; + identity ; ReLU
    BasicBlock.forward      :83-113    3x3(stride, dilation) -> 3x3 ; + identity ; ReLU
    make_res_layer          :235-306   downsample = 1x1/stride conv + BN; first block dilation // 2 (:285)
    ConvModule              mmcv-full 1.2.1: conv(bias=False) -> norm -> act; BN eps 1e-5, momentum 0.1

State-dict keys are the reference's (``conv1.conv.weight``, ``layer1.0.conv1.bn.running_mean`` ...).
Pinned against the reference itself by tests/golden/make_golden.py (run where /root/reference exists).
"""
import torch
import torch.nn.functional as F

ARCH = {
    18: ('basic', (2, 2, 2, 2)),
    34: ('basic', (3, 4, 6, 3)),
    50: ('bottleneck', (3, 4, 6, 3)),
    101: ('bottleneck', (3, 4, 23, 3)),
    152: ('bottleneck', (3, 8, 36, 3)),
}


def conv_module(x, sd, prefix, stride=1, padding=0, dilation=1, relu=True, bn_training=False):
    """mmcv ConvModule: conv -> BN -> (ReLU).  In training mode uses batch statistics (and, like torch, would
    update running stats; the oracle leaves the state dict untouched by cloning them)."""
    y = F.conv2d(x, sd[prefix + '.conv.weight'], sd.get(prefix + '.conv.bias'), stride, padding, dilation)
    if prefix + '.bn.weight' in sd:
        rm, rv = sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var']
        if bn_training:
            rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(y, rm, rv, sd[prefix + '.bn.weight'], sd[prefix + '.bn.bias'], bn_training, 0.1, 1e-5)
    return F.relu(y) if relu else y


def _block(x, sd, prefix, kind, stride, dilation, has_downsample, bn_training):
    identity = x
    if kind == 'bottleneck':
        out = conv_module(x, sd, prefix + '.conv1', 1, 0, 1, True, bn_training)
        out = conv_module(out, sd, prefix + '.conv2', stride, dilation, dilation, True, bn_training)
        out = conv_module(out, sd, prefix + '.conv3', 1, 0, 1, False, bn_training)
    else:
        out = conv_module(x, sd, prefix + '.conv1', stride, dilation, dilation, True, bn_training)
        out = conv_module(out, sd, prefix + '.conv2', 1, 1, 1, False, bn_training)
    if has_downsample:
        identity = conv_module(x, sd, prefix + '.downsample', stride, 0, 1, False, bn_training)
    return F.relu(out + identity)


def resnet_forward(sd, x, depth, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(3, ),
                   bn_training=False, block_index=None):
    """Returns the tensor (one index) or tuple of stage outputs, like ResNet.forward; with ``block_index`` the
    output of that residual block (forward_block, resnet.py:577-587)."""
    kind, stage_blocks = ARCH[depth]
    expansion = 4 if kind == 'bottleneck' else 1
    x = conv_module(x, sd, 'conv1', 2, 3, 1, True, bn_training)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    inplanes = 64
    bidx = 0
    for i, nblocks in enumerate(stage_blocks[:len(strides)]):
        planes = 64 * 2**i
        for b in range(nblocks):
            first = b == 0
            stride = strides[i] if first else 1
            dil = dilations[i]
            if first and dil != 1:
                dil = dil // 2
            has_ds = first and (strides[i] != 1 or inplanes != planes * expansion)
            x = _block(x, sd, f'layer{i + 1}.{b}', kind, stride, dil, has_ds, bn_training)
            if block_index is not None and bidx == block_index:
                return x
            bidx += 1
        inplanes = planes * expansion
        if i in out_indices:
            outs.append(x)
    return outs[0] if len(outs) == 1 else tuple(outs)


def seeded_state_dict(module_or_sd, seed=0, bn_affine=True):
    """Deterministic, init-order-independent parameter fill keyed by state-dict name (so the reference module,
    the B200 module and the oracle can be given identical weights from a seed alone):
    conv/linear weights ~ kaiming-normal(fan_out), BN gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1),
    running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5)  (zero-init-residual makes the raw init
    degenerate)."""
    import zlib
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    out = {}
    for name, t in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2**31))
        if name.endswith('num_batches_tracked'):
            out[name] = torch.zeros_like(t)
        elif name.endswith('running_var'):
            out[name] = torch.rand(t.shape, generator=g) + 0.5
        elif name.endswith('running_mean'):
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        elif t.ndim == 1 and name.endswith('.weight'):
            out[name] = (torch.rand(t.shape, generator=g) + 0.5) if bn_affine else torch.ones_like(t)
        elif t.ndim == 1:  # biases
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        elif t.ndim >= 2:
            fan_out = t.shape[0] * (t[0, 0].numel() if t.ndim > 2 else 1)
            out[name] = torch.randn(t.shape, generator=g) * (2.0 / fan_out) ** 0.5
        else:
            out[name] = t.clone()
    return out
