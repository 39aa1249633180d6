"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every test drives the CUDA path through the C ABI
(libvfs_b200.so via ctypes) and checks it against

  * the committed golden fixtures = outputs of the unmodified reference (tests/golden/vfs_golden.npz), and
  * the oracle (oracle/*.py, CPU restatement pinned to those fixtures) on seeded inputs,

within the north-star tolerance: 1e-3 relative for floating point (relative to the tensor's max magnitude for
feature maps, elementwise for scalars), exact top-k index sets modulo fp32 near-ties.
"""
import numpy as np
import pytest
import torch

import oracle
from tests.golden import cases

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3  # north_star: "within 1e-3 rel fp32"


def stable_oracle(fn, tries=5):
    """Evaluate a CPU-oracle expression until two consecutive evaluations are bit-identical (a cheap guard against
    host-side nondeterminism of torch-CPU; the GPU result is not involved in the vote).  History: the dense-softmax
    window case failed in ~1 of 10 fresh processes on the GPU boxes with the CUDA output right (1e-7 from an fp64
    re-evaluation) and the oracle 0.32 off; the vote did not catch it because the wrong value was reproducible inside
    the process -- the float32 ``(dy**2 + dx**2)**0.5 < r`` neighbour mask came out different in those processes.
    oracle.spatial_neighbor now uses integer arithmetic; the fp64 arbiter in the test below reports any recurrence."""
    prev = fn()
    for _ in range(tries):
        cur = fn()
        if torch.equal(prev, cur):
            return cur
        print('[oracle] two CPU evaluations of the oracle disagreed (max abs diff %.3e); re-evaluating'
              % float((prev - cur).abs().max()))
        prev = cur
    raise AssertionError('the CPU oracle does not give a reproducible result')


def rel_err(got, ref):
    got = torch.as_tensor(got).detach().double().cpu()
    ref = torch.as_tensor(ref).detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _load(module, sd):
    module.load_state_dict(sd)
    return module.cuda()


# --------------------------------------------------------------------------------------------- backbone
@pytest.mark.parametrize('name', sorted(cases.BACKBONE_CASES))
def test_backbone_eval_matches_reference_golden(golden, name):
    from vfs_b200.backbones import ResNet
    c = cases.BACKBONE_CASES[name]
    net = ResNet(c['depth'], norm_cfg=dict(type='SyncBN', requires_grad=True), strides=c['strides'],
                 dilations=c['dilations'], out_indices=c['out_indices'])
    net = _load(net, oracle.seeded_state_dict(net, seed=c['seed']))
    net.train(False)
    y = net(cases.backbone_input(c).cuda())
    ref = golden[f'backbone/{name}/eval']
    assert tuple(y.shape) == ref.shape
    assert rel_err(y, ref) < REL_TOL


@pytest.mark.parametrize('depth', [18, 50])
def test_backbone_480x854_davis_features_match_oracle(depth):
    """DAVIS inference geometry of the configs (480x854 frame, strides (1,2,1,1), out index 2 -> [1,C,60,107]; W = 107
    is prime, HW = 6420 is not a multiple of the 128-pixel tile) against the CPU oracle."""
    from vfs_b200.backbones import ResNet
    net = ResNet(depth, norm_cfg=dict(type='SyncBN', requires_grad=True), strides=(1, 2, 1, 1), out_indices=(2, ))
    sd = oracle.seeded_state_dict(net, seed=40 + depth)
    net = _load(net, sd)
    net.train(False)
    x = torch.randn(1, 3, 480, 854, generator=torch.Generator().manual_seed(depth))
    y = net(x.cuda())
    with torch.no_grad():
        ref = oracle.resnet_forward(sd, x, depth, strides=(1, 2, 1, 1), out_indices=(2, ))
    assert tuple(y.shape) == tuple(ref.shape) == (1, 1024 if depth == 50 else 256, 60, 107)
    assert rel_err(y, ref) < REL_TOL
    # elementwise view of the same bar: 99.9 % of the activations within 1e-3 of their own magnitude (+ 1e-3 of the
    # mean magnitude for values near zero)
    yc, rf = y.cpu().double(), ref.double()
    ok = (yc - rf).abs() <= REL_TOL * (rf.abs() + rf.abs().mean())
    assert float(ok.double().mean()) > 0.999


@pytest.mark.parametrize('name', sorted(cases.BACKBONE_CASES))
def test_backbone_train_mode_bn_matches_reference_golden(golden, name):
    """Batch-statistics BatchNorm (train mode): outputs and the running-statistics update."""
    from vfs_b200.backbones import ResNet
    c = cases.BACKBONE_CASES[name]
    net = ResNet(c['depth'], norm_cfg=dict(type='SyncBN', requires_grad=True), strides=c['strides'],
                 dilations=c['dilations'], out_indices=c['out_indices'])
    sd = oracle.seeded_state_dict(net, seed=c['seed'])
    net = _load(net, sd)
    net.train(True)
    x = cases.backbone_input(c)
    y = net(x.cuda())
    ref = golden[f'backbone/{name}/train']
    assert tuple(y.shape) == ref.shape
    assert rel_err(y, ref) < REL_TOL
    # running statistics of the stem BN after one step == torch's update rule on the oracle's stem conv output
    with torch.no_grad():
        z = torch.nn.functional.conv2d(x, sd['conv1.conv.weight'], None, 2, 3)
        exp_mean = 0.9 * sd['conv1.bn.running_mean'] + 0.1 * z.mean(dim=(0, 2, 3))
        exp_var = 0.9 * sd['conv1.bn.running_var'] + 0.1 * z.var(dim=(0, 2, 3), unbiased=True)
    assert rel_err(net.conv1.bn.running_mean, exp_mean) < REL_TOL
    assert rel_err(net.conv1.bn.running_var, exp_var) < REL_TOL
    assert int(net.conv1.bn.num_batches_tracked) == 1
    # and switching back to eval uses the UPDATED running statistics (cached folds are refreshed)
    net.train(False)
    y_eval = net(x.cuda())
    sd2 = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref_eval = oracle.resnet_forward(sd2, x, c['depth'], c['strides'], c['dilations'], c['out_indices'])
    assert rel_err(y_eval, ref_eval) < REL_TOL


@pytest.mark.parametrize('name', sorted(cases.TRACKER_TRAIN_CASES))
def test_simsiam_forward_train_matches_reference_golden(golden, name):
    """SimSiamBaseTracker.forward_train in TRAIN mode (batch-stat BN in backbone and head) built from the
    reference's model dicts through build_model, against the reference's own losses."""
    import vfs_b200
    c = cases.TRACKER_TRAIN_CASES[name]
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    model.load_state_dict(oracle.seeded_state_dict(model, seed=c['seed']))
    model = model.cuda()
    model.train()
    with torch.no_grad():
        losses = model.forward_train(cases.tracker_train_input(c).cuda())
    keys = [k for k in golden if k.startswith(f'tracker_train/{name}/')]
    assert len(keys) == len(losses)
    for k in keys:
        got = losses[k.split('/', 2)[2]]
        np.testing.assert_allclose(got.cpu().numpy(), golden[k], rtol=REL_TOL, atol=2e-5)


@pytest.mark.parametrize('depth,shape,strides,out_indices', [
    (50, (2, 3, 256, 256), (1, 2, 2, 2), (3, )),      # BASELINE cfg-2 frame size
    (50, (1, 3, 240, 432), (1, 2, 1, 1), (2, )),      # DAVIS-style res4, stride 8
    (18, (2, 3, 224, 224), (1, 2, 2, 2), (3, )),      # BASELINE cfg-1 frame size
    (50, (2, 3, 128, 160), (1, 2, 2, 2), (0, 1, 2, 3)),
])
def test_backbone_eval_matches_oracle(depth, shape, strides, out_indices):
    from vfs_b200.backbones import ResNet
    net = ResNet(depth, norm_cfg=dict(type='SyncBN', requires_grad=True), strides=strides,
                 out_indices=out_indices)
    sd = oracle.seeded_state_dict(net, seed=depth)
    net = _load(net, sd)
    net.train(False)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(depth + shape[2]))
    y = net(x.cuda())
    with torch.no_grad():
        ref = oracle.resnet_forward(sd, x, depth, strides, (1, 1, 1, 1), out_indices)
    if len(out_indices) == 1:
        y, ref = (y, ), (ref, )
    assert len(y) == len(ref)
    for a, b in zip(y, ref):
        assert rel_err(a, b) < REL_TOL


def test_backbone_api_contract():
    """Constructor validation / mode switches pinned by the reference's tests/test_models/test_backbone.py:26-103."""
    from vfs_b200.backbones import ResNet
    with pytest.raises(KeyError):
        ResNet(20)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=0)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=5)
    with pytest.raises(AssertionError):
        ResNet(50, strides=(1, ), dilations=(1, 1), num_stages=3)
    with pytest.raises(TypeError):
        ResNet(50, pretrained=0).init_weights()
    with pytest.raises(AssertionError):
        ResNet(18, style='tensorflow')
    net = ResNet(18, norm_eval=True)
    net.init_weights()
    net.train()
    assert all(not m.training for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm))
    net = ResNet(18, frozen_stages=1).cuda()
    net.init_weights()
    net.train()
    assert not net.conv1.bn.training and all(not p.requires_grad for p in net.layer1.parameters())
    net.train(False)
    assert tuple(net(torch.randn(1, 3, 64, 64).cuda()).shape) == (1, 512, 2, 2)       # test_backbone.py:109-113
    net50 = ResNet(50).cuda()
    net50.train(False)
    assert tuple(net50(torch.randn(1, 3, 64, 64).cuda()).shape) == (1, 2048, 2, 2)    # test_backbone.py:116-120
    # stride switching (resnet.py:624-637) changes the output resolution and is reversible
    net50.switch_strides((1, 2, 1, 1))
    net50.switch_out_indices((2, ))
    assert tuple(net50(torch.randn(1, 3, 64, 64).cuda()).shape) == (1, 1024, 8, 8)
    net50.switch_strides()
    net50.switch_out_indices()
    assert tuple(net50(torch.randn(1, 3, 64, 64).cuda()).shape) == (1, 2048, 2, 2)
    with pytest.raises(RuntimeError):
        net50(torch.randn(1, 3, 64, 64))  # CPU tensor: no fallback


# --------------------------------------------------------------------------------------------- conv kernel units
CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, dil, relu, residual
    (2, 16, 16, 64, 64, 1, 1, 1, 0, 0), (1, 15, 13, 128, 256, 1, 1, 1, 1, 1), (2, 16, 16, 64, 64, 3, 1, 1, 1, 0),
    (1, 15, 13, 64, 128, 3, 1, 1, 1, 1), (2, 16, 16, 128, 128, 3, 2, 1, 1, 0), (1, 15, 13, 64, 64, 3, 2, 1, 0, 0),
    (1, 20, 20, 64, 64, 3, 1, 2, 1, 0), (1, 20, 20, 64, 64, 3, 1, 4, 1, 0), (2, 16, 16, 256, 512, 1, 2, 1, 0, 0),
    (1, 60, 107, 128, 128, 3, 1, 1, 1, 0), (8, 64, 64, 64, 256, 1, 1, 1, 1, 1), (8, 7, 7, 512, 512, 3, 1, 1, 1, 0),
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_bn_act_against_fp64_and_simt(case):
    import torch.nn.functional as F
    from vfs_b200 import ops
    N, H, W, Cin, Cout, k, stride, dil, relu, use_res = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k)**0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1
    xs, wp = ops.to_split(x.cuda()), ops.pack_conv_weight(w.cuda())
    Ho, Wo = ops.conv_out_hw(H, W, k, stride, dil)
    rs = ops.to_split(torch.randn(N, Cout, Ho, Wo, generator=g).cuda()) if use_res else None
    out_split, out32 = ops.conv_bn_act(xs, wp, scale.cuda(), shift.cuda(), k, stride, dil, relu, rs, True, True)
    simt = ops.debug_conv_bn_act_simt(xs, wp, scale.cuda(), shift.cuda(), k, stride, dil, relu, rs)
    xr = ops.from_split(xs).cpu().double()
    wr = (wp[0].float() + wp[1].float()).cpu().double().view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xr, wr, stride=stride, padding=0 if k == 1 else dil, dilation=dil if k == 3 else 1)
    ref = ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    if use_res:
        ref = ref + ops.from_split(rs).cpu().double()
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert rel_err(out32, ref) < 5e-5
    assert rel_err(ops.from_split(out_split).permute(0, 2, 3, 1), ref) < 5e-5
    assert rel_err(out32, simt) < 5e-5


PAIR_CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, dil, relu, residual     (split-only output = TMA epilogue; pair tiles 256 x BN)
    (4, 16, 16, 64, 256, 1, 1, 1, 1, 0),      # flat 1x1, 8 m-tiles, BN 256
    (3, 13, 11, 128, 256, 1, 1, 1, 1, 1),     # odd number of m-tiles (phantom tile), residual
    (2, 24, 20, 64, 128, 3, 1, 1, 1, 0),      # 3x3, BN 128 (64 weight rows per CTA)
    (2, 16, 16, 128, 128, 3, 1, 1, 0, 1),     # BasicBlock conv2: 3x3 + residual, BN 128
    (2, 33, 29, 256, 256, 3, 2, 1, 1, 0),     # stride-2 parity views, ragged edges
    (1, 30, 30, 64, 256, 3, 1, 2, 1, 0),      # dilation 2
    (2, 32, 32, 256, 1024, 1, 1, 1, 1, 1),    # layer3 expand: 4 n-tiles x 8 pairs, residual
    (16, 32, 32, 256, 256, 3, 1, 1, 1, 0),    # bench layer3 conv2: 64 pairs -> several tiles per cluster
    (10, 60, 107, 512, 256, 1, 1, 1, 1, 0),   # 480p DAVIS layer3 conv1: 502 m-tiles, persistent loop
]


@pytest.mark.parametrize('case', PAIR_CONV_CASES)
def test_conv_cta_pair_kernel_matches_single_cta_and_fp64(case):
    """The cta_group::2 form of the conv kernel (clusters of two CTAs, M = 256) against the 1-CTA form (must be
    bit-identical: same products, same accumulation order per output element) and against an fp64 convolution."""
    import torch.nn.functional as F
    from vfs_b200 import ops
    N, H, W, Cin, Cout, k, stride, dil, relu, use_res = case
    g = torch.Generator().manual_seed(sum(case) + 7)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k)**0.5
    scale, shift = (torch.rand(Cout, generator=g) + 0.5).cuda(), (torch.randn(Cout, generator=g) * 0.1).cuda()
    xs, wp = ops.to_split(x.cuda()), ops.pack_conv_weight(w.cuda())
    Ho, Wo = ops.conv_out_hw(H, W, k, stride, dil)
    rs = ops.to_split(torch.randn(N, Cout, Ho, Wo, generator=g).cuda()) if use_res else None
    try:
        ops.conv_set_pair_policy(0)
        single, _ = ops.conv_bn_act(xs, wp, scale, shift, k, stride, dil, relu, rs)
        ops.conv_set_pair_policy(1)
        pair, _ = ops.conv_bn_act(xs, wp, scale, shift, k, stride, dil, relu, rs)
        pair2, _ = ops.conv_bn_act(xs, wp, scale, shift, k, stride, dil, relu, rs)   # back to back (PDL overlap)
    finally:
        ops.conv_set_pair_policy(2, 48)
    torch.cuda.synchronize()
    assert torch.equal(pair, single)
    assert torch.equal(pair2, single)
    if N * Ho * Wo * Cout * Cin * k * k <= 2**32:
        xr = ops.from_split(xs).cpu().double()
        wr = (wp[0].float() + wp[1].float()).cpu().double().view(Cout, k, k, Cin).permute(0, 3, 1, 2)
        ref = F.conv2d(xr, wr, stride=stride, padding=0 if k == 1 else dil, dilation=dil if k == 3 else 1)
        ref = ref * scale.cpu().double().view(1, -1, 1, 1) + shift.cpu().double().view(1, -1, 1, 1)
        if use_res:
            ref = ref + ops.from_split(rs).cpu().double()
        if relu:
            ref = torch.relu(ref)
        assert rel_err(ops.from_split(pair), ref) < 5e-5
    assert ops.overflow_count() == 0


DGRAD_CASES = [
    # N, H, W, Cin, Cout, k, stride, dil, with_add
    (2, 16, 16, 64, 128, 1, 1, 1, 0), (2, 16, 16, 64, 64, 3, 1, 1, 1), (1, 15, 13, 128, 64, 3, 1, 1, 0),
    (2, 16, 16, 64, 128, 3, 2, 1, 1), (1, 15, 13, 64, 64, 3, 2, 1, 0), (2, 16, 16, 128, 256, 1, 2, 1, 1),
    (1, 15, 13, 64, 128, 1, 2, 1, 0), (1, 20, 20, 64, 64, 3, 1, 2, 0), (4, 8, 8, 256, 256, 3, 1, 1, 1),
]


@pytest.mark.parametrize('case', DGRAD_CASES)
def test_conv_dgrad_matches_autograd(case):
    """Data gradient of the conv (training backward) against torch autograd in fp64 on the CPU."""
    import torch.nn.functional as F
    from vfs_b200 import ops
    N, H, W, Cin, Cout, k, stride, dil, with_add = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k)**0.5
    Ho, Wo = ops.conv_out_hw(H, W, k, stride, dil)
    dz = torch.randn(N, Cout, Ho, Wo, generator=g)
    add = torch.randn(N, Cin, H, W, generator=g) if with_add else None
    dzs = ops.to_split(dz.cuda())
    wt = ops.pack_conv_weight_dgrad(w.cuda())
    adds = ops.to_split(add.cuda()) if with_add else None
    dx = ops.from_split(ops.conv_dgrad(dzs, wt, (H, W), k, stride, dil, adds))
    x = torch.zeros(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
    z = F.conv2d(x, w.double(), stride=stride, padding=0 if k == 1 else dil, dilation=dil if k == 3 else 1)
    (ref, ) = torch.autograd.grad(z, x, ops.from_split(dzs).cpu().double())
    if with_add:
        ref = ref + ops.from_split(adds).cpu().double()
    assert rel_err(dx, ref) < 5e-5


WGRAD_CASES = [
    # N, H, W, Cin, Cout, k, stride, dil
    (2, 16, 16, 64, 128, 1, 1, 1), (2, 16, 16, 128, 128, 3, 1, 1), (1, 15, 13, 64, 64, 3, 1, 1),
    (2, 16, 16, 64, 128, 3, 2, 1), (1, 15, 13, 128, 256, 3, 2, 1), (2, 16, 16, 256, 512, 1, 2, 1),
    (1, 20, 20, 64, 128, 3, 1, 2), (8, 32, 32, 256, 256, 3, 1, 1), (4, 8, 8, 512, 128, 1, 1, 1),
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_conv_wgrad_matches_autograd(case):
    """Weight gradient of the conv (tcgen05, MN-major operands) against torch autograd in fp64 on the CPU."""
    import torch.nn.functional as F
    from vfs_b200 import ops
    N, H, W, Cin, Cout, k, stride, dil = case
    g = torch.Generator().manual_seed(sum(case) + 2)
    x = torch.randn(N, Cin, H, W, generator=g)
    Ho, Wo = ops.conv_out_hw(H, W, k, stride, dil)
    dz = torch.randn(N, Cout, Ho, Wo, generator=g)
    xs, dzs = ops.to_split(x.cuda()), ops.to_split(dz.cuda())
    dw = ops.conv_wgrad(xs, dzs, k, stride, dil)
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    z = F.conv2d(ops.from_split(xs).cpu().double(), w, stride=stride, padding=0 if k == 1 else dil,
                 dilation=dil if k == 3 else 1)
    (ref, ) = torch.autograd.grad(z, w, ops.from_split(dzs).cpu().double())
    assert rel_err(dw, ref) < 5e-5
    # accumulate into an existing gradient (second view of the SimSiam step)
    dw2 = ops.conv_wgrad(xs, dzs, k, stride, dil, out=dw.clone(), accumulate=True)
    assert rel_err(dw2, 2 * ref) < 5e-5


def test_layout_roundtrip_and_stem():
    import torch.nn.functional as F
    from vfs_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 72, 9, 11, generator=g)
    assert rel_err(ops.from_split(ops.to_split(x.cuda())), x) < 1e-5
    img = torch.randn(2, 3, 67, 93, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    out = ops.from_split(ops.stem_forward(img.cuda(), w.cuda(), scale.cuda(), shift.cuda()))
    ref = F.conv2d(img.double(), w.double(), stride=2, padding=3)
    ref = F.max_pool2d(torch.relu(ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)), 3, 2, 1)
    assert rel_err(out, ref) < 2e-5


# --------------------------------------------------------------------------------------------- head / loss
@pytest.mark.parametrize('name', sorted(cases.HEAD_CASES))
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_head_matches_reference_golden(golden, name, mode):
    from vfs_b200.heads import SimSiamHead
    c = cases.HEAD_CASES[name]
    head = SimSiamHead(**c['cfg'])
    sd = oracle.seeded_state_dict(head, seed=c['seed'])
    head = _load(head, sd)
    head.train(mode == 'train')
    x1, x2 = cases.head_inputs(c)
    z1, p1 = head(x1.cuda())
    z2, p2 = head(x2.cuda())
    loss = head.loss(p1, z1, p2, z2)['loss_feat']
    assert rel_err(z1, golden[f'head/{name}/{mode}/z1']) < REL_TOL
    assert rel_err(p1, golden[f'head/{name}/{mode}/p1']) < REL_TOL
    np.testing.assert_allclose(loss.detach().cpu().numpy(), golden[f'head/{name}/{mode}/loss'], rtol=REL_TOL,
                               atol=1e-5)
    if mode == 'train':  # running statistics must have been updated like torch's BatchNorm1d does
        with torch.no_grad():
            x = torch.nn.functional.adaptive_avg_pool2d(x1, 1).flatten(1)
            y = torch.nn.functional.linear(x, sd['projection_fcs.0.weight'], sd['projection_fcs.0.bias'])
            exp_mean = 0.9 * sd['projection_fcs.1.running_mean'] + 0.1 * y.mean(0)
        # two forward calls happened (x1 then x2); check the first update is contained by recomputing both
        y2 = torch.nn.functional.linear(torch.nn.functional.adaptive_avg_pool2d(x2, 1).flatten(1),
                                        sd['projection_fcs.0.weight'], sd['projection_fcs.0.bias'])
        exp_mean = 0.9 * exp_mean + 0.1 * y2.mean(0)
        assert rel_err(head.projection_fcs[1].running_mean, exp_mean) < REL_TOL
        assert int(head.projection_fcs[1].num_batches_tracked) == 2


def test_cosine_loss_matches_reference_golden(golden):
    from vfs_b200.losses import CosineSimLoss
    p, z = cases.loss_inputs()
    for neg in (False, True):
        got = CosineSimLoss(negative=neg)(p.cuda(), z.cuda())
        np.testing.assert_allclose(got.detach().cpu().numpy(), golden[f'loss/cosine/neg{int(neg)}'], rtol=REL_TOL, atol=1e-6)


# --------------------------------------------------------------------------------------------- attention
# A disagreement with the fp64 top-k set is only accepted when every disputed key's fp64 affinity lies within
# TIE_TOL * max(1, |k-th affinity|) of the selection boundary: 5e-6 relative is ~3 fp32 ulps of a 1024-term dot product
# (the reference's own fp32 einsum cannot resolve it either).  TIE_EXCUSES counts how often that was needed.
TIE_TOL = 5e-6
TIE_EXCUSES = [0]


def _check_topk_modulo_ties(q, k, c, mask_dense, tv, ti, topk, non_mask_len):
    """Index parity: the selected key set must equal the fp64 top-k set, except where the k-th / (k+1)-th fp64
    affinities are closer than the fp32 noise floor (a tie for any fp32 implementation, including the reference)."""
    import torch.nn.functional as F
    qn = F.normalize(q.double(), dim=1).reshape(q.shape[1], -1)
    kn = F.normalize(k.double(), dim=1).reshape(k.shape[1], -1)
    aff = (kn.t() @ qn) / c['temperature']                                   # [T*HW, HW]
    if mask_dense is not None:
        T = k.shape[2]
        HW = qn.shape[1]
        full = mask_dense.reshape(1, HW, HW).expand(T, -1, -1).clone()
        full[:non_mask_len] = True
        aff = aff.masked_fill(~full.reshape(T * HW, HW), float('-inf'))
    srt, order = aff.sort(dim=0, descending=True)
    ti = ti.cpu().long()
    n_q = aff.shape[1]
    bad = 0
    for qi in range(n_q):
        exact = set(order[:topk, qi].tolist())
        got = set(ti[:, qi].tolist())
        if exact == got:
            continue
        # every disagreement must be explained by a near-tie at the selection boundary
        kth = srt[topk - 1, qi]
        diff = (exact ^ got)
        vals = aff[list(diff), qi]
        if not bool(((vals - kth).abs() <= TIE_TOL * max(1.0, float(kth.abs()))).all()):
            bad += 1
        elif not bool((vals == kth).all()):
            # (bit-equal fp64 affinities -- the same frame twice in the key set -- are exact ties: either copy is a
            # correct answer and costs no excuse)
            TIE_EXCUSES[0] += 1
    return bad, n_q


@pytest.mark.parametrize('name', sorted(cases.ATTENTION_CASES))
def test_attention_matches_reference_golden(golden, name):
    from vfs_b200.common import masked_attention_efficient, spatial_neighbor
    from vfs_b200 import ops
    c = cases.ATTENTION_CASES[name]
    q, k, v = cases.attention_inputs(c)
    # C must be a multiple of 64 for the tensor-core path: zero-pad channels (does not change dot products/norms)
    pad = (-c['C']) % 64
    qp = torch.nn.functional.pad(q, (0, 0, 0, 0, 0, pad))
    kp = torch.nn.functional.pad(k, (0, 0, 0, 0, 0, 0, 0, pad))
    mask = spatial_neighbor(1, c['H'], c['W'], c['range'], mode=c.get('mask_mode', 'circle')) if c['range'] else None
    out = masked_attention_efficient(qp.cuda(), kp.cuda(), v.cuda(), mask, temperature=c['temperature'],
                                     topk=c['topk'], non_mask_len=c.get('non_mask_len', 0),
                                     mode=c.get('mode', 'softmax'))
    ref = golden[f'attention/{name}/out']
    assert rel_err(out, ref) < REL_TOL
    if mask is not None:  # analytic mask == the reference's materialised one
        dense = mask.dense()
        np.testing.assert_array_equal(np.packbits(dense.numpy()), golden[f'attention/{name}/mask_packed'])
    # index parity (exact modulo near-ties)
    _, tv, ti = ops.masked_attention(qp.cuda(), kp.cuda(), v.cuda(), mask, c['temperature'], c['topk'], True,
                                     c.get('non_mask_len', 0), c.get('mode', 'softmax'), return_topk=True)
    dense = None
    if mask is not None:
        dense = mask.dense()
        dense = dense[0] if dense.ndim == 3 else dense
    TIE_EXCUSES[0] = 0
    bad, n_q = _check_topk_modulo_ties(q, k, c, dense, tv[0], ti[0], c['topk'], c.get('non_mask_len', 0))
    print(f'[top-k parity] {name}: {TIE_EXCUSES[0]}/{n_q} queries needed the near-tie window ({TIE_TOL:g} rel)')
    assert bad == 0, f'{bad}/{n_q} queries selected a different key set outside the near-tie tolerance'
    assert TIE_EXCUSES[0] <= max(2, n_q // 200), f'{TIE_EXCUSES[0]}/{n_q} queries hid behind the near-tie window'
    # selected affinities are sorted and finite
    assert bool((tv[0][:-1] >= tv[0][1:]).all())


GENERIC_ATTN_CASES = [
    # name, N, C, Cv, T, (Hq, Wq), (Hk, Wk), mask kind, topk, mode, non_mask_len
    ('bool2d_topk', 1, 64, 3, 2, (9, 11), (9, 11), 'random2d', 5, 'softmax', 0),
    ('bool2d_first_free', 1, 64, 4, 3, (8, 8), (8, 8), 'random2d', 10, 'softmax', 1),
    ('window_dense_softmax', 1, 32, 3, 2, (9, 11), (9, 11), 'window', None, 'softmax', 0),
    ('nomask_dense_softmax', 2, 64, 20, 1, (7, 9), (7, 9), None, None, 'softmax', 0),
    ('bool3d_topk', 2, 64, 3, 1, (8, 9), (8, 9), 'random3d', 4, 'softmax', 0),
    ('dense_cosine', 1, 64, 2, 2, (6, 7), (6, 7), 'random2d', None, 'cosine', 0),
    ('rect_query_vs_key', 1, 64, 3, 2, (6, 7), (8, 9), None, 6, 'softmax', 0),
]


@pytest.mark.parametrize('case', GENERIC_ATTN_CASES, ids=[c[0] for c in GENERIC_ATTN_CASES])
def test_attention_general_masks_and_dense_softmax_match_oracle(case):
    """Arbitrary boolean mask tensors, topk=None and query / key maps of different sizes (general kernel,
    csrc/dense.cu) against the oracle restatement of local_attention.py:237-348."""
    from vfs_b200.common import masked_attention_efficient, spatial_neighbor
    _, N, C, Cv, T, (Hq, Wq), (Hk, Wk), kind, topk, mode, nml = case
    g = torch.Generator().manual_seed(len(case[0]) * 17 + N)
    q = torch.relu(torch.randn(N, C, Hq, Wq, generator=g))
    k = torch.relu(torch.randn(N, C, T, Hk, Wk, generator=g))
    v = torch.rand(N, Cv, T, Hk, Wk, generator=g)
    hwk, hwq = Hk * Wk, Hq * Wq
    mask = ref_mask = None
    if kind == 'random2d':
        ref_mask = torch.rand(hwk, hwq, generator=g) > 0.4
        ref_mask[:16] = True                                # every query keeps at least 16 keys
        mask = ref_mask
    elif kind == 'random3d':
        ref_mask = torch.rand(N, hwk, hwq, generator=g) > 0.4
        ref_mask[:, :16] = True
        mask = ref_mask
    elif kind == 'window':
        mask = spatial_neighbor(1, Hk, Wk, 8)
        ref_mask = oracle.spatial_neighbor(Hk, Wk, 8)
    out = masked_attention_efficient(q.cuda(), k.cuda(), v.cuda(), mask.cuda() if torch.is_tensor(mask) else mask,
                                     temperature=0.07 if mode == 'softmax' else 1.0, topk=topk, non_mask_len=nml,
                                     mode=mode)
    temp = 0.07 if mode == 'softmax' else 1.0
    ref = stable_oracle(lambda: oracle.masked_attention_efficient(q, k, v, ref_mask, temperature=temp, topk=topk,
                                                                  non_mask_len=nml, mode=mode))
    assert tuple(out.shape) == tuple(ref.shape) == (N, Cv, Hq, Wq)
    err = rel_err(out, ref)
    if err >= REL_TOL and topk is None:
        # The torch-CPU oracle has given a wrong answer for this very case in ~1 of 10 fresh processes on the GPU boxes
        # (tools/dense_flake_probe.py: the CUDA output matched an fp64 re-evaluation of the same inputs to 1e-7 while
        # two oracle evaluations disagreed by 0.12).  Arbitrate with an independent fp64 evaluation written out below
        # (mask from integer arithmetic): the CUDA path must match it; the oracle's deviation is reported.
        ref64 = _dense_attention_fp64(q, k, v, ref_mask if kind != 'window' else _window_mask_int(Hk, Wk, 8), temp, nml,
                                      mode)
        again = oracle.masked_attention_efficient(q, k, v, ref_mask, temperature=temp, topk=topk, non_mask_len=nml,
                                                  mode=mode)
        print(f'[oracle glitch?] {case[0]}: cuda vs oracle {err:.3e}, cuda vs fp64 {rel_err(out, ref64):.3e}, '
              f'oracle vs fp64 {rel_err(ref, ref64):.3e}, oracle re-evaluated vs fp64 {rel_err(again, ref64):.3e}')
        assert rel_err(ref, ref64) >= REL_TOL, 'the oracle agrees with fp64 but the CUDA path does not'
        err = rel_err(out, ref64)
    assert err < REL_TOL


def _window_mask_int(h, w, neighbor_range):
    """spatial_neighbor (circle) in integer arithmetic: dy^2 + dx^2 < (neighbor_range // 2)^2, bool [HW, HW]."""
    r = neighbor_range // 2
    ys, xs = torch.arange(h).view(h, 1, 1, 1), torch.arange(w).view(1, w, 1, 1)
    d2 = (ys - torch.arange(h).view(1, 1, h, 1))**2 + (xs - torch.arange(w).view(1, 1, 1, w))**2
    return (d2 < r * r).reshape(h * w, h * w)


def _dense_attention_fp64(q, k, v, mask, temperature, non_mask_len, mode):
    """masked_attention_efficient with topk=None written out in float64, one batch item at a time (independent of the
    oracle's chunked einsum formulation): rows = keys (t, y, x), columns = queries."""
    N, C, Hq, Wq = q.shape
    T, Hk, Wk = k.shape[2:]
    qn = torch.nn.functional.normalize(q.double(), p=2, dim=1).reshape(N, C, Hq * Wq)
    kn = torch.nn.functional.normalize(k.double(), p=2, dim=1).reshape(N, C, T * Hk * Wk)
    out = torch.zeros(N, v.shape[1], Hq * Wq, dtype=torch.float64)
    for n in range(N):
        aff = kn[n].t() @ qn[n] / temperature                              # [T*HWk, HWq]
        if mask is not None:
            m = (mask if mask.ndim == 2 else mask[n]).bool()               # [HWk, HWq]
            full = m.repeat(T, 1)
            full[:non_mask_len * Hk * Wk] = True
            aff = aff.masked_fill(~full, float('-inf'))
        wgt = aff.softmax(dim=0) if mode == 'softmax' else aff.clamp(min=0)**2
        out[n] = v[n].double().reshape(v.shape[1], -1) @ wgt
    return out.reshape(N, v.shape[1], Hq, Wq)


ATTN_FORM_CASES = [
    # C, Cv, T, H, W, range, non_mask_len
    (64, 3, 1, 9, 21, 8, 0),        # ragged right edge
    (128, 4, 3, 17, 40, 12, 0),     # odd number of key tiles in some windows (phantom second tile)
    (64, 2, 2, 16, 32, None, 0),    # no mask: the window is the whole map
    (64, 4, 3, 12, 35, 10, 1),      # first key frame exempt from the mask
    (64, 2, 1, 7, 9, 6, 0),         # a single key tile: every step has a phantom partner
    (1024, 4, 1, 32, 32, 36, 0),    # bench shape (256x256 input, stride 8)
    (256, 5, 2, 60, 107, 36, 0),    # 480p DAVIS map
]


@pytest.mark.parametrize('case', ATTN_FORM_CASES)
def test_attention_wide_form_matches_narrow_form(case):
    """Two key tiles per step (N = 256 MMAs) select exactly the keys, values and outputs of one key tile per step."""
    from vfs_b200 import ops
    from vfs_b200.common import spatial_neighbor
    C, Cv, T, H, W, rng, nml = case
    g = torch.Generator().manual_seed(sum(x or 0 for x in case))
    q = torch.relu(torch.randn(1, C, H, W, generator=g)).cuda()
    k = torch.relu(torch.randn(1, C, T, H, W, generator=g)).cuda()
    v = torch.rand(1, Cv, T, H, W, generator=g).cuda()
    mask = spatial_neighbor(1, H, W, rng) if rng else None
    try:
        ops.attention_set_wide(0)
        o0, tv0, ti0 = ops.masked_attention(q, k, v, mask, 0.07, 10, True, nml, 'softmax', return_topk=True)
        ops.attention_set_wide(1)
        o1, tv1, ti1 = ops.masked_attention(q, k, v, mask, 0.07, 10, True, nml, 'softmax', return_topk=True)
    finally:
        ops.attention_set_wide(1)
    torch.cuda.synchronize()
    assert torch.equal(tv0, tv1)
    assert torch.equal(ti0, ti1)
    assert torch.equal(o0, o1)


def test_attention_multi_batch_matches_oracle():
    """N > 1 batch items (the reference API allows it) run as problems of one launch."""
    from vfs_b200.common import masked_attention_efficient, spatial_neighbor
    g = torch.Generator().manual_seed(11)
    N, C, Cv, T, H, W = 3, 64, 3, 2, 11, 19
    q = torch.relu(torch.randn(N, C, H, W, generator=g))
    k = torch.relu(torch.randn(N, C, T, H, W, generator=g))
    v = torch.rand(N, Cv, T, H, W, generator=g)
    out = masked_attention_efficient(q.cuda(), k.cuda(), v.cuda(), spatial_neighbor(1, H, W, 12), temperature=0.07,
                                     topk=10)
    ref = oracle.masked_attention_efficient(q, k, v, oracle.spatial_neighbor(H, W, 12), temperature=0.07, topk=10)
    err = rel_err(out, ref)
    assert err < REL_TOL, (err, int(torch.isnan(out).sum()), [float(rel_err(out[i:i + 1], ref[i:i + 1])) for i in range(N)])


def _fullsize_attention_inputs(T, dup_first):
    H, W, C, Cv = 60, 107, 1024, 4
    g = torch.Generator().manual_seed(1000 + T)
    q = torch.relu(torch.randn(1, C, H, W, generator=g))
    k = torch.relu(torch.randn(1, C, T, H, W, generator=g))
    v = torch.rand(1, Cv, T, H, W, generator=g)
    if dup_first:      # VanillaTracker's key set while frame_idx <= 20: frame 0 twice (vanilla_tracker.py:133-149)
        k[:, :, 1] = k[:, :, 0]
        v[:, :, 1] = v[:, :, 0]
    return q, k, v


def _fp64_topk_parity_on_device(q, k, ti, topk, radius, temperature, chunk=1070):
    """Index parity at sizes whose [T*HW, HW] fp64 affinity (6.9 GB at T = 21) does not fit a host test: the fp64
    affinities are formed on the device, a chunk of queries at a time (torch fp64 matmul as the yardstick).  Returns
    (#queries whose key set differs beyond the near-tie window, #queries that needed the window, #queries)."""
    import torch.nn.functional as F
    _, C, T, H, W = k.shape
    HW = H * W
    dev = torch.device('cuda')
    qn = F.normalize(q.to(dev).double(), dim=1).reshape(C, HW)
    kn = F.normalize(k.to(dev).double(), dim=1).reshape(C, T * HW)
    ky = (torch.arange(T * HW, device=dev) % HW) // W
    kx = (torch.arange(T * HW, device=dev) % HW) % W
    ti = ti.to(dev).long()                                                   # [topk, HW]
    bad = excused = 0
    for q0 in range(0, HW, chunk):
        q1 = min(HW, q0 + chunk)
        aff = (kn.t() @ qn[:, q0:q1]) / temperature                          # [T*HW, n]
        qy = (torch.arange(q0, q1, device=dev) // W)[None]
        qx = (torch.arange(q0, q1, device=dev) % W)[None]
        inside = ((ky[:, None] - qy)**2 + (kx[:, None] - qx)**2).double().sqrt() < radius   # affinity_utils.py:150
        aff = aff.masked_fill(~inside, float('-inf'))
        top = aff.topk(topk, dim=0)
        kth = top.values[topk - 1]                                           # [n]
        got = ti[:, q0:q1]
        got_vals = aff.gather(0, got)
        same = (got.sort(dim=0).values == top.indices.sort(dim=0).values).all(dim=0)
        tol = TIE_TOL * kth.abs().clamp_min(1.0)
        # a differing set is excused iff every selected key and every exact key is within tol of the boundary or in both
        # sets; equivalently: all selected values >= kth - tol (nothing clearly worse was picked)
        ok_vals = (got_vals >= (kth - tol)[None]).all(dim=0) & torch.isfinite(got_vals).all(dim=0)
        bad += int((~same & ~ok_vals).sum())
        excused += int((~same & ok_vals).sum())
        del aff, inside, top
    return bad, excused, HW


@pytest.mark.parametrize('T,dup_first', [(1, False), (21, True)], ids=['T1', 'T21_first_frame_twice'])
def test_attention_full_size_matches_oracle(T, dup_first):
    """BASELINE cfg-3 at its real size (480p: 60x107 map, C = 1024, radius 18, top-k 10, temperature 0.07), one key
    frame and the 21-frame steady state with frame 0 in the key set twice: VALUES against the CPU oracle
    (masked_attention_efficient, local_attention.py:237-348) and INDEX sets against fp64 affinities."""
    from vfs_b200 import ops
    from vfs_b200.common import spatial_neighbor
    q, k, v = _fullsize_attention_inputs(T, dup_first)
    H, W = q.shape[2:]
    mask = spatial_neighbor(1, H, W, 36, mode='circle')
    out, tv, ti = ops.masked_attention(q.cuda(), k.cuda(), v.cuda(), mask, 0.07, 10, True, 0, 'softmax',
                                       return_topk=True)
    assert ops.overflow_count() == 0
    torch.set_num_threads(max(1, min(16, len(__import__('os').sched_getaffinity(0)))))
    with torch.no_grad():
        ref = oracle.masked_attention_efficient(q, k, v, oracle.spatial_neighbor(H, W, 36), temperature=0.07, topk=10)
    bad, excused, n_q = _fp64_topk_parity_on_device(q, k, ti[0], 10, 18, 0.07)
    print(f'[top-k parity] 480p T={T}: {excused}/{n_q} queries needed the near-tie window ({TIE_TOL:g} rel), {bad} bad')
    assert bad == 0
    assert excused <= max(2, n_q // 200)
    # values: a query whose 10th/11th keys are an fp32 tie may legitimately propagate a different value vector, so
    # at most `excused` queries may exceed the tolerance (the reference's own fp32 einsum decides those by rounding)
    err = ((out.cpu() - ref).abs().amax(dim=1) / ref.abs().max()).reshape(-1)          # per query
    assert int((err > REL_TOL).sum()) <= 2 * max(excused, 1) + 2, float(err.max())
    assert float(err.median()) < 1e-5


def test_attention_full_size_properties():
    """480p DAVIS size (HW = 60x107, C = 1024, T = 2): size-independent properties -- propagated one-hot labels
    sum to 1 per query (softmax weights sum to 1), every selected key lies inside the radius, indices are unique
    per query, affinities sorted."""
    from vfs_b200 import ops
    from vfs_b200.common import spatial_neighbor
    H, W, C, T, Cv, r = 60, 107, 1024, 2, 4, 36
    g = torch.Generator().manual_seed(7)
    q = torch.relu(torch.randn(1, C, H, W, generator=g)).cuda()
    k = torch.relu(torch.randn(1, C, T, H, W, generator=g)).cuda()
    lab = torch.randint(0, Cv, (1, T, H, W), generator=g)
    v = torch.nn.functional.one_hot(lab, Cv).permute(0, 4, 1, 2, 3).float().contiguous().cuda()
    mask = spatial_neighbor(1, H, W, r, mode='circle')
    out, tv, ti = ops.masked_attention(q, k, v, mask, 0.07, 10, True, 0, 'softmax', return_topk=True)
    assert torch.isfinite(out).all()
    assert float((out.sum(dim=1) - 1).abs().max()) < 1e-4
    ti = ti[0].long().cpu()
    pos = ti % (H * W)
    ky, kx = pos // W, pos % W
    qy = torch.arange(H * W) // W
    qx = torch.arange(H * W) % W
    d2 = (ky - qy)**2 + (kx - qx)**2
    assert bool((d2 < (r // 2)**2).all())
    srt = ti.sort(dim=0)[0]
    assert bool((srt[1:] != srt[:-1]).all())
    assert bool((tv[0][:-1] >= tv[0][1:]).all())
    # cosine similarities of normalised vectors, divided by the temperature
    assert float(tv.max()) <= 1.0 / 0.07 * (1 + 1e-4)


# --------------------------------------------------------------------------------------------- SiamFC
@pytest.mark.parametrize('name', sorted(cases.XCORR_CASES))
def test_xcorr_matches_reference_golden(golden, name):
    from vfs_b200.siamfc import SiamConvFC, SiamFC
    c = cases.XCORR_CASES[name]
    z, x = cases.xcorr_inputs(c)
    got = SiamFC(out_scale=c['out_scale'])(z.cuda(), x.cuda())
    assert rel_err(got, golden[f'xcorr/{name}/siamfc']) < REL_TOL
    m = SiamConvFC(c['C'], c['C'], out_scale=c['out_scale'])
    m = _load(m, oracle.seeded_state_dict(m, seed=c['seed']))
    got = m(z.cuda(), x.cuda())
    assert rel_err(got, golden[f'xcorr/{name}/siamconvfc']) < REL_TOL


def test_graphed_train_step_equals_eager_steps():
    """vfs_b200.GraphedTrainStep (forward + backward + SGD captured in one CUDA graph) must walk exactly the eager
    trajectory: same losses and same parameters after three steps; building the graph must not train the model."""
    import vfs_b200
    from vfs_b200.optim import build_optimizer
    c = cases.TRACKER_TRAIN_CASES['r18_intra']
    opt_cfg = dict(type='SGD', lr=0.05, momentum=0.9, weight_decay=1e-4)

    def make():
        m = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
        m.load_state_dict(oracle.seeded_state_dict(m, seed=c['seed']))
        m = m.cuda()
        m.train()
        return m, build_optimizer(m, opt_cfg)

    g = torch.Generator().manual_seed(77)
    batches = [torch.randn(4, 2, 3, 2, 64, 64, generator=g).cuda() for _ in range(3)]

    def run_eager(n):
        m, opt = make()
        losses = []
        for b in batches[:n]:
            out = m.train_step(dict(imgs=b), opt)
            opt.zero_grad(set_to_none=True)
            out['loss'].backward()
            opt.step()
            losses.append(out['log_vars']['loss'])
        return m.state_dict(), losses

    # One step: parameters must agree to fp32-atomics noise.  Three steps: tiny-batch BatchNorm amplifies that noise
    # chaotically in the parameters (two eager runs already differ by percents in conv1), so only the losses are held.
    sd_e1, _ = run_eager(1)
    _, losses_e = run_eager(3)
    graphed, opt_g = make()
    before = {k: v.clone() for k, v in graphed.state_dict().items()}
    step = vfs_b200.GraphedTrainStep(graphed, opt_g, dict(imgs=batches[0]))
    for k, v in graphed.state_dict().items():
        assert torch.equal(v, before[k]), f'building the graph changed {k}'
    losses_g = [step(dict(imgs=batches[0]))['log_vars']['loss']]
    torch.cuda.synchronize()
    sd_g1 = {k: v.clone() for k, v in graphed.state_dict().items()}
    for k in sd_e1:
        if sd_e1[k].dtype.is_floating_point:
            assert rel_err(sd_g1[k], sd_e1[k]) < 1e-3, k
        else:
            assert torch.equal(sd_g1[k], sd_e1[k]), k
    losses_g += [step(dict(imgs=b))['log_vars']['loss'] for b in batches[1:]]
    assert losses_g[0] == pytest.approx(losses_e[0], rel=1e-5)
    assert losses_g == pytest.approx(losses_e, rel=2e-3)


# --------------------------------------------------------------------------------------------- data feed
@pytest.mark.parametrize('to_bgr', [False, True])
def test_device_normalize_format_matches_pipeline_oracle(to_bgr):
    """uint8 HWC frames -> fp32 NCTHW on the device == Normalize + FormatShape('NCTHW') of the reference pipeline
    (cv2 arithmetic), bit for bit."""
    import vfs_b200
    rng = np.random.RandomState(3 + int(to_bgr))
    B, T, H, W = 3, 2, 22, 34
    frames = rng.randint(0, 256, (B, T, H, W, 3)).astype(np.uint8)
    cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_bgr=to_bgr)   # configs/*:85-86
    feed = vfs_b200.DeviceNormalizeFormat(**cfg)
    got = feed(torch.from_numpy(frames).pin_memory()).cpu().numpy()
    ref = np.concatenate([oracle.normalize_format_ncthw(frames[b], num_clips=1, **cfg) for b in range(B)], axis=0)
    assert got.shape == ref.shape == (B, 3, T, H, W) and got.dtype == np.float32
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize('name', sorted(cases.TRAIN_PIPELINE_CASES))
def test_device_train_augment_matches_reference_golden(name):
    """RandomResizedCrop -> Resize -> Flip -> Normalize -> FormatShape on the device (one kernel) == the output of the
    unmodified reference pipeline classes for the same seeds, bit for bit (cv2-exact fixed-point bilinear resize)."""
    import os
    import random
    import vfs_b200
    c = cases.TRAIN_PIPELINE_CASES[name]
    with np.load(os.path.join(os.path.dirname(__file__), 'golden', 'train_pipeline_golden.npz')) as z:
        ref = z[name]
    frames = cases.train_pipeline_frames(c)
    aug = vfs_b200.DeviceTrainAugment(scale=c['scale'], area_range=c['area_range'], flip_ratio=c['flip_ratio'],
                                      same_on_clip=c['same_on_clip'], same_across_clip=c['same_across_clip'],
                                      to_bgr=c['to_bgr'], **cases.NORM_CFG)
    np.random.seed(c['seed'])
    random.seed(c['seed'])
    boxes, flips = aug.sample((c['H'], c['W']), len(frames), c['clip_len'])
    got = aug(torch.from_numpy(np.stack(frames)).pin_memory(), boxes, flips, clip_len=c['clip_len']).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_array_equal(got, ref)
    # frames of different sizes in one launch (K400 videos are not uniform): each against the oracle
    big = [np.ascontiguousarray(np.pad(f, ((0, 7 * i), (0, 5 * i), (0, 0)), mode='edge')) for i, f in enumerate(frames)]
    boxes2 = [(1, 2, f.shape[1] - 3, f.shape[0] - 1) for f in big]
    got2 = aug([torch.from_numpy(f) for f in big], boxes2, flips, clip_len=c['clip_len']).cpu().numpy()
    exp2 = oracle.train_augment_ncthw(big, boxes2, flips, c['scale'], to_bgr=c['to_bgr'], num_clips=c['num_clips'],
                                      **cases.NORM_CFG)
    np.testing.assert_array_equal(got2, exp2)


def test_pinned_ring_overlapped_feed_returns_batches_in_order():
    import vfs_b200
    ring = vfs_b200.PinnedRing(slots=2)
    batches = [torch.full((3, 1000), float(i)) for i in range(5)]
    ring.put(batches[0])
    for i in range(5):
        if i + 1 < 5:
            ring.put(batches[i + 1])            # next batch's H2D overlaps this batch's consumer
        dev = ring.get()
        assert dev.is_cuda and float(dev.sum()) == 3000.0 * i
    with pytest.raises(RuntimeError):
        ring.get()


# --------------------------------------------------------------------------------------------- SiamFC tracker
def _smooth_maps(gen, S, R):
    """Response-like maps: a few Gaussian bumps + small noise (a pure-noise map makes the arg-max a coin toss)."""
    ys, xs = torch.meshgrid(torch.arange(R).float(), torch.arange(R).float(), indexing='ij')
    maps = 0.02 * torch.randn(S, R, R, generator=gen)
    for s in range(S):
        for _ in range(3):
            cy, cx = (torch.rand(2, generator=gen) * (R - 1)).tolist()
            amp, sig = float(torch.rand(1, generator=gen)) + 0.3, float(torch.rand(1, generator=gen)) * 2 + 1.0
            maps[s] += amp * torch.exp(-((ys - cy)**2 + (xs - cx)**2) / (2 * sig * sig))
    return maps


@pytest.mark.parametrize('seed', range(8))
def test_siamfc_response_peak_matches_cv2_oracle(seed):
    """Device-side bicubic upsample + penalty + Hann blend + arg-max against the reference's cv2 / numpy sequence
    (siamfc_tracker_base.py:263-291)."""
    from oracle import siamfc as o_siamfc
    from vfs_b200 import ops
    gen = torch.Generator().manual_seed(900 + seed)
    S, R, up = 3, 17, 16
    U = R * up
    maps = _smooth_maps(gen, S, R)
    if seed % 3 == 0:
        maps[1] *= 0.3                       # make a non-centre scale win
    hann = np.outer(np.hanning(U), np.hanning(U))
    hann /= hann.sum()
    got = ops.siamfc_response_peak(maps.cuda(), torch.from_numpy(hann).cuda(), U, 0.9745, 0.176).cpu().tolist()
    sid, loc, blended = o_siamfc.response_peak(maps.numpy().copy(), hann, U, S, 0.9745, 0.176)
    assert got[0] == sid
    if (got[1], got[2]) != loc:              # only an fp32-level near-tie may move the peak
        assert abs(blended[got[1], got[2]] - blended[loc]) <= 1e-6 * abs(blended[loc])
        assert abs(got[1] - loc[0]) <= 1 and abs(got[2] - loc[1]) <= 1


def test_siamfc_tracker_matches_oracle():
    """TrackerSiamFC.init / update (R18, SiamConvFC head, 127 / 255 crops) on a synthetic moving blob against the
    CPU oracle tracker: response maps within 1e-3, boxes within a fraction of a pixel."""
    import vfs_b200  # noqa: F401
    from oracle import siamfc as o_siamfc
    from vfs_b200.siamfc import TrackerSiamFC, build_cfg
    cfg = build_cfg(dict(type='ResNet', depth=18, pretrained=None, norm_cfg=dict(type='BN', requires_grad=True)),
                    exemplar_sz=127, out_scale=1e-3)
    trk = TrackerSiamFC(cfg)
    bsd = oracle.seeded_state_dict(trk.net.backbone, seed=71)
    trk.net.backbone.load_state_dict(bsd)
    hsd = oracle.seeded_state_dict(trk.net.head, seed=72)
    trk.net.head.load_state_dict(hsd)
    trk.net.to('cuda')
    rng = np.random.RandomState(5)
    base = (rng.rand(240, 320, 3) * 60 + 60)
    frames = []
    for f in range(4):
        img = base.copy()
        cy, cx = 120 + 6 * f, 150 + 9 * f
        yy, xx = np.mgrid[0:240, 0:320]
        blob = np.exp(-(((yy - cy) / 14.0)**2 + ((xx - cx) / 20.0)**2))
        img += 150 * blob[..., None] * np.array([1.0, 0.6, 0.2])
        frames.append(np.clip(img, 0, 255).astype(np.uint8))
    box0 = [150 - 30 + 1, 120 - 21 + 1, 60, 42]          # 1-indexed x, y, w, h
    ref = o_siamfc.TrackerOracle({k: cfg[k] for k in cfg}, {k: v.cpu() for k, v in bsd.items()},
                                 {k: v.cpu() for k, v in hsd.items()}, 18)
    trk.init(frames[0], box0)
    ref.init(frames[0], box0)
    assert rel_err(trk.kernel, ref.kernel) < REL_TOL
    for img in frames[1:]:
        r_gpu = trk.responses(img)
        r_ref = ref.responses(img)
        assert tuple(r_gpu.shape) == r_ref.shape == (3, 17, 17)
        assert rel_err(r_gpu, r_ref) < REL_TOL
        b_gpu = trk.update(img)
        b_ref = ref.update(img, responses=r_ref)
        assert np.abs(b_gpu - b_ref).max() < 0.75, (b_gpu, b_ref)    # <= one upsampled-response step (0.5 px)
        # keep the two trackers on the same trajectory so that later frames compare the same crops
        trk.center, trk.target_sz = ref.center.copy(), ref.target_sz.copy()
        trk.z_sz, trk.x_sz = ref.z_sz, ref.x_sz


@pytest.mark.parametrize('name', sorted(cases.SIAMFC_TRACKER_CASES))
def test_siamfc_tracker_matches_reference_golden(name):
    """TrackerSiamFC.init / update against the fixture written by the UNMODIFIED reference tracker
    (siamfc_tracker_base.py:200-319 via oracle/ref_shim.py::load_reference_siamfc_tracker): exemplar kernel and raw
    responses within 1e-3, boxes within one upsampled-response step, for the reference-default 120 px exemplar, the
    127 px of BASELINE cfg-5 and the head without adapters."""
    import os
    import vfs_b200  # noqa: F401
    from vfs_b200.siamfc import TrackerSiamFC, build_cfg
    with np.load(os.path.join(os.path.dirname(__file__), 'golden', 'siamfc_tracker_golden.npz')) as z:
        gold = {k: z[k] for k in z.files}
    c = cases.SIAMFC_TRACKER_CASES[name]
    full = cases.siamfc_tracker_cfg(c)
    backbone = dict(full['model']['backbone'])
    backbone['norm_cfg'] = dict(type='BN', requires_grad=True)
    cfg = build_cfg(backbone, exemplar_sz=c['exemplar_sz'], out_scale=c['out_scale'], extra_conv=c['extra_conv'])
    trk = TrackerSiamFC(cfg)
    trk.net.backbone.load_state_dict(oracle.seeded_state_dict(trk.net.backbone, seed=c['seed']))
    if c['extra_conv']:
        trk.net.head.load_state_dict(oracle.seeded_state_dict(trk.net.head, seed=c['seed'] + 1))
    trk.net.to('cuda')
    frames, box0 = cases.siamfc_tracker_frames()
    trk.init(frames[0], box0)
    assert rel_err(trk.kernel, gold[f'{name}/kernel']) < REL_TOL
    for i, img in enumerate(frames[1:]):
        r_gpu = trk.responses(img)
        assert rel_err(r_gpu, gold[f'{name}/responses'][i]) < REL_TOL
        b_gpu = trk.update(img)
        assert np.abs(b_gpu - gold[f'{name}/boxes'][i]).max() < 0.75, (b_gpu, gold[f'{name}/boxes'][i])
    state = np.concatenate([trk.center, trk.target_sz, [trk.z_sz, trk.x_sz]])
    assert np.abs(state - gold[f'{name}/state']).max() < 0.75


@pytest.mark.parametrize('name', sorted(cases.SIAMFC_TRAIN_CASES))
def test_siamfc_train_step_matches_reference_golden(name):
    """TrackerSiamFC.train_step on the device (frozen tcgen05 backbone, 1x1 adapters + x-corr, fused loss + gradient,
    x-corr backward, tcgen05 weight gradients, Adam / SGD kernels) against two steps of the UNMODIFIED reference class
    (tests/golden/siamfc_train_golden.npz): losses of both steps (the second depends on the first update) and the head
    gradients of the first step."""
    import os
    import vfs_b200  # noqa: F401
    from vfs_b200.siamfc import TrackerSiamFC, build_cfg
    with np.load(os.path.join(os.path.dirname(__file__), 'golden', 'siamfc_train_golden.npz')) as z:
        gold = {k: z[k] for k in z.files}
    c = cases.SIAMFC_TRAIN_CASES[name]
    full = cases.siamfc_train_cfg(c)
    backbone = dict(full['model']['backbone'])
    backbone['norm_cfg'] = dict(type='BN', requires_grad=True)
    cfg = build_cfg(backbone, exemplar_sz=c['exemplar_sz'], out_scale=c['out_scale'], extra_conv=True, loss=c['loss'],
                    optimizer=c['optimizer'], lr_schedule='fixed')
    trk = TrackerSiamFC(cfg)
    trk.net.backbone.load_state_dict(oracle.seeded_state_dict(trk.net.backbone, seed=c['seed']))
    trk.net.head.load_state_dict(oracle.seeded_state_dict(trk.net.head, seed=c['seed'] + 1))
    trk.net.to('cuda')
    losses = []
    for i, batch in enumerate(cases.siamfc_train_batches(c)):
        losses.append(trk.train_step(batch, backward=True))
        if i == 0:
            for k, p in trk.net.head.named_parameters():
                ref = gold[f'{name}/grad/{k}']
                err = float(np.linalg.norm(p.grad.cpu().numpy() - ref) / np.linalg.norm(ref))
                assert err < 2e-3, (k, err)
    np.testing.assert_allclose(losses, gold[f'{name}/losses'], rtol=1e-3)
    # evaluation mode of the same call: no update, same loss definition
    before = {k: p.detach().clone() for k, p in trk.net.head.named_parameters()}
    val = trk.train_step(cases.siamfc_train_batches(c)[0], backward=False)
    assert np.isfinite(val)
    for k, p in trk.net.head.named_parameters():
        assert torch.equal(p, before[k])


# --------------------------------------------------------------------------------------------- trackers
def test_vanilla_tracker_multi_level_equals_single_level_runs():
    """Several backbone out indices / test_cfg.all_blocks (reference vanilla_tracker.py:30-46, 92-93, 196-206): the
    propagation runs once per feature level and the results are stacked [L,T,H,W]; every level must equal the
    single-level tracker configured for it, all_blocks' last block of a stage equals that stage's output."""
    import vfs_b200
    name = sorted(cases.TRACKER_TEST_CASES)[0]
    c = cases.TRACKER_TEST_CASES[name]
    imgs, seg = cases.tracker_test_inputs(c)
    meta = [dict(original_shape=(c['H'], c['W'], 3))]

    def run(out_indices, **extra):
        bb = dict(c['backbone'])
        bb['out_indices'] = out_indices
        tc = dict(c['test_cfg'])
        tc['out_indices'] = out_indices
        tc.update(extra)
        m = vfs_b200.build_model(dict(type='VanillaTracker', backbone=bb), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(tc))
        m.backbone.load_state_dict(oracle.seeded_state_dict(m.backbone, seed=c['seed']))
        m = m.cuda()
        m.eval()
        return np.asarray(m.forward_test(imgs.cuda(), seg.cuda(), meta)[0])

    lvl1, lvl2 = run((1, )), run((2, ))
    both = run((1, 2))
    assert both.shape == (2, ) + lvl1.shape
    np.testing.assert_array_equal(both[0], lvl1)
    np.testing.assert_array_equal(both[1], lvl2)
    blocks = run((2, ), all_blocks=True)
    assert blocks.ndim == 4 and blocks.shape[1:] == lvl2.shape
    np.testing.assert_array_equal(blocks[-1], lvl2)


@pytest.mark.parametrize('name', sorted(cases.TRACKER_TEST_CASES))
def test_vanilla_tracker_matches_reference_golden(golden, name):
    """End-to-end DAVIS-style propagation through build_model + the config dicts; label maps are argmaxes, so the
    criterion is pixel agreement (tiny logit differences may flip isolated boundary pixels)."""
    import vfs_b200
    c = cases.TRACKER_TEST_CASES[name]
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=c['backbone']), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(c['test_cfg']))
    model.backbone.load_state_dict(oracle.seeded_state_dict(model.backbone, seed=c['seed']))
    model = model.cuda()
    model.eval()
    imgs, seg = cases.tracker_test_inputs(c)
    preds = model.forward_test(imgs.cuda(), seg.cuda(), [dict(original_shape=(c['H'], c['W'], 3))])
    got = np.asarray(preds[0]).astype(np.uint8)
    ref = golden[f'tracker_test/{name}/preds']
    assert got.shape == ref.shape
    agree = float((got == ref).mean())
    assert agree > 0.995, f'pixel agreement {agree:.4f}'


def test_vanilla_tracker_batched_videos_equal_single_video_calls():
    """forward_test over B videos at once (extension of the reference, which asserts B == 1) must return exactly
    what B separate calls return -- including videos with fewer objects than the batch's largest label id and a
    feature pass split into several ``batch_step`` chunks."""
    import vfs_b200
    c = cases.TRACKER_TEST_CASES['r18_clip5']
    test_cfg = dict(c['test_cfg'], batch_step=7)                 # 3 videos x 5 frames = 15 frames -> chunks 7+7+1
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=c['backbone']), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(test_cfg))
    model.backbone.load_state_dict(oracle.seeded_state_dict(model.backbone, seed=c['seed']))
    model = model.cuda()
    model.eval()
    g = torch.Generator().manual_seed(77)
    B = 3
    imgs = torch.randn(B, 1, 3, c['T'], c['H'], c['W'], generator=g)
    seg = torch.zeros(B, c['H'], c['W'])
    for b in range(B):
        for o in range(1, 2 + b):                                # video b has b+1 objects (labels 1..b+1)
            y0, x0 = 6 * o + 3 * b, 10 * o + 5 * b
            seg[b, y0:y0 + 20, x0:x0 + 26] = o
    meta = [dict(original_shape=(c['H'], c['W'], 3))]
    singles = [model.forward_test(imgs[b:b + 1].cuda(), seg[b:b + 1].cuda(), meta)[0] for b in range(B)]
    batched = model.forward_test(imgs.cuda(), seg.cuda(), meta * B)
    from_host_labels = model.forward_test(imgs.cuda(), seg.pin_memory(), meta * B)    # class count read on the host
    assert len(batched) == B
    for b in range(B):
        np.testing.assert_array_equal(from_host_labels[b], batched[b])
    for b in range(B):
        assert batched[b].shape == singles[b].shape == (c['T'], c['H'], c['W'])
        assert batched[b].dtype == singles[b].dtype
        np.testing.assert_array_equal(batched[b], singles[b])
        assert set(np.unique(batched[b])) <= set(range(b + 2))


@pytest.mark.parametrize('M,N,K', [(32, 2048, 2048), (8, 512, 2048), (20, 48, 64), (5, 16, 128), (33, 64, 192),
                                   (7, 24, 36)])
def test_linear_forward_backward_against_fp64(M, N, K):
    """Head Linear layers (csrc/linear_mma.cu: warp-MMA 3xTF32 kernels when N % 16 == 0 and K % 64 == 0, csrc/head.cu /
    train.cu SIMT kernels otherwise -- the last shape) against fp64: fp32-grade error (the fp32 matmul of the oracle
    itself is 1e-6..1e-5 from fp64 at K = 2048), with and without accumulation into existing gradients."""
    from vfs_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    y = ops.linear_forward(x.cuda(), W.cuda(), b.cuda()).cpu().double()
    ref = x.double() @ W.double().t() + b.double()
    f32 = (x @ W.t() + b).double()
    tol = max(4.0 * float((f32 - ref).abs().max()), 2e-6 * float(ref.abs().max()))
    assert float((y - ref).abs().max()) <= tol, (float((y - ref).abs().max()), tol)
    dx, dW, db = ops.linear_backward(dy.cuda(), x.cuda(), W.cuda())
    for got, want64, want32 in ((dx, dy.double() @ W.double(), dy @ W), (dW, dy.double().t() @ x.double(), dy.t() @ x),
                                (db, dy.double().sum(0), dy.sum(0))):
        err = float((got.cpu().double() - want64).abs().max())
        tol = max(4.0 * float((want32.double() - want64).abs().max()), 2e-6 * float(want64.abs().max()))
        assert err <= tol, (tuple(got.shape), err, tol)
    # accumulate into existing gradient tensors (flat-gradient sinks), no dx
    dW0, db0 = torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    dWa, dba = dW0.cuda(), db0.cuda()
    none_dx, _, _ = ops.linear_backward(dy.cuda(), x.cuda(), W.cuda(), need_dx=False, dW_out=dWa, db_out=dba)
    assert none_dx is None
    np.testing.assert_allclose(dWa.cpu().numpy(), (dW0 + dW.cpu()).numpy(), rtol=0, atol=1e-6 * float(dW.abs().max()) + 1e-6)
    np.testing.assert_allclose(dba.cpu().numpy(), (db0 + db.cpu()).numpy(), rtol=0, atol=1e-5)


def test_single_gpu_test_pipelined_driver_equals_blocking_calls():
    """vfs_b200.apis.single_gpu_test (reference mmaction/apis/test.py:15-45) over a loader of HOST batches: the
    prefetching device feed + two calls in flight must return exactly what one blocking forward_test call per batch on
    device tensors returns, in loader order -- pinned and pageable sources, uint8 label maps (loader dtype)."""
    import vfs_b200
    from vfs_b200.apis import single_gpu_test
    c = cases.TRACKER_TEST_CASES['r18_clip5']
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=c['backbone']), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(c['test_cfg']))
    model.backbone.load_state_dict(oracle.seeded_state_dict(model.backbone, seed=c['seed']))
    model = model.cuda()
    g = torch.Generator().manual_seed(78)
    meta = [dict(original_shape=(c['H'], c['W'], 3))]
    loader = []
    for i in range(5):
        B = 1 + i % 2
        imgs = torch.randn(B, 1, 3, c['T'], c['H'], c['W'], generator=g)
        seg = torch.zeros(B, c['H'], c['W'], dtype=torch.uint8)
        for b in range(B):
            seg[b, 5 + 3 * i:30 + 3 * i, 8 + 2 * b:40 + 2 * b] = 1 + (i + b) % 3
        if i % 2 == 0:
            imgs, seg = imgs.pin_memory(), seg.pin_memory()
        loader.append(dict(imgs=imgs, ref_seg_map=seg, img_meta=meta * B))
    model.eval()
    expect = []
    for d in loader:
        expect.extend(model.forward_test(d['imgs'].cuda(), d['ref_seg_map'].cuda(), d['img_meta']))
    for depth, coalesce in ((1, None), (2, None), (3, None), (2, 1), (2, 3)):   # None: 8 videos per call at depth > 1
        got = single_gpu_test(model, loader, pipeline_depth=depth, coalesce=coalesce)
        assert len(got) == len(expect) == 7
        for a_, b_ in zip(got, expect):
            assert a_.dtype == np.uint8 and a_.shape == b_.shape
            np.testing.assert_array_equal(a_, b_)
    # an unsupported test_cfg.topk is a clear error, not a ctypes TypeError
    dense = vfs_b200.build_model(dict(type='VanillaTracker', backbone=c['backbone']), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(dict(c['test_cfg'], topk=None))).cuda()
    with pytest.raises(NotImplementedError, match='topk'):
        dense.forward_test(loader[0]['imgs'].cuda(), loader[0]['ref_seg_map'], loader[0]['img_meta'])
    handle = model.forward_test_async(loader[0]['imgs'].cuda(), loader[0]['ref_seg_map'], loader[0]['img_meta'])
    np.testing.assert_array_equal(handle.result()[0], expect[0])
    assert handle.result() is handle.result()


@pytest.mark.parametrize('P', [1, 3])
def test_seg_postprocess_matches_torch(P):
    """Fused bilinear upsample + min-max + arg-max (csrc/post.cu) against the reference's torch sequence
    (vanilla_tracker.py:162-181); one all-zero channel exercises the ``max > 0`` branch."""
    from vfs_b200 import ops
    g = torch.Generator().manual_seed(5 + P)
    Cv, fh, fw, H, W = 5, 9, 13, 70, 101
    logit = torch.rand(P, Cv, fh * fw, generator=g)
    logit[:, 3] = 0
    got = ops.seg_postprocess(logit.cuda() if P > 1 else logit[0].cuda(), fh, fw, (H, W)).cpu()
    up = torch.nn.functional.interpolate(logit.view(P, Cv, fh, fw), size=(H, W), mode='bilinear', align_corners=False)
    mn = up.flatten(2).min(-1)[0].view(P, Cv, 1, 1)
    mx = up.flatten(2).max(-1)[0].view(P, Cv, 1, 1)
    ref = torch.where(mx > 0, (up - mn) / (mx - mn + 1e-12), up).argmax(1).byte()
    got = got.view(P, H, W)
    agree = float((got == ref).float().mean())
    assert agree > 0.999, agree                                   # fp32 interpolation order may flip exact ties


@pytest.mark.parametrize('name', sorted(cases.TRACKER_TRAIN_CASES))
def test_simsiam_forward_eval_mode_matches_oracle(name):
    """SimSiamBaseTracker.forward_train through build_model on the reference's model dicts, BN in eval mode
    (train-mode batch statistics are not native yet) against the oracle composition."""
    import vfs_b200
    c = cases.TRACKER_TRAIN_CASES[name]
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    sd = oracle.seeded_state_dict(model, seed=c['seed'])
    model.load_state_dict(sd)
    model = model.cuda()
    model.eval()
    imgs = cases.tracker_train_input(c)
    losses = model.forward_train(imgs.cuda())
    depth = c['model']['backbone']['depth']
    bsd = {k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')}
    hsd = {k[len('img_head.'):]: v for k, v in sd.items() if k.startswith('img_head.')}
    from vfs_b200.common import images2video, video2images
    clip_len = imgs.size(3)
    with torch.no_grad():
        i1 = video2images(imgs[:, 0].contiguous().reshape(-1, *imgs.shape[2:]))
        i2 = video2images(imgs[:, 1].contiguous().reshape(-1, *imgs.shape[2:]))
        z1, p1 = oracle.simsiam_head_forward(hsd, oracle.resnet_forward(bsd, i1, depth))
        z2, p2 = oracle.simsiam_head_forward(hsd, oracle.resnet_forward(bsd, i2, depth))
        intra = c['train_cfg'].get('intra_video', False)
        w = 1. / clip_len if intra else 1.
        exp = {'img_head.0.loss_feat': oracle.simsiam_loss(p1, z1, p2, z2, weight=w)}
        if intra:
            z2v, p2v = images2video(z2, clip_len), images2video(p2, clip_len)
            for i in range(1, clip_len):
                exp[f'img_head.{i}.loss_feat'] = oracle.simsiam_loss(
                    p1, z1, video2images(p2v.roll(i, dims=2)), video2images(z2v.roll(i, dims=2)), weight=w)
    assert set(losses) == set(exp)
    for k_ in exp:
        np.testing.assert_allclose(losses[k_].detach().cpu().numpy(), exp[k_].numpy(), rtol=REL_TOL, atol=2e-5)
    out = model.train_step(dict(imgs=imgs.cuda()), None)
    assert set(out) == {'loss', 'log_vars', 'num_samples'} and out['num_samples'] == imgs.shape[0]
    assert abs(out['log_vars']['loss'] - float(sum(v.mean() for v in exp.values()))) < 1e-3


@pytest.mark.parametrize('depth', [18, 50])
def test_backward_through_eval_mode_bn_matches_oracle_autograd(depth):
    """norm_eval=True fine-tuning with a frozen prefix (reference resnet.py:593-609, 645-654; the SiamFC / transfer
    setting): every BatchNorm runs on its running statistics, stage 1 and the stem are frozen, the rest trains.
    Gradients of all trainable parameters against torch autograd over the CPU oracle, judged against an fp64 run."""
    from vfs_b200.backbones import ResNet
    net = ResNet(depth, norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, frozen_stages=1,
                 out_indices=(3, ), zero_init_residual=True)
    sd = oracle.seeded_state_dict(net, seed=60 + depth)
    g = torch.Generator().manual_seed(depth)
    x = torch.randn(4, 3, 96, 96, generator=g)
    # Realistic running statistics: a fine-tuned network's BN buffers describe its activations.  (With random buffers
    # the eval-mode network is not normalised at all, activation / gradient magnitudes drift by > 1e5 across the 16
    # blocks of R50 and leave the range in which both fp16 planes of a split gradient are normal numbers.)
    cal = ResNet(depth, norm_cfg=dict(type='BN', requires_grad=True), out_indices=(3, ))
    cal = _load(cal, sd)
    cal.train()
    for m in cal.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = 1.0
    with torch.no_grad():
        cal(x.cuda())
    sd = {k: v.detach().cpu().clone() for k, v in cal.state_dict().items()}
    # zero-init-residual state on some blocks: dgamma must survive gamma == 0 (it cannot be recovered from y).  Zeroed
    # AFTER the calibration pass: a BN calibrated behind a gamma == 0 layer sees a constant input, running_var = 0 and a
    # 316x gradient gain (1 / sqrt(eps)) -- fp64 measures |dL/dz| = 447 there, outside any training regime.
    for k in list(sd):
        if k.endswith(('layer2.1.conv2.bn.weight', 'layer3.0.conv3.bn.weight', 'layer3.1.conv2.bn.weight')):
            sd[k] = torch.zeros_like(sd[k])
    net = _load(net, sd)
    net.train()                                   # norm_eval keeps the BNs in eval mode, frozen stages stay frozen
    wout = torch.randn(4, 2048 if depth == 50 else 512, 3, 3, generator=g) * 1e-2

    def oracle_grads(dtype):
        params = {k: (v.to(dtype).clone() if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
        for k, v in params.items():
            frozen = k.startswith(('conv1.', 'layer1.')) or 'running' in k or not v.dtype.is_floating_point
            v.requires_grad_(not frozen)
        y = oracle.resnet_forward(params, x.to(dtype), depth, out_indices=(3, ), bn_training=False)
        (y * wout.to(dtype)).sum().backward()
        return {k: v.grad for k, v in params.items() if v.requires_grad}

    ref, ref64 = oracle_grads(torch.float32), oracle_grads(torch.float64)
    y = net(x.cuda())
    (y * wout.cuda()).sum().backward()
    from vfs_b200 import ops
    assert ops.overflow_count() == 0
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(ref), set(got) ^ set(ref)
    gnorm = max(float(r.norm()) for r in ref64.values())
    bad = []
    for k, gk in got.items():
        r64 = ref64[k]
        denom = max(float(r64.norm()), 1e-6 * gnorm)
        mine = float((gk.cpu().double() - r64).norm()) / denom
        base = float((ref[k].double() - r64).norm()) / denom
        # Floor 4e-2: the tcgen05 forward is ~7e-5 (max-relative) away from fp64 (fp32 accumulation in the tensor core
        # truncates; the fp32 oracle is at 4e-6), which flips the ReLU mask of the few activations that sit within 1e-4
        # of zero.  On these small maps ONE flipped element moves every upstream gradient by ~1e-2: the error is a step
        # function of depth that starts at the flipped layer with d(beta) hit 20x harder than d(gamma)
        # (tools/evalbn_debug.py) -- not a rounding drift.  A wrong formula (missing gamma or invstd, batch-statistic
        # terms left in) is O(1) and a wrong parameter set is caught above.  Measured over repeated runs: R18 <= 1.2e-2,
        # R50 up to 2.4e-2 (layer4.1.conv2.bn.bias: 4 images x 3 x 3 positions = 36 terms per channel, one flipped
        # term changes that channel's d(beta) by a full summand); the cosine check below bounds the same quantity at
        # 6e-2.
        if mine > max(10 * base, 4e-2):
            bad.append((k, mine, base))
    assert not bad, bad[:6]
    cos = []
    for k, gk in got.items():
        a, b = gk.cpu().double().reshape(-1), ref64[k].reshape(-1)
        if float(b.norm()) > 1e-6 * gnorm:
            cos.append(float(torch.dot(a, b) / (a.norm() * b.norm())))
    assert min(cos) > 0.998, min(cos)
    # frozen parameters received nothing, running statistics did not move
    for k, p in net.named_parameters():
        if k.startswith(('conv1.', 'layer1.')):
            assert p.grad is None
    for k, v in net.state_dict().items():
        if 'running' in k:
            assert torch.equal(v.cpu(), sd[k]), k


@pytest.mark.parametrize('pair_mode', [0, 1], ids=['one_cta', 'cta_pair'])
@pytest.mark.parametrize('shape', [(3, 7, 9, 64, 64, 3, 1), (2, 14, 14, 128, 256, 1, 1), (5, 6, 6, 64, 128, 3, 2),
                                   (4, 8, 8, 256, 512, 1, 1), (9, 5, 5, 64, 256, 3, 1)])
def test_conv_stats_tma_epilogue_matches_legacy_epilogue(shape, pair_mode):
    """Train-mode forward conv through the TMA epilogue (split z + statistics from warp-shuffle column sums) against
    the fp32-output epilogue: same raw conv output (to split precision), same per-channel sum / sum of squares --
    ragged tiles, stride 2, several N tiles, the CTA-pair form with its phantom tile."""
    from vfs_b200 import ops
    N, H, W, Cin, Cout, k, stride = shape
    g = torch.Generator().manual_seed(sum(shape))
    xs = ops.to_split(torch.randn(N, Cin, H, W, generator=g).cuda())
    w = ops.pack_conv_weight((torch.randn(Cout, Cin, k, k, generator=g) * 0.1).cuda())
    try:
        ops.conv_set_pair_policy(pair_mode, 1)
        z_ref, st_ref = ops.conv_stats(xs, w, k, stride, 1)
        z_new, st_new = ops.conv_stats_split(xs, w, k, stride, 1)
    finally:
        ops.conv_set_pair_policy(2, 48)
    zn = (z_new[0].float() + z_new[1].float())
    assert tuple(zn.shape) == tuple(z_ref.shape)
    assert rel_err(zn, z_ref) < 2e-6
    assert float((st_new - st_ref).abs().max() / st_ref.abs().max()) < 1e-6
    # and against the values themselves (fp64 sums of the fp32 output)
    exp = torch.cat([z_ref.double().sum(dim=(0, 1, 2)), (z_ref.double()**2).sum(dim=(0, 1, 2))])
    assert float((st_new - exp).abs().max() / exp.abs().max()) < 1e-5


# --------------------------------------------------------------------------------------------- training step
def _oracle_train_reference(c, sd, imgs):
    """Loss and parameter gradients of SimSiamBaseTracker.forward_train by torch autograd through the oracle
    (CPU fp32; the oracle functions are plain differentiable torch code)."""
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd.items()}
    losses = oracle.simsiam_forward_train(params, imgs, c['model']['backbone']['depth'],
                                          intra_video=c['train_cfg'].get('intra_video', False)).values()
    loss = sum(l.mean() for l in losses)
    loss.backward()
    return float(loss), {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize('name', sorted(cases.TRACKER_TRAIN_CASES))
def test_train_step_gradients_match_oracle_autograd(name):
    """Full SimSiam training step through the public API (train_step -> loss.backward()): every parameter gradient
    against torch autograd over the CPU oracle, then one native SGD update against torch.optim.SGD's rule."""
    import vfs_b200
    from vfs_b200 import ops
    from vfs_b200.optim import SGD
    c = cases.TRACKER_TRAIN_CASES[name]
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    sd = oracle.seeded_state_dict(model, seed=c['seed'])
    model.load_state_dict(sd)
    model = model.cuda()
    model.train()
    # 8 clips: with the golden fixtures' 2-3 clips the train-mode BatchNorms see 3-12 samples per channel and their
    # backward (a projection orthogonal to 2 directions of a 3-dimensional space) amplifies fp32 rounding to O(1)
    shape = (8, ) + tuple(c['shape'][1:])
    imgs = torch.randn(shape, generator=torch.Generator().manual_seed(900 + c['seed']))
    ref_loss, ref_grads = _oracle_train_reference(c, sd, imgs)
    # fp64 run of the same oracle: the yardstick.  ReLU / max-pool make the gradient discontinuous, so single
    # elements flip under ANY rounding change (the fp32 oracle itself is up to 1e-1 away from fp64 on some tensors);
    # a tensor passes if its relative L2 error against fp64 is within 5x the fp32 oracle's own error (floor 3e-3):
    # split-fp16 operands carry 22 significant bits against fp32's 24, i.e. a few times fp32's rounding per op.
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    _, ref64 = _oracle_train_reference(c, sd64, imgs.double())

    opt = SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    opt.zero_grad()
    out = model.train_step(dict(imgs=imgs.cuda()), opt)
    assert abs(out['log_vars']['loss'] - ref_loss) < 1e-3 * max(1.0, abs(ref_loss))
    out['loss'].backward()
    assert ops.overflow_count() == 0
    got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(ref_grads), set(got) ^ set(ref_grads)
    gnorm = max(float(r.norm()) for r in ref64.values())
    failures, ratios = [], []
    # R50 (the configs' model): within 5x the fp32 oracle's own error, median < 3x (measured: median 1.2x, max 3.8x --
    # fp32's own summation error over K up to 4608 dominates).  The tiny R18 case (64x64 input, 2x2 layer4 maps) is
    # dominated by OPERAND rounding instead: split-fp16 carries 2^-22 against fp32's 2^-24, i.e. 4x per op, and the
    # ratio sits at that 4.4x for every tensor independent of the backward scale (tools/grad_scale_probe.py) -> 10x / 6x.
    bound, median_bound = (5.0, 3.0) if name == 'r50' else (10.0, 6.0)
    for k, g in got.items():
        r64 = ref64[k]
        denom = max(float(r64.norm()), 1e-6 * gnorm)   # mathematically-zero gradients (bias before a BN) are noise
        mine = float((g.cpu().double() - r64).norm()) / denom
        base = float((ref_grads[k].double() - r64).norm()) / denom
        ratios.append(mine / max(base, 1e-7))
        if mine > max(bound * base, 3e-3):
            failures.append((k, mine, base))
    ratios = sorted(ratios)
    print(f'[grad parity] {name}: error vs fp64 relative to the fp32 oracle\'s own error: median '
          f'{ratios[len(ratios) // 2]:.2f}x, p90 {ratios[int(len(ratios) * 0.9)]:.2f}x, max {ratios[-1]:.2f}x '
          f'over {len(ratios)} tensors')
    assert not failures, failures[:8]
    assert ratios[len(ratios) // 2] < median_bound, 'median gradient error too far above the fp32 oracle\'s own'

    # one SGD step (first step: momentum buffer = g + wd*p)
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        exp = before[k] - 0.05 * (p.grad + 1e-4 * before[k])
        assert rel_err(p.detach(), exp) < 1e-5, k


# --------------------------------------------------------------------------------------------- dense affinity helpers
@pytest.mark.parametrize('name', sorted(cases.AFFINITY_CASES))
def test_dense_affinity_and_propagate_match_reference_golden(golden, name):
    """compute_affinity / propagate (exported by the reference, common/affinity_utils.py:6-50)."""
    from vfs_b200.common import compute_affinity, propagate
    c = cases.AFFINITY_CASES[name]
    a, b, img = cases.affinity_inputs(c)
    aff = compute_affinity(a.cuda(), b.cuda(), temperature=c['temperature'], softmax_dim=c['softmax_dim'])
    assert rel_err(aff, golden[f'affinity/{name}/aff']) < REL_TOL
    prop = propagate(img.cuda(), aff, topk=c['topk'])
    assert rel_err(prop, golden[f'affinity/{name}/prop']) < REL_TOL


def test_dense_affinity_masked_matches_oracle():
    from vfs_b200.common import compute_affinity, spatial_neighbor
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(2, 48, 9, 11, generator=g), torch.randn(2, 48, 9, 11, generator=g)
    for dim in (1, 2, None):
        got = compute_affinity(a.cuda(), b.cuda(), temperature=0.07, softmax_dim=dim, mask=spatial_neighbor(1, 9, 11, 8))
        ref = oracle.compute_affinity(a, b, temperature=0.07, softmax_dim=dim, mask=oracle.spatial_neighbor(9, 11, 8))
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isfinite(got.cpu()), fin)
        assert rel_err(got.cpu()[fin], ref[fin]) < REL_TOL
