"""CPU tests: the C-ABI library loads and exports every declared symbol, and the host-side mirror of the reference
interface (registries, configs, module/state-dict contract, layout helpers) behaves like the reference."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CONFIGS = '/root/reference/configs'


def test_library_exports_every_declared_symbol():
    from vfs_b200 import _native
    header = open(os.path.join(ROOT, 'include', 'vfs_b200.h')).read()
    declared = set(re.findall(r'\b(vfs_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/vfs_b200.h but not exported'
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    assert lib.vfs_abi_version() == 1


def test_library_reports_errors_without_gpu():
    from vfs_b200 import _native
    lib = _native.lib()
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.vfs_check_device() != 0
    assert len(lib.vfs_last_error_string()) > 0
    with pytest.raises(RuntimeError):
        _native.check(lib.vfs_check_device(), 'check_device')


def test_registry_and_build_from_cfg():
    from vfs_b200.mmcv_lite import Registry, build_from_cfg
    R = Registry('thing')

    @R.register_module()
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    with pytest.raises(KeyError):
        R.register_module()(A)
    R.register_module(force=True)(A)
    a = build_from_cfg(dict(type='A', x=1), R, default_args=dict(y=5, x=9))
    assert (a.x, a.y) == (1, 5)
    assert isinstance(build_from_cfg(dict(type=A, x=3), R), A)
    with pytest.raises(KeyError):
        build_from_cfg(dict(type='B'), R)
    with pytest.raises(KeyError):
        build_from_cfg(dict(x=1), R)
    with pytest.raises(TypeError):
        build_from_cfg([], R)
    assert 'A' in R and 'B' not in R and len(R) == 1


def test_registries_hold_the_reference_type_names():
    import vfs_b200
    assert vfs_b200.BACKBONES.get('ResNet') is not None
    assert vfs_b200.HEADS.get('SimSiamHead') is not None
    assert vfs_b200.LOSSES.get('CosineSimLoss') is not None
    assert vfs_b200.TRACKERS.get('SimSiamBaseTracker') is not None
    assert vfs_b200.TRACKERS.get('VanillaTracker') is not None
    with pytest.raises(KeyError):
        vfs_b200.build_model(dict(type='NoSuchModel'))


def _config_files():
    files = [os.path.join(ROOT, 'tests', 'data', 'sample_simsiam_config.py')]
    if os.path.isdir(REF_CONFIGS):
        files += sorted(os.path.join(REF_CONFIGS, f) for f in os.listdir(REF_CONFIGS) if f.endswith('.py'))
    return files


@pytest.mark.parametrize('path', _config_files())
def test_configs_build_unchanged(path):
    """The reference's configs/*.py (when /root/reference is present) and the local fixture load through Config and
    build through build_model; tools/test.py:129-133's VanillaTracker rebuild works too."""
    import vfs_b200
    cfg = vfs_b200.Config.fromfile(path)
    assert cfg.model.type == 'SimSiamBaseTracker'
    cfg.merge_from_dict({'model.backbone.norm_eval': True, 'optimizer.lr': 0.1})
    assert cfg.model.backbone.norm_eval is True and cfg.optimizer.lr == 0.1
    model = vfs_b200.build_model(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    depth = cfg.model.backbone.depth
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == {18: 12099520, 50: 38210112}[depth]
    assert model.intra_video == cfg.train_cfg.get('intra_video', False)
    assert float(model.iteration) == 0.0
    backbone_cfg = cfg.model.backbone.copy()
    backbone_cfg['out_indices'] = cfg.test_cfg.out_indices
    backbone_cfg['strides'] = cfg.test_cfg.strides
    tracker = vfs_b200.build_model(dict(type='VanillaTracker', backbone=backbone_cfg), train_cfg=None,
                                   test_cfg=cfg.test_cfg)
    assert tracker.stride == 8
    assert tracker.backbone.layer3[0].conv2.conv.stride == (1, 1) if depth == 50 else True


def test_resnet_state_dict_contract_matches_torchvision_mapping():
    """Checkpoint key names are part of the drop-in contract: they must map 1:1 onto torchvision's
    through the convert_to_pretrained.py renaming."""
    import torchvision
    from vfs_b200.backbones import ResNet
    for depth, tv in ((18, torchvision.models.resnet18), (50, torchvision.models.resnet50)):
        ours = ResNet(depth).state_dict()
        theirs = {k: v for k, v in tv(weights=None).state_dict().items() if not k.startswith('fc.')}

        def to_tv(k):
            k = re.sub(r'^conv1\.conv\.', 'conv1.', k)
            k = re.sub(r'^conv1\.bn\.', 'bn1.', k)
            k = re.sub(r'\.downsample\.conv\.', '.downsample.0.', k)
            k = re.sub(r'\.downsample\.bn\.', '.downsample.1.', k)
            k = re.sub(r'\.conv(\d)\.conv\.', r'.conv\1.', k)
            k = re.sub(r'\.conv(\d)\.bn\.', r'.bn\1.', k)
            return k

        mapped = {to_tv(k): v.shape for k, v in ours.items()}
        assert mapped == {k: v.shape for k, v in theirs.items()}
        # and the loader accepts a torchvision state dict
        net = ResNet(depth)
        left = net._load_torchvision_checkpoint(dict(tv(weights=None).state_dict()))
        assert left == ['fc.bias', 'fc.weight']


def test_head_state_dict_indices():
    from vfs_b200.heads import SimSiamHead
    keys = set(SimSiamHead(in_channels=64, projection_mid_channels=32, projection_out_channels=32,
                           predictor_mid_channels=16, predictor_out_channels=32).state_dict())
    for i in (0, 1, 3, 4, 6, 7):
        assert f'projection_fcs.{i}.weight' in keys
    for i in (0, 1, 3):
        assert f'predictor_fcs.{i}.weight' in keys
    assert 'projection_fcs.2.weight' not in keys


def test_cpu_tensors_are_rejected_not_silently_computed():
    from vfs_b200.backbones import ResNet
    from vfs_b200.losses import CosineSimLoss
    net = ResNet(18)
    net.train(False)
    with pytest.raises(RuntimeError):
        net(torch.randn(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        CosineSimLoss()(torch.randn(2, 8), torch.randn(2, 8))
    with pytest.raises(RuntimeError):
        net.conv1(torch.randn(1, 3, 64, 64))  # ConvModule is a parameter container


def test_resnet_constructor_errors():
    from vfs_b200.backbones import ResNet
    with pytest.raises(KeyError):
        ResNet(20)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=0)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=5)
    with pytest.raises(AssertionError):
        ResNet(50, strides=(1, ), dilations=(1, 1), num_stages=3)
    with pytest.raises(AssertionError):
        ResNet(18, style='tensorflow')
    with pytest.raises(TypeError):
        ResNet(50, pretrained=0).init_weights()
    net = ResNet(50, zero_init_residual=True)
    net.init_weights()
    assert float(net.layer1[0].conv3.bn.weight.abs().sum()) == 0.0
    assert net.output_stride == 32 and net.feat_dim == 2048
    assert ResNet(18, dilations=(1, 1, 2, 4)).layer4[0].conv1.conv.dilation == (2, 2)   # dilation // 2 (:285)
    assert ResNet(18, dilations=(1, 1, 2, 4)).layer4[1].conv1.conv.dilation == (4, 4)


def test_neighbor_mask_equals_oracle_and_layout_helpers():
    import oracle
    from vfs_b200.common import images2video, spatial_neighbor, video2images
    for (h, w, r, mode) in ((9, 11, 8, 'circle'), (12, 17, 10, 'circle'), (10, 10, 6, 'square'), (60, 107, 36, 'circle')):
        ours = spatial_neighbor(1, h, w, r, mode=mode).dense()
        ref = oracle.spatial_neighbor(h, w, r, mode=mode)
        assert torch.equal(ours, ref)
    m = spatial_neighbor(1, 60, 107, 36).dense()
    assert int(m[30 * 107 + 50].sum()) == 1005                      # interior query: 1005 neighbours
    x = torch.randn(2, 3, 4, 5, 6)
    assert torch.equal(images2video(video2images(x), 4), x)
    assert video2images(x[:, :, :1]).shape == (2, 3, 5, 6)


def test_pil_nearest_matches_pillow():
    from PIL import Image
    from vfs_b200.common import pil_nearest_interpolate
    g = torch.Generator().manual_seed(0)
    for (h, w, oh, ow) in ((480, 854, 60, 107), (64, 96, 8, 12), (37, 53, 10, 7), (480, 910, 60, 114)):
        seg = torch.randint(0, 5, (1, 1, h, w), generator=g).float()
        ours = pil_nearest_interpolate(seg, (oh, ow))[0, 0].numpy()
        pil = np.array(Image.fromarray(seg[0, 0].numpy()).resize((ow, oh), Image.NEAREST))
        np.testing.assert_array_equal(ours, pil)


def _parse_losses_worker(rank, world, port, q):
    import torch.distributed as dist
    from vfs_b200.trackers import BaseTracker
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    losses = {'img_head.0.loss_feat': torch.full((4, ), float(rank + 1)), 'acc': torch.tensor([0.5 * (rank + 1)])}
    loss, log_vars = BaseTracker._parse_losses(losses)
    q.put((rank, float(loss), dict(log_vars)))
    dist.destroy_process_group()


def test_parse_losses_averages_over_ranks_gloo_world2():
    """world_size-2 gloo: logged scalars are the rank average (reference base.py:103-108); the loss tensor used
    for backward stays local."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_parse_losses_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [1.0, 2.0]
    for _, _, lv in res:
        assert abs(lv['img_head.0.loss_feat'] - 1.5) < 1e-6 and abs(lv['loss'] - 1.5) < 1e-6
        assert abs(lv['acc'] - 0.75) < 1e-6


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    from vfs_b200.optim import allreduce_grads
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    allreduce_grads(params, average=True)          # params[2] has no gradient: skipped
    q.put((rank, params[0].grad.clone(), params[1].grad.clone(), params[2].grad))
    dist.destroy_process_group()


def test_allreduce_grads_gloo_world2():
    """world_size-2 gloo: the data-parallel gradient exchange of the training step (one flat bucket, averaged)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for _, g0, g1, g2 in res:
        assert torch.equal(g0, torch.full((3, 4), 1.5))
        assert torch.equal(g1, torch.arange(5, dtype=torch.float32) * 1.5)
        assert g2 is None


def test_sgd_and_autograd_reject_cpu():
    from vfs_b200.optim import SGD, build_optimizer
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        SGD([p], lr=0.1).step()
    with pytest.raises(NotImplementedError):
        SGD([p], lr=0.1, nesterov=True)
    with pytest.raises(KeyError):
        build_optimizer(torch.nn.Linear(2, 2), dict(type='Adam', lr=1e-3))
    opt = build_optimizer(torch.nn.Linear(2, 2), dict(type='SGD', lr=0.05, momentum=0.9, weight_decay=1e-4))
    assert opt.param_groups[0]['momentum'] == 0.9


# --------------------------------------------------------------------------------------------- SiamFC tracker (host)
def test_siamfc_crop_matches_reference_golden():
    """Product crop helper and oracle restatement == the reference's siamfc/ops.py::crop_and_resize (cv2) bit for bit:
    committed fixtures everywhere, plus the live reference where /root/reference exists."""
    import os
    import numpy as np
    from oracle import ref_shim, siamfc as o_siamfc
    from tests.golden import cases
    from vfs_b200.siamfc import image_ops
    img = cases.siamfc_image()
    with np.load(os.path.join(os.path.dirname(__file__), 'golden', 'siamfc_crop_golden.npz')) as z:
        gold = {k: z[k] for k in z.files}
    refops = ref_shim.load_reference_siamfc_ops() if ref_shim.available() else None
    for name, (cy, cx, size, out_size) in cases.SIAMFC_CROP_CASES.items():
        center = np.array([cy, cx], dtype=np.float32)
        mine = image_ops.crop_and_resize(img, center, size, out_size)
        orc = o_siamfc.crop_and_resize(img, center, size, out_size)
        assert mine.dtype == gold[name].dtype and mine.shape == gold[name].shape == (out_size, out_size, 3)
        np.testing.assert_array_equal(mine, gold[name])
        np.testing.assert_array_equal(orc, gold[name])
        if refops is not None:
            live = refops.crop_and_resize(img, center, size, out_size=out_size, border_value=np.mean(img, axis=(0, 1)))
            np.testing.assert_array_equal(live, gold[name])


def test_siamfc_tracker_config_and_registry():
    """default_config_base.py values and the backbone override used by TrackerSiamFC."""
    from vfs_b200.siamfc import DEFAULT_CFG, build_cfg
    cfg = build_cfg(dict(type='ResNet', depth=18, pretrained=None), exemplar_sz=127)
    assert cfg.exemplar_sz == 127 and cfg.instance_sz == 255 and cfg.response_sz * cfg.response_up == 272
    b = cfg.model.backbone
    assert tuple(b.strides) == (1, 2, 1, 1) and tuple(b.dilations) == (1, 1, 2, 4) and b.norm_eval and b.depth == 18
    assert DEFAULT_CFG['exemplar_sz'] == 120          # the reference default, not BASELINE's 127


def test_synthetic_weights_match_oracle_seeding():
    """vfs_b200.synthetic (bench / tools) and oracle.seeded_state_dict (tests) implement the same name-keyed fill."""
    import oracle
    from vfs_b200.backbones import ResNet
    from vfs_b200.synthetic import seeded_state_dict
    net = ResNet(18)
    a, b = seeded_state_dict(net, seed=5), oracle.seeded_state_dict(net, seed=5)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_gpu_only_entry_points_fail_loudly_without_cuda():
    """No silent CPU path: the graphed training step and the SiamFC tracker refuse to run without a CUDA device."""
    import vfs_b200
    from vfs_b200.siamfc import TrackerSiamFC, build_cfg
    with pytest.raises(RuntimeError):
        vfs_b200.GraphedTrainStep(torch.nn.Linear(2, 2), None, dict(imgs=torch.zeros(1)))
    with pytest.raises(RuntimeError):
        TrackerSiamFC(build_cfg(dict(type='ResNet', depth=18, pretrained=None)))


def test_masked_attention_argument_checks_mirror_the_reference():
    """local_attention.py:263-272, 294: mode, batch equality, value/key shapes, non_mask_len range and mask shape are
    asserted before anything is computed (here: before the CUDA-only path is even reached)."""
    from vfs_b200.common import masked_attention_efficient, spatial_neighbor
    q = torch.zeros(1, 64, 4, 5)
    k = torch.zeros(1, 64, 2, 4, 5)
    v = torch.zeros(1, 3, 2, 4, 5)
    with pytest.raises(AssertionError):
        masked_attention_efficient(q, k, v, None, topk=5, mode='dot')
    with pytest.raises(AssertionError):
        masked_attention_efficient(torch.zeros(2, 64, 4, 5), k, v, None, topk=5)
    with pytest.raises(AssertionError):
        masked_attention_efficient(q, k, torch.zeros(1, 3, 1, 4, 5), None, topk=5)
    with pytest.raises(AssertionError):
        masked_attention_efficient(q, k, v, None, topk=5, non_mask_len=2)
    with pytest.raises(AssertionError):
        masked_attention_efficient(q, k, v, torch.ones(19, 20, dtype=torch.bool), topk=5)       # mask must be [HWk, HWq]
    with pytest.raises(AssertionError):
        masked_attention_efficient(q, k, v, spatial_neighbor(1, 4, 5, 4, mode='square'), topk=5)  # 3-D mask needs T == 1
    with pytest.raises(RuntimeError):                                                           # valid call, CPU tensors
        masked_attention_efficient(q, k, v, spatial_neighbor(1, 4, 5, 4), topk=5)


def test_checkpoint_conversion_roundtrips_through_torchvision_names(tmp_path):
    """tools/convert_weights/convert_to_pretrained.py equivalent: tracker checkpoint -> torchvision-style backbone
    checkpoint that torchvision itself accepts and that ``ResNet(pretrained=...)`` loads back unchanged."""
    import torchvision
    import vfs_b200
    from tests.golden import cases
    from vfs_b200.backbones import ResNet
    from vfs_b200.convert import backbone_to_torchvision, convert
    from vfs_b200.synthetic import seeded_state_dict
    c = cases.TRACKER_TRAIN_CASES['r50']
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    model.load_state_dict(seeded_state_dict(model, seed=3))
    tv_sd = backbone_to_torchvision(model.state_dict())
    tv = torchvision.models.resnet50(weights=None)
    missing, unexpected = tv.load_state_dict(tv_sd, strict=False)
    assert sorted(missing) == ['fc.bias', 'fc.weight'] and not unexpected
    src, dst = tmp_path / 'latest.pth', tmp_path / 'backbone_tv.pth'
    torch.save(dict(state_dict=model.state_dict(), meta=dict(epoch=1)), src)
    convert(str(src), str(dst))
    net = ResNet(50, pretrained=str(dst), torchvision_pretrain=True)
    net.init_weights()
    for k, v in net.state_dict().items():
        assert torch.equal(v, model.state_dict()['backbone.' + k]), k
    with pytest.raises(RuntimeError):
        backbone_to_torchvision({'backbone.layer1.0.mystery.weight': torch.zeros(1)})


class _EchoModel(torch.nn.Module):
    """forward(return_loss=False, idx=...) -> one result per sample (stands in for a tracker's forward_test)."""

    def forward(self, return_loss=True, idx=None):
        assert not return_loss
        return [dict(sample=int(i), doubled=int(i) * 2) for i in idx]


def _multi_gpu_test_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.utils.data import DataLoader, Dataset
    from torch.utils.data.distributed import DistributedSampler
    from vfs_b200.apis import multi_gpu_test

    class Videos(Dataset):
        def __len__(self):
            return 7                      # odd: the sampler pads one sample

        def __getitem__(self, i):
            return dict(idx=i)

    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    ds = Videos()
    loader = DataLoader(ds, batch_size=2, sampler=DistributedSampler(ds, world, rank, shuffle=False))
    q.put((rank, multi_gpu_test(_EchoModel(), loader)))
    dist.destroy_process_group()


def test_multi_gpu_test_orders_results_like_the_reference_gloo_world2():
    """mmaction/apis/test.py:47-194 on two gloo ranks: rank 0 gets the 7 results in dataset order, rank 1 gets None."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_multi_gpu_test_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[1] is None
    assert [r['sample'] for r in res[0]] == list(range(7))
    assert all(r['doubled'] == 2 * r['sample'] for r in res[0])


# --------------------------------------------------------------------------------------------- data-parallel plumbing
def test_flat_layout_is_16_byte_aligned_and_dense():
    """dp.FlatTrainState places every parameter at a 4-float (16-byte) aligned offset: the peer-memory all-reduce and
    the vectorised flat SGD kernel rely on it; padding stays below 3 floats per tensor."""
    from vfs_b200.dp import _layout
    params = [torch.zeros(s) for s in [(64, 3, 7, 7), (64, ), (5, ), (2048, 512), (3, ), (1, )]]
    offs, total = _layout(params)
    assert offs[0] == 0 and all(o % 4 == 0 for o in offs) and total % 4 == 0
    for (o, p), nxt in zip(zip(offs, params), offs[1:] + [total]):
        assert 0 <= nxt - (o + p.numel()) < 4


def _cross_rank_sum_worker(rank, world, port, q):
    import torch.distributed as dist
    from vfs_b200 import ops, peer
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    assert peer.active() is None               # no communicator installed: torch.distributed carries the exchange
    stats = torch.arange(6, dtype=torch.float64) * (rank + 1)
    w = ops.cross_rank_sum_(stats)
    q.put((rank, w, stats.clone()))
    dist.destroy_process_group()


def test_syncbn_statistics_exchange_falls_back_to_torch_distributed_gloo_world2():
    """The SyncBN [sum | sum of squares] exchange (ops.cross_rank_sum_): without an installed peer communicator it is a
    plain all-reduce over the default group -- the path the eager multi-rank step and these CPU tests take."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33700 + os.getpid() % 2000
    procs = [ctx.Process(target=_cross_rank_sum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for _, w, stats in res:
        assert w == 2
        assert torch.equal(stats, torch.arange(6, dtype=torch.float64) * 3)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_peer_communicator_and_device_feeds_fail_loudly_without_cuda():
    import vfs_b200
    from vfs_b200 import peer
    with pytest.raises(RuntimeError):
        peer.PeerComm(data_bytes=1024, rank=0, world=1)
    with pytest.raises(RuntimeError):
        vfs_b200.PinnedRing()
    aug = vfs_b200.DeviceTrainAugment(mean=[0, 0, 0], std=[1, 1, 1], device='cpu')
    with pytest.raises(RuntimeError):
        aug(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), [(0, 0, 8, 8)], [False])


def test_evaluation_driver_coalesces_only_batches_of_identical_layout():
    """vfs_b200.apis._coalesced: consecutive loader batches are merged along the batch axis up to the limit, never
    across a change of clip length / resolution / dtype, and a batch that already reaches the limit passes through."""
    import torch
    from vfs_b200.apis import _coalesced

    def mk(b, t=3, tag=0):
        return dict(imgs=torch.full((b, 1, 3, t, 4, 4), float(tag)), ref_seg_map=torch.zeros(b, 4, 4, dtype=torch.uint8),
                    img_meta=[dict(tag=tag)] * b)

    batches = [mk(1, tag=0), mk(2, tag=1), mk(1, tag=2), mk(1, 5, tag=3), mk(8, tag=4), mk(3, tag=5), mk(3, tag=6),
               mk(3, tag=7)]
    out = list(_coalesced(batches, 8))
    assert [(d['imgs'].shape[0], d['imgs'].shape[3]) for d in out] == [(4, 3), (1, 5), (8, 3), (6, 3), (3, 3)]
    assert [m['tag'] for m in out[0]['img_meta']] == [0, 1, 1, 2]
    assert out[0]['imgs'][:, 0, 0, 0, 0, 0].tolist() == [0.0, 1.0, 1.0, 2.0]          # sample order is kept
    assert out[2] is batches[4]
    assert [d['imgs'].shape[0] for d in _coalesced(batches[:3], 1)] == [1, 2, 1]
    odd = dict(imgs=torch.zeros(1, 1, 3, 3, 4, 4), extra=7)                           # a non-list, non-tensor entry
    assert list(_coalesced([mk(1), odd, mk(1)], 8))[1] is odd


def test_pending_predictions_handle_and_cpu_driver_loop():
    """Host logic of the evaluation driver without a GPU: a finished handle returns its value unchanged (paths that
    cannot be deferred), and single_gpu_test on a CPU model is the reference's plain loop (list results flattened,
    other results appended, no device feed, no merging)."""
    import torch
    from vfs_b200.apis import single_gpu_test
    from vfs_b200.trackers.vanilla_tracker import PendingPredictions

    value = [1, 2, 3]
    handle = PendingPredictions.finished(value)
    assert handle.result() is value and handle.result() is value

    class Echo(torch.nn.Module):
        def forward(self, imgs, return_loss=True, **kw):
            assert return_loss is False and not imgs.is_cuda
            return [int(x) for x in imgs] if imgs.numel() > 1 else {'single': int(imgs)}

    loader = [dict(imgs=torch.tensor([1, 2])), dict(imgs=torch.tensor([3])), dict(imgs=torch.tensor([4, 5, 6]))]
    assert single_gpu_test(Echo(), loader) == [1, 2, {'single': 3}, 4, 5, 6]
