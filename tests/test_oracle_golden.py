"""CPU: pins the oracle (oracle/*.py) against outputs of the reference itself (tests/golden/vfs_golden.npz,
produced by tests/golden/make_golden.py from the unmodified /root/reference modules)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import resnet as o_resnet
from tests.golden import cases


def _module_like_state_dict(builder):
    """state-dict *names and shapes* of the corresponding vfs_b200 module (names are part of the contract)."""
    return builder().state_dict()


@pytest.mark.parametrize('name', sorted(cases.BACKBONE_CASES))
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_backbone_matches_reference(golden, name, mode):
    from vfs_b200.backbones import ResNet
    c = cases.BACKBONE_CASES[name]
    net = ResNet(c['depth'], norm_cfg=dict(type='SyncBN', requires_grad=True), strides=c['strides'],
                 dilations=c['dilations'], out_indices=c['out_indices'])
    sd = oracle.seeded_state_dict(net, seed=c['seed'])
    x = cases.backbone_input(c)
    with torch.no_grad():
        y = o_resnet.resnet_forward(sd, x, c['depth'], c['strides'], c['dilations'], c['out_indices'],
                                    bn_training=(mode == 'train'))
    ref = golden[f'backbone/{name}/{mode}']
    assert tuple(y.shape) == ref.shape
    np.testing.assert_allclose(y.numpy(), ref, rtol=0, atol=0)  # same ATen ops, same order: bit-exact


@pytest.mark.parametrize('name', sorted(cases.HEAD_CASES))
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_head_matches_reference(golden, name, mode):
    from vfs_b200.heads import SimSiamHead
    c = cases.HEAD_CASES[name]
    sd = oracle.seeded_state_dict(SimSiamHead(**c['cfg']), seed=c['seed'])
    x1, x2 = cases.head_inputs(c)
    with torch.no_grad():
        z1, p1 = oracle.simsiam_head_forward(sd, x1, bn_training=(mode == 'train'))
        z2, p2 = oracle.simsiam_head_forward(sd, x2, bn_training=(mode == 'train'))
        loss = oracle.simsiam_loss(p1, z1, p2, z2)
    np.testing.assert_allclose(z1.numpy(), golden[f'head/{name}/{mode}/z1'], rtol=0, atol=0)
    np.testing.assert_allclose(p1.numpy(), golden[f'head/{name}/{mode}/p1'], rtol=0, atol=0)
    np.testing.assert_allclose(loss.numpy(), golden[f'head/{name}/{mode}/loss'], rtol=0, atol=0)


def test_cosine_loss_matches_reference(golden):
    p, z = cases.loss_inputs()
    for neg in (False, True):
        got = oracle.cosine_sim_loss(p, z, negative=neg).numpy()
        np.testing.assert_allclose(got, golden[f'loss/cosine/neg{int(neg)}'], rtol=0, atol=0)


@pytest.mark.parametrize('name', sorted(cases.ATTENTION_CASES))
def test_attention_matches_reference(golden, name):
    c = cases.ATTENTION_CASES[name]
    q, k, v = cases.attention_inputs(c)
    mask = None
    if c['range']:
        mask = oracle.spatial_neighbor(c['H'], c['W'], c['range'], mode=c.get('mask_mode', 'circle'))
        packed = np.packbits(mask.numpy())
        np.testing.assert_array_equal(packed, golden[f'attention/{name}/mask_packed'])
    out = oracle.masked_attention_efficient(q, k, v, mask, temperature=c['temperature'], topk=c['topk'],
                                            non_mask_len=c.get('non_mask_len', 0), mode=c.get('mode', 'softmax'))
    np.testing.assert_allclose(out.numpy(), golden[f'attention/{name}/out'], rtol=0, atol=0)


@pytest.mark.parametrize('name', sorted(cases.AFFINITY_CASES))
def test_affinity_propagate_match_reference(golden, name):
    c = cases.AFFINITY_CASES[name]
    a, b, img = cases.affinity_inputs(c)
    aff = oracle.compute_affinity(a, b, temperature=c['temperature'], softmax_dim=c['softmax_dim'])
    np.testing.assert_allclose(aff.numpy(), golden[f'affinity/{name}/aff'], rtol=0, atol=0)
    prop = oracle.propagate(img, aff, topk=c['topk'])
    np.testing.assert_allclose(prop.numpy(), golden[f'affinity/{name}/prop'], rtol=0, atol=0)


@pytest.mark.parametrize('name', sorted(cases.XCORR_CASES))
def test_xcorr_matches_reference(golden, name):
    from vfs_b200.siamfc import SiamConvFC
    c = cases.XCORR_CASES[name]
    z, x = cases.xcorr_inputs(c)
    np.testing.assert_allclose(oracle.xcorr(z, x, c['out_scale']).numpy(), golden[f'xcorr/{name}/siamfc'],
                               rtol=0, atol=0)
    sd = oracle.seeded_state_dict(SiamConvFC(c['C'], c['C'], out_scale=c['out_scale']), seed=c['seed'])
    with torch.no_grad():
        got = oracle.siam_conv_fc(sd, z, x, c['out_scale']).numpy()
    np.testing.assert_allclose(got, golden[f'xcorr/{name}/siamconvfc'], rtol=0, atol=0)
