"""ORACLE -- test infrastructure only (see oracle/__init__.py).

    SimSiamHead.forward / loss   mmaction/models/heads/sim_siam_head.py:143-174
    CosineSimLoss._forward       mmaction/models/losses/sim_loss.py:42-63
"""
import torch
import torch.nn.functional as F


def _bn1d(x, sd, prefix, training):
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if training:
        rm, rv = rm.clone(), rv.clone()
    return F.batch_norm(x, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], training, 0.1, 1e-5)


def simsiam_head_forward(sd, x, num_projection_fcs=3, num_predictor_fcs=2, bn_training=False, prefix=''):
    """x [B,C,h,w] -> (z, p).  AdaptiveAvgPool2d(1) -> flatten -> projector (Linear, BN1d, ReLU except last) ->
    predictor (Linear, BN1d, ReLU, ..., Linear) with the reference's Sequential indices
    (projection_fcs.{0,1,3,4,6,7}, predictor_fcs.{0,1,3}; sim_siam_head.py:76-111)."""
    x = F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)
    idx = 0
    for i in range(num_projection_fcs):
        last = i == num_projection_fcs - 1
        x = F.linear(x, sd[f'{prefix}projection_fcs.{idx}.weight'], sd[f'{prefix}projection_fcs.{idx}.bias'])
        x = _bn1d(x, sd, f'{prefix}projection_fcs.{idx + 1}', bn_training)
        idx += 2
        if not last:
            x = F.relu(x)
            idx += 1
    z = x
    idx = 0
    p = z
    for i in range(num_predictor_fcs):
        last = i == num_predictor_fcs - 1
        p = F.linear(p, sd[f'{prefix}predictor_fcs.{idx}.weight'], sd[f'{prefix}predictor_fcs.{idx}.bias'])
        idx += 1
        if not last:
            p = F.relu(_bn1d(p, sd, f'{prefix}predictor_fcs.{idx}', bn_training))
            idx += 2
    return z, p


def cosine_sim_loss(cls_score, label, with_norm=True, negative=False, loss_weight=1.0):
    """Per-sample 2 - 2*cos (or -cos), sim_loss.py:42-63 (non-pairwise branch) times loss_weight (base.py:37)."""
    if with_norm:
        cls_score = F.normalize(cls_score, p=2, dim=1)
        label = F.normalize(label, p=2, dim=1)
    prod = torch.sum(cls_score * label, dim=1).view(cls_score.size(0), -1)
    loss = -prod.mean(dim=-1) if negative else 2 - 2 * prod.mean(dim=-1)
    return loss * loss_weight


def simsiam_loss(p1, z1, p2, z2, weight=1., **loss_kw):
    """0.5*L(p1, sg(z2)) + 0.5*L(p2, sg(z1)) (sim_siam_head.py:171-173)."""
    return (cosine_sim_loss(p1, z2.detach(), **loss_kw) * 0.5 + cosine_sim_loss(p2, z1.detach(), **loss_kw) * 0.5) * weight
