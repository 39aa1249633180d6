import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from tests.golden import cases
from vfs_b200.heads import SimSiamHead
from vfs_b200 import ops

for name, B in (('r18_head', 4), ('r50_head', 3), ('r50_head', 8)):
    c = cases.HEAD_CASES[name]
    head = SimSiamHead(**c['cfg'])
    sd = oracle.seeded_state_dict(head, seed=c['seed'])
    head.load_state_dict(sd)
    head = head.cuda(); head.train()
    g = torch.Generator().manual_seed(5)
    x1 = torch.relu(torch.randn(B, c['cfg']['in_channels'], 2, 2, generator=g))
    x2 = torch.relu(torch.randn(B, c['cfg']['in_channels'], 2, 2, generator=g))
    xa, xb = x1.cuda().requires_grad_(True), x2.cuda().requires_grad_(True)
    z1, p1 = head(xa); z2, p2 = head(xb)
    loss = head.loss(p1, z1, p2, z2)['loss_feat'].mean()
    loss.backward()
    params = {k: v.clone().requires_grad_('running' not in k and v.dtype.is_floating_point) for k, v in sd.items()}
    ra, rb = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    oz1, op1 = oracle.simsiam_head_forward(params, ra, bn_training=True)
    oz2, op2 = oracle.simsiam_head_forward(params, rb, bn_training=True)
    ol = oracle.simsiam_loss(op1, oz1, op2, oz2).mean()
    ol.backward()
    print(name, B, 'loss', float(loss), float(ol))
    print('   dx1 rel', float((xa.grad.cpu() - ra.grad).abs().max() / ra.grad.abs().max()))
    for k, p in head.named_parameters():
        r = params[k].grad
        if r is None or p.grad is None: continue
        print('   %-28s rel %.3e refmax %.3e' % (k, float((p.grad.cpu() - r).abs().max() / r.abs().max().clamp_min(1e-30)), float(r.abs().max())))
