import sys
sys.path.insert(0, '/root/repo')
import torch, oracle
from vfs_b200.backbones import ResNet
from vfs_b200 import ops
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
calib = '--nocal' not in sys.argv
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-2
net = ResNet(depth, norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, frozen_stages=1, out_indices=(3,), zero_init_residual=True)
sd = oracle.seeded_state_dict(net, seed=60 + depth)
g = torch.Generator().manual_seed(depth)
x = torch.randn(4, 3, 96, 96, generator=g)
if calib:
    cal = ResNet(depth, norm_cfg=dict(type='BN', requires_grad=True), out_indices=(3,))
    cal.load_state_dict(sd); cal = cal.cuda(); cal.train()
    for m in cal.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm): m.momentum = 1.0
    with torch.no_grad(): cal(x.cuda())
    sd = {k: v.detach().cpu().clone() for k, v in cal.state_dict().items()}
net.load_state_dict(sd); net = net.cuda(); net.train()
wout = torch.randn(4, 2048 if depth == 50 else 512, 3, 3, generator=g) * scale
def og(dtype):
    params = {k: (v.to(dtype).clone() if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
    for k, v in params.items():
        frozen = k.startswith(('conv1.', 'layer1.')) or 'running' in k or not v.dtype.is_floating_point
        v.requires_grad_(not frozen)
    y = oracle.resnet_forward(params, x.to(dtype), depth, out_indices=(3,), bn_training=False)
    (y * wout.to(dtype)).sum().backward()
    return y.detach(), {k: v.grad for k, v in params.items() if v.requires_grad}
y32, ref = og(torch.float32); y64, ref64 = og(torch.float64)
y = net(x.cuda())
print('fwd rel err', float((y.cpu().double() - y64).abs().max() / y64.abs().max()), 'fp32 oracle', float((y32.double() - y64).abs().max() / y64.abs().max()))
(y * wout.cuda()).sum().backward()
print('overflow', ops.overflow_count())
for k, p in reversed(list(net.named_parameters())):
    if p.grad is None: continue
    r = ref64[k]
    print(f'{float((p.grad.cpu().double()-r).norm()/r.norm().clamp_min(1e-30)):.2e} base {float((ref[k].double()-r).norm()/r.norm().clamp_min(1e-30)):.2e} norm {float(r.norm()):.2e} {k}')
