"""Debug helper: per-block relative error of the train-mode (batch-stat BN) backbone against the oracle."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from tests.golden import cases
from vfs_b200.backbones import ResNet

for name in ('r50_default', 'r50_siamfc', 'r18_default', 'r50_davis'):
    c = cases.BACKBONE_CASES[name]
    x = cases.backbone_input(c)
    nblocks = sum(ResNet.arch_settings[c['depth']][1])
    errs = []
    for bi in range(nblocks):
        net = ResNet(c['depth'], norm_cfg=dict(type='SyncBN', requires_grad=True), strides=c['strides'],
                     dilations=c['dilations'], out_indices=c['out_indices'])
        sd = oracle.seeded_state_dict(net, seed=c['seed'])
        net.load_state_dict(sd)
        net = net.cuda()
        net.train(True)
        y = net.forward_block(x.cuda(), bi).cpu()
        with torch.no_grad():
            ref = oracle.resnet_forward(sd, x, c['depth'], c['strides'], c['dilations'], c['out_indices'],
                                        bn_training=True, block_index=bi)
        errs.append((bi, tuple(ref.shape[1:]), float((y - ref).abs().max() / ref.abs().max()),
                     float(ref.abs().max())))
    print(name)
    for e in errs:
        print('   block %2d %-18s rel %.2e  absmax %.3g' % e)
