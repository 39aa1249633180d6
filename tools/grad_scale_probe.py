"""Gradient parity of the SimSiam step against the fp64 oracle for several backward scales (autograd.GRAD_SCALE):
split-fp16 gradients lose their lo plane to fp16 subnormals below ~0.1, so too small a scale costs precision and too
large a one overflows 65504.  Prints, per scale: overflow count, median / p90 / max of (error vs fp64) / (fp32 oracle's
own error vs fp64) over all parameter tensors, and the largest |scaled activation gradient| seen."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
import vfs_b200  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.test_gpu_parity import _oracle_train_reference  # noqa: E402
from vfs_b200 import autograd as ag, ops  # noqa: E402


def main():
    for name in sys.argv[1:] or ['r18_intra', 'r50']:
        c = cases.TRACKER_TRAIN_CASES[name]
        model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
        sd = oracle.seeded_state_dict(model, seed=c['seed'])
        shape = (8, ) + tuple(c['shape'][1:])
        imgs = torch.randn(shape, generator=torch.Generator().manual_seed(900 + c['seed']))
        _, ref = _oracle_train_reference(c, sd, imgs)
        sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
        _, ref64 = _oracle_train_reference(c, sd64, imgs.double())
        gnorm = max(float(r.norm()) for r in ref64.values())
        for log2s in (8, 12, 16, 20, 24):
            ag.GRAD_SCALE = float(2**log2s)
            model.load_state_dict(sd)
            model = model.cuda()
            model.train()
            for p in model.parameters():
                p.grad = None
            out = model.train_step(dict(imgs=imgs.cuda()), None)
            out['loss'].backward()
            ovf = ops.overflow_count()
            ratios, worst = [], (0, '')
            for k, p in model.named_parameters():
                if p.grad is None:
                    continue
                r64 = ref64[k]
                denom = max(float(r64.norm()), 1e-6 * gnorm)
                mine = float((p.grad.cpu().double() - r64).norm()) / denom
                base = float((ref[k].double() - r64).norm()) / denom
                ratios.append(mine / max(base, 1e-7))
                if mine > worst[0]:
                    worst = (mine, k)
            ratios.sort()
            print(f'{name} S=2^{log2s}: overflow {ovf}, ratio median {ratios[len(ratios) // 2]:.2f} p90 '
                  f'{ratios[int(len(ratios) * .9)]:.2f} max {ratios[-1]:.2f}; worst rel err {worst[0]:.2e} ({worst[1]})',
                  flush=True)


if __name__ == '__main__':
    main()
