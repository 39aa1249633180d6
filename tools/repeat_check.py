"""Run-to-run reproducibility of the single-process SimSiam training step (fresh model each run, same weights and
batch): prints the worst relative L2 difference of any parameter gradient between consecutive runs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vfs_b200  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device('cuda')
    cfg = dict(type='SimSiamBaseTracker',
               backbone=dict(type='ResNet', pretrained=None, depth=18, out_indices=(3, ),
                             norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False,
                             zero_init_residual=True),
               img_head=dict(type='SimSiamHead', in_channels=512, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                             projection_mid_channels=512, projection_out_channels=512, num_predictor_fcs=2,
                             predictor_mid_channels=128, predictor_out_channels=512, with_norm=True,
                             loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg'))
    g = torch.Generator().manual_seed(99)
    full = torch.randn(clips, 2, 3, 1, size, size, generator=g).to(dev)

    def run():
        m = vfs_b200.build_model(cfg, train_cfg=vfs_b200.ConfigDict(dict(intra_video=False)), test_cfg=None)
        m.load_state_dict(seeded_state_dict(m, seed=1))
        m = m.to(dev)
        m.train()
        feats = []
        h = m.backbone.register_forward_hook(lambda mod, i, o: feats.append(o.detach().clone()))
        losses = m(imgs=full)
        h.remove()
        loss = sum(v.mean() for k, v in losses.items() if 'loss' in k)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), feats, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}

    runs = [run() for _ in range(4)]
    for i in range(1, len(runs)):
        l0, f0, g0 = runs[i - 1]
        l1, f1, g1 = runs[i]
        fe = max(float((a - b).abs().max() / a.abs().max()) for a, b in zip(f0, f1))
        rows = sorted(((float((g1[k] - g0[k]).norm()) / (float(g0[k].norm()) + 1e-30), k) for k in g0
                       if 'fcs' not in k or 'bias' not in k), reverse=True)
        print(f'clips {clips} size {size} run {i - 1}->{i}: loss {l0:.8f} {l1:.8f} feat diff {fe:.2e}; worst grads: ' +
              ', '.join(f'{e:.2e} {k}' for e, k in rows[:3]), flush=True)
        if '--all' in sys.argv and rows[0][0] > 1e-3:
            for k in reversed(list(g0)):
                print(f'    {float((g1[k] - g0[k]).norm()) / (float(g0[k].norm()) + 1e-30):.2e} {k}')
            break


if __name__ == '__main__':
    main()
