"""Debug helper: per-parameter gradient comparison of the native training step against oracle autograd."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import vfs_b200
from tests.golden import cases
from tests.test_gpu_parity import _oracle_train_reference

for name in ('r18_intra', 'r50'):
    c = cases.TRACKER_TRAIN_CASES[name]
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    sd = oracle.seeded_state_dict(model, seed=c['seed'])
    model.load_state_dict(sd)
    model = model.cuda()
    model.train()
    imgs = cases.tracker_train_input(c)
    ref_loss, ref = _oracle_train_reference(c, sd, imgs)
    out = model.train_step(dict(imgs=imgs.cuda()), None)
    out['loss'].backward()
    print(name, 'loss', out['log_vars']['loss'], ref_loss)
    rows = []
    for k, p in model.named_parameters():
        if p.grad is None:
            print('  NO GRAD', k)
            continue
        r = ref[k]
        err = float((p.grad.cpu() - r).abs().max())
        rows.append((err / max(float(r.abs().max()), 1e-30), k, float(r.abs().max()), err))
    rows.sort(reverse=True)
    for row in rows[:25]:
        print('   rel %.3e  %-45s refmax %.3e  err %.3e' % row)
    print('   ... median rel %.3e' % sorted(r[0] for r in rows)[len(rows) // 2])
