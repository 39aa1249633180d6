"""Repeat the multi-batch attention parity case to expose run-to-run differences (debug instrument)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from vfs_b200.common import masked_attention_efficient, spatial_neighbor  # noqa: E402

g = torch.Generator().manual_seed(11)
N, C, Cv, T, H, W = 3, 64, 3, 2, 11, 19
q = torch.relu(torch.randn(N, C, H, W, generator=g))
k = torch.relu(torch.randn(N, C, T, H, W, generator=g))
v = torch.rand(N, Cv, T, H, W, generator=g)
ref = oracle.masked_attention_efficient(q, k, v, oracle.spatial_neighbor(H, W, 12), temperature=0.07, topk=10)
mask = spatial_neighbor(1, H, W, 12)
outs = []
junk = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    if it % 3 == 1:   # perturb the caching allocator / leave garbage in freed blocks
        junk = [torch.full((1 << (10 + it % 12), ), float('nan'), device='cuda') for _ in range(4)]
        del junk
    out = masked_attention_efficient(q.cuda(), k.cuda(), v.cuda(), mask, temperature=0.07, topk=10).cpu()
    err = float((out - ref).abs().max() / ref.abs().max())
    outs.append(out)
    same = bool(torch.equal(out, outs[0]))
    if err > 1e-3 or not same:
        bad = (out - ref).abs().flatten().argmax()
        print(it, 'err', err, 'same_as_first', same, 'argmax', int(bad), 'nan', int(torch.isnan(out).sum()))
print('done', len(outs), 'max err first', float((outs[0] - ref).abs().max() / ref.abs().max()))
