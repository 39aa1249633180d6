"""CUDA-graph timing of the restricted-attention kernels (scores + top-k, merge + propagate) in isolation.

    python tools/profile_attention.py            # bench shape (8 problems, 32x32 map, C 1024) and the 480p DAVIS map
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vfs_b200 import ops  # noqa: E402
from vfs_b200.common import spatial_neighbor  # noqa: E402


def timed_graph(fn, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def case(name, P, T, H, W, C, Cv, rng):
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device='cuda').manual_seed(0)
    F = P * (T + 1)
    feats = torch.relu(torch.randn(F, C, H, W, device=dev, generator=gen))
    bank = ops.features_to_split(feats, normalize=True)                       # [2,F,H,W,C]
    vals = torch.rand(F, Cv, H * W, device=dev, generator=gen)
    mask = spatial_neighbor(1, H, W, rng)
    q_ids = [p * (T + 1) + T for p in range(P)]
    ids = [[p * (T + 1) + t for t in range(T)] for p in range(P)]
    us = timed_graph(lambda: ops.attention_bank_batched(bank, q_ids, bank, ids, vals, ids, 0, Cv * H * W, H * W, Cv,
                                                        mask, 0.07, 10))
    hw = H * W
    dense = 2.0 * P * T * hw * hw * C
    print(f'{name:28s} {us:8.1f} us   dense-equivalent {3 * dense / us / 1e6:7.1f} TF/s issued-equiv '
          f'({dense / 1e9:.1f} GFLOP dense)')


def main():
    case('bench 8x(32x32) T=1 C=1024', 8, 1, 32, 32, 1024, 4, 36)
    case('480p 60x107 T=1 C=1024', 1, 1, 60, 107, 1024, 4, 36)
    case('480p 60x107 T=5 C=1024', 1, 5, 60, 107, 1024, 4, 36)
    case('480p 60x107 T=21 C=1024', 1, 21, 60, 107, 1024, 4, 36)


if __name__ == '__main__':
    main()
