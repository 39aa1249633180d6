"""Per-shape timing of the BatchNorm backward / apply kernels of the training step (HBM-bound elementwise passes):
R50 at 224^2, 64 images per view (cfg-4).  Prints achieved GB/s against the algorithmic bytes of each kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vfs_b200 import ops  # noqa: E402

SHAPES = [  # (C, H, W, residual?) of distinct BN layers, R50 @224
    (64, 56, 56, False), (256, 56, 56, True), (128, 56, 56, False), (128, 28, 28, False), (512, 28, 28, True),
    (256, 28, 28, False), (256, 14, 14, False), (1024, 14, 14, True), (512, 14, 14, False), (512, 7, 7, False),
    (2048, 7, 7, True)]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device('cuda')
    tot = dict(reduce=0.0, apply=0.0, fwd_apply=0.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for C, H, W, res in SHAPES:
        M = N * H * W
        z32 = torch.randn(N, H, W, C, device=dev)
        y = ops.bn_apply(z32, torch.ones(C, device=dev), torch.zeros(C, device=dev), None, True)
        z = ops.bn_apply(z32, torch.ones(C, device=dev), torch.zeros(C, device=dev), None, False)   # split, like the engine
        del z32
        dy = ops.bn_apply(torch.randn(N, H, W, C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev),
                          None, False)
        mean = torch.zeros(C, device=dev)
        invstd = torch.ones(C, device=dev)
        bn = torch.nn.BatchNorm2d(C).to(dev)

        def both():
            return ops.bn_backward(dy, y, z, mean, invstd, bn, want_g=res)

        # separate timings through the raw entry points
        from vfs_b200 import _native as nat
        from vfs_b200._native import current_stream, ptr
        sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)

        def red():
            nat.lib().vfs_bn_bwd_reduce(ptr(dy), None, ptr(y), None, None, ptr(z), ptr(mean), ptr(invstd), ptr(sums), M, C,
                                        current_stream())

        dz = torch.empty_like(dy)
        g = torch.empty_like(dy) if res else None
        dg = torch.empty(C, device=dev)
        db = torch.empty(C, device=dev)

        def app():
            nat.lib().vfs_bn_bwd_apply(ptr(dy), None, ptr(y), None, None, ptr(z), ptr(mean), ptr(invstd), ptr(bn.weight),
                                       ptr(sums), float(M), ptr(dz), None, ptr(g), ptr(dg), ptr(db), 0, 1.0, M, C,
                                       current_stream())

        sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)

        def fapp():
            ops.bn_apply(z, sc, sh, y if res else None, True)

        def with_flush(fn):
            def run():
                flush.fill_(0)
                fn()
            return run

        t_flush = timed(with_flush(lambda: None))
        t_red = timed(with_flush(red)) - t_flush
        t_app = timed(with_flush(app)) - t_flush
        t_fa = timed(with_flush(fapp)) - t_flush
        e = M * C
        print(f'C={C:5d} {H:3d}x{W:<3d} res={int(res)} M={M:8d}: reduce {t_red:7.1f} us {10 * e / t_red / 1e3:7.0f} GB/s | '
              f'apply {t_app:7.1f} us {(14 + 4 * res) * e / t_app / 1e3:7.0f} GB/s | fwd apply {t_fa:7.1f} us '
              f'{(8 + 4 * res) * e / t_fa / 1e3:7.0f} GB/s', flush=True)
        tot['reduce'] += t_red
        tot['apply'] += t_app
        tot['fwd_apply'] += t_fa
    print('sum over the 11 distinct shapes (us):', {k: round(v, 1) for k, v in tot.items()})


if __name__ == '__main__':
    main()
