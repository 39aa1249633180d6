"""In-kernel timeline of conv_tc_kernel for one layer (vfs_debug_conv_trace): where does a CTA's time go?

    python tools/conv_trace.py l3_expand_256_1024_res [--cta 0]

Prints, for one CTA: kernel span, per-role wait/busy summary, and the first tiles' event list in ns (SM clock
converted with the measured clock rate).  Codes -- TMA: 1 = slot acquired (load issued); MMA: 8 = accumulator stage
acquired, 2 = operands landed, 3 = tile committed; epilogue: 4 = addressing done / waiting for the accumulator,
5 = accumulator ready, 6 = chunk staged (phase A done), 7 = chunk stored (phase B done)."""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.profile_layers import LAYERS  # noqa: E402
from vfs_b200 import _native as nat  # noqa: E402
from vfs_b200 import ops  # noqa: E402

CAP = 2048


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('layer')
    ap.add_argument('--cta', type=int, default=0)
    ap.add_argument('--tiles', type=int, default=3)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    N, H, W, Cin, Cout, k, s, res = LAYERS[a.layer]
    g = torch.Generator(device='cuda').manual_seed(0)
    x = ops.to_split(torch.randn(N, Cin, H, W, device=dev, generator=g))
    w = torch.randn(Cout, Cin, k, k, device=dev, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    wp = ops.pack_conv_weight(w)
    scale = torch.rand(Cout, device=dev, generator=g) + 0.5
    shift = torch.randn(Cout, device=dev, generator=g) * 0.1
    Ho, Wo = ops.conv_out_hw(H, W, k, s, 1)
    r = ops.to_split(torch.randn(N, Cout, Ho, Wo, device=dev, generator=g)) if res else None
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(2):
        ops.conv_bn_act(x, wp, scale, shift, k, s, 1, True, r)
    buf = torch.zeros(148 * 3 * CAP * 2, dtype=torch.int64, device=dev)
    nat.check(nat.lib().vfs_debug_conv_trace(ctypes.c_void_p(buf.data_ptr()), CAP), 'trace on')
    flush.fill_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_bn_act(x, wp, scale, shift, k, s, 1, True, r)
    e1.record()
    torch.cuda.synchronize()
    nat.check(nat.lib().vfs_debug_conv_trace(None, 0), 'trace off')
    ghz = 1.965  # SM clock in GHz (cycles -> ns); B200 boost clock, see bench.py clocks
    t = buf.cpu().view(148, 3, CAP, 2)
    print(f'{a.layer}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us (traced), assuming {ghz:.3f} GHz')
    roles = ('tma', 'mma', 'epi')
    cta = t[a.cta]
    t0 = int(cta[:, 0, 1].min())
    for ri, name in enumerate(roles):
        ev = [(int(c), int(v)) for c, v in cta[ri].tolist() if v != 0]
        if not ev:
            continue
        span = (ev[-1][1] - t0) / ghz
        print(f'-- {name}: {len(ev)} events, last at {span:.0f} ns')
        lim = {'tma': 40, 'mma': 40, 'epi': 16 * a.tiles + 2}[name]
        prev = t0
        line = []
        for c, v in ev[:lim]:
            line.append(f'{c}@{(v - t0) / ghz:.0f}(+{(v - prev) / ghz:.0f})')
            prev = v
        print('   ' + ' '.join(line))
    # per-tile epilogue statistics over all CTAs
    import statistics
    waits, chunkA, chunkB, tiles = [], [], [], []
    for b in range(148):
        ev = [(int(c), int(v)) for c, v in t[b, 2].tolist() if v != 0]
        last7 = None
        for i, (c, v) in enumerate(ev):
            if c == 5 and i > 0:
                waits.append((v - ev[i - 1][1]) / ghz)
            if c == 6:
                chunkA.append((v - ev[i - 1][1]) / ghz)
            if c == 7:
                chunkB.append((v - ev[i - 1][1]) / ghz)
        t4 = [v for c, v in ev if c == 4]
        tiles += [(b_ - a_) / ghz for a_, b_ in zip(t4, t4[1:])]
    for name, arr in (('epilogue wait for accumulator', waits), ('phase A per chunk', chunkA),
                      ('phase B per chunk', chunkB), ('tile period (epilogue)', tiles)):
        if arr:
            print(f'{name:32s} n={len(arr):5d} median {statistics.median(arr):7.0f} ns  mean {statistics.fmean(arr):7.0f} ns')
    spans = []
    for b in range(148):
        vals = t[b, :, :, 1]
        nz = vals[vals != 0]
        if nz.numel():
            spans.append((int(nz.max()) - int(nz.min())) / ghz)
    print(f'CTA span: median {statistics.median(spans):.0f} ns, max {max(spans):.0f} ns')


if __name__ == '__main__':
    main()
