"""Per-call host timeline of the pipelined evaluation driver (vfs_b200.apis.single_gpu_test): time spent enqueuing a
forward_test_async call and waiting for the previous call's predictions."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import vfs_b200  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=bench.BACKBONE_CFG), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(bench.TEST_CFG))
    model.backbone.load_state_dict(seeded_state_dict(model.backbone, seed=0))
    model = model.to(dev).eval()
    model.backbone.engine.check_versions = False
    g = torch.Generator().manual_seed(1234)
    C, F, S = bench.CLIPS, bench.FRAMES, bench.SIZE
    imgs_host = torch.randn(C, 1, 3, F, S, S, generator=g).pin_memory()
    seg_u8 = bench.seg_input(g, torch).expand(C, S, S).contiguous().to(torch.uint8).pin_memory()
    meta = [dict(original_shape=(S, S, 3))] * C
    imgs_dev = imgs_host.to(dev)

    for mode in ('resident', 'h2d_same_stream', 'ring'):
        ring = vfs_b200.PinnedRing(slots=3)
        for keep in (True, False):
            rows, kept, prev = [], [], None
            torch.cuda.synchronize()
            t_start = time.perf_counter()
            if mode == 'ring':
                ring.put(imgs_host)
            for i in range(24):
                t0 = time.perf_counter()
                if mode == 'resident':
                    imgs = imgs_dev
                elif mode == 'h2d_same_stream':
                    imgs = imgs_host.to(dev, non_blocking=True)
                else:
                    imgs = ring.get()
                    ring.put(imgs_host)
                t1 = time.perf_counter()
                h = model.forward_test_async(imgs, seg_u8, meta)
                t2 = time.perf_counter()
                if prev is not None:
                    r = prev.result()
                    if keep:
                        kept.append(r)
                t3 = time.perf_counter()
                prev = h
                rows.append((t1 - t0, t2 - t1, t3 - t2))
            prev.result()
            if mode == 'ring':
                ring.get()
            torch.cuda.synchronize()
            total = time.perf_counter() - t_start
            tail = rows[8:]
            print('%-16s keep=%d  %.3f ms/call | feed %.3f  enqueue %.3f  wait-prev %.3f (ms, mean of last 16)' %
                  (mode, keep, total / 24 * 1e3, *(sum(r[j] for r in tail) / len(tail) * 1e3 for j in range(3))))


if __name__ == '__main__':
    main()
