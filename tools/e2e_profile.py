"""Where does one VanillaTracker.forward_test call (batch of 8 two-frame 256x256 videos, the bench's e2e unit) spend
its time?
cProfile over 50 calls + a GPU-side view (CUDA events around the call, torch profiler kernel table)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import vfs_b200  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=bench.BACKBONE_CFG), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(bench.TEST_CFG))
    model.backbone.load_state_dict(seeded_state_dict(model.backbone, seed=0))
    model = model.to(dev)
    model.eval()
    model.backbone.engine.check_versions = False
    g = torch.Generator().manual_seed(0)
    imgs_host = torch.randn(bench.CLIPS, 1, 3, bench.FRAMES, bench.SIZE, bench.SIZE, generator=g).pin_memory()
    seg_host = bench.seg_input(g, torch).expand(bench.CLIPS, bench.SIZE, bench.SIZE).contiguous().pin_memory()
    meta = [dict(original_shape=(bench.SIZE, bench.SIZE, 3))]

    def call():
        imgs = imgs_host.to(dev, non_blocking=True)
        seg = seg_host.to(dev, non_blocking=True)
        return model.forward_test(imgs, seg, meta * bench.CLIPS)[0]

    for _ in range(5):
        call()
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    torch.cuda.synchronize()
    print(f'wall per call: {(time.perf_counter() - t0) / n * 1e3:.3f} ms')
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        call()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(28)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(10):
            call()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))


if __name__ == '__main__':
    main()
