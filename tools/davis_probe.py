"""DAVIS-style inference at the real size (BASELINE configs[2] / SURVEY cfg-3): one 480 x 854 video of T frames through
VanillaTracker.forward_test (R50 res4 features with the test_cfg strides, 20 preceding frames + first frame as keys,
radius 18, top-k 10), host frames in pinned memory -> uint8 label maps on the host.  Prints frames/s end to end and
the device time of the two phases (feature bank | propagation loop)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import vfs_b200  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    H, W = 480, 854
    dev = torch.device('cuda', 0)
    test_cfg = dict(bench.TEST_CFG, batch_step=10)
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=bench.BACKBONE_CFG), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(test_cfg))
    model.backbone.load_state_dict(seeded_state_dict(model.backbone, seed=0))
    model = model.to(dev).eval()
    model.backbone.engine.check_versions = False
    g = torch.Generator().manual_seed(0)
    imgs = torch.randn(1, 1, 3, T, H, W, generator=g).pin_memory()
    seg = torch.zeros(1, H, W)
    for o in range(1, 4):
        seg[0, 60 * o:60 * o + 150, 120 * o:120 * o + 200] = o
    seg = seg.pin_memory()
    meta = [dict(original_shape=(H, W, 3))]

    def call():
        return model.forward_test(imgs.to(dev, non_blocking=True), seg.to(dev, non_blocking=True), meta)[0]

    call()
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = call()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = min(times)
    # device phases
    x = imgs.to(dev)[0]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    e[0].record()
    model.get_feat_bank(x)
    e[1].record()
    torch.cuda.synchronize()
    feat_ms = e[0].elapsed_time(e[1])
    print(json.dumps(dict(workload=f'DAVIS-style propagation, 1 video x {T} frames x {H}x{W}, R50 res4 (60x107 map), '
                                   f'20+1 key frames, radius 18, top-k 10', frames_per_s_e2e=T / dt,
                          ms_per_video=dt * 1e3, feature_bank_ms=feat_ms, feature_ms_per_frame=feat_ms / T,
                          propagation_ms_per_frame=(dt * 1e3 - feat_ms) / max(T - 1, 1),
                          h2d_mb=imgs.numel() * 4 / 1e6, out_shape=list(out.shape), labels=sorted(set(out.flatten().tolist()))[:6])))


if __name__ == '__main__':
    main()
