"""Experiment: the bench's conv segment (R50 residual stages over 16 frames) as ONE chain on one stream versus TWO
independent half-batch chains on two streams inside one CUDA graph (kernel boundaries of one chain covered by the
other chain's kernels).  VFS_CONV_PDL=0 recommended for the two-chain form."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import vfs_b200  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402


def main():
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device('cuda', 0)
    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=bench.BACKBONE_CFG), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(bench.TEST_CFG))
    model.backbone.load_state_dict(seeded_state_dict(model.backbone, seed=0))
    model = model.to(dev).eval()
    eng = model.backbone.engine
    eng.check_versions = False
    g = torch.Generator().manual_seed(1)
    frames = torch.randn(16, 3, 256, 256, generator=g).to(dev)
    stem = eng.stem(frames)                                          # [2,16,64,64,64]
    parts = [stem[:, i * (16 // chains):(i + 1) * (16 // chains)].contiguous() for i in range(chains)]
    streams = [torch.cuda.Stream() for _ in range(chains)]
    outs = [None] * chains

    def run():
        cur = torch.cuda.current_stream()
        for i, s in enumerate(streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs[i] = eng.run_stages(parts[i], 2)
        for s in streams:
            cur.wait_stream(s)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref = torch.cat([o.clone() for o in outs], dim=1)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        run()
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    times = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    times.sort()
    full = eng.run_stages(stem, 2)
    same = torch.equal(torch.cat(outs, dim=1), full)
    print(f'chains={chains} PDL={os.environ.get("VFS_CONV_PDL", "1")}: conv segment median {times[len(times) // 2]:.1f} us '
          f'(min {times[0]:.1f}); equals single-chain result: {same}')


if __name__ == '__main__':
    main()
