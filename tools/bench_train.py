"""Training-step timing of the SimSiam pre-training path (BASELINE configs[1] / cfg-2 and cfg-4 of SURVEY 8d):
ResNet-50 SimSiamBaseTracker, `imgs [B,2,3,1,S,S]` per GPU, full step = forward (2 backbone passes, 2 head passes, loss)
+ native backward + gradient all-reduce (N > 1) + SGD(lr .05, momentum .9, wd 1e-4), SyncBN in backbone and head.

    python tools/bench_train.py [--clips 8] [--size 256] [--steps 10]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py

Prints one JSON line (rank 0): frame-pairs/s over all ranks, ms/step (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vfs_b200  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402
from vfs_b200 import ops  # noqa: E402
from vfs_b200.optim import allreduce_grads, build_optimizer  # noqa: E402

MODEL = dict(
    type='SimSiamBaseTracker',
    backbone=dict(type='ResNet', pretrained=None, depth=50, out_indices=(3, ),
                  norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False, zero_init_residual=True),
    img_head=dict(type='SimSiamHead', in_channels=2048, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                  projection_mid_channels=2048, projection_out_channels=2048, num_predictor_fcs=2,
                  predictor_mid_channels=512, predictor_out_channels=2048, with_norm=True,
                  loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--clips', type=int, default=8)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--profile', action='store_true', help='print the torch-profiler kernel table of 2 eager steps')
    ap.add_argument('--graph', action='store_true', help='one CUDA graph per step (vfs_b200.GraphedTrainStep)')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    model = vfs_b200.build_model(MODEL, train_cfg=vfs_b200.ConfigDict(dict(intra_video=False)), test_cfg=None)
    model.load_state_dict(seeded_state_dict(model, seed=0))
    model = model.to(dev)
    model.train()
    opt = build_optimizer(model, dict(type='SGD', lr=0.05, momentum=0.9, weight_decay=1e-4))
    params = [p for p in model.parameters() if p.requires_grad]
    g = torch.Generator().manual_seed(1000 + rank)
    imgs = torch.randn(a.clips, 2, 3, 1, a.size, a.size, generator=g).to(dev)

    def step():
        out = model.train_step(dict(imgs=imgs), opt)
        opt.zero_grad(set_to_none=True)
        out['loss'].backward()
        if world > 1:
            allreduce_grads(params, average=True)
        opt.step()
        return out

    if a.graph:
        graphed = vfs_b200.GraphedTrainStep(model, opt, dict(imgs=imgs))

        def step():   # noqa: F811
            return graphed(dict(imgs=imgs), log=False)

    for _ in range(a.warmup):
        out = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ops.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    # algorithmic conv FLOPs of one step: forward of 2*clips frames through all 4 stages, backward ~ 2x forward
    layers = model.backbone.engine.conv_layer_list((2 * a.clips, 3, a.size, a.size), 3)
    fwd = sum(l['flops'] for l in layers) + 2.0 * 2 * a.clips * (a.size // 2)**2 * 64 * 147
    line = dict(workload=f'SimSiam R50 train step (fwd + bwd + allreduce + SGD), {a.clips} clips x 2 views x {a.size}^2 per GPU, SyncBN',
                n_gpus=world, ms_per_step=ms, frame_pairs_per_s=world * a.clips / (ms * 1e-3),
                loss=float(out['loss']), mode='cuda-graph' if a.graph else 'eager', native_launches_per_step=(ops.LAUNCHES[0] - launches0) / a.steps,
                conv_gflop_fwd=fwd / 1e9, conv_tflops_algorithmic_fwd_bwd=3 * fwd / (ms * 1e-3) / 1e12,
                overflow=ops.overflow_count(), timing='CUDA events over the timed steps, max over ranks')
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
