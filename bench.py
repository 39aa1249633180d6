#!/usr/bin/env python
"""Bench of the VFS hot path on B200:  frame-pairs/sec (R50 res4 features + affinity/top-k label propagation).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one batch of BASELINE configs[1]-shaped synthetic input per GPU: 8 clips x 2 frames x 256x256
(fp32, randn).  For every clip the pair is processed like DAVIS inference does it (reference
VanillaTracker.forward_test with a 2-frame video, tools/test.py:129-133 model): ResNet-50 with the test_cfg strides
(1,2,1,1) up to res4 (layer3, 1024 ch, stride 8 -> 32x32) for both frames, then restricted attention
(radius 18, top-k 10, temperature 0.07) propagating a 4-channel one-hot label map from frame 0 to frame 1.

  value      device-timed (CUDA events) throughput with inputs resident in HBM; L2 flushed between steps
  e2e        the same work through the reference's evaluation entry point: vfs_b200.apis.single_gpu_test
             (mmaction/apis/test.py:15) running build_model(VanillaTracker) over a loader of pinned HOST batches of 8
             two-frame videos; every step's frames go H2D and its predictions D2H inside the timed region (wall
             clock); the driver keeps two forward_test calls in flight.  `e2e.blocking_driver` collects every call
             before issuing the next, `e2e.single_call` is one blocking forward_test call per step (round 1's e2e),
             `e2e.per_video_calls` is the same batch issued the way the reference must issue it (one video per
             forward_test call, vanilla_tracker.py:56 asserts B == 1); `e2e.from_uint8_frames` feeds uint8 HWC
             frames through vfs_b200.DeviceNormalizeFormat (Normalize + FormatShape on the device).
             The sub-benches `train_cfg2` / `train` (SimSiam train step, cfg-2 / cfg-4), `affinity_480p` (cfg-3) and
             `siamfc` (cfg-5) carry their own device and e2e numbers.
  roofline   tcgen05 conv kernel: algorithmic conv FLOPs of a step / CUDA-event time of the conv segment
  roofline_affinity  the fused affinity / top-k / propagation kernels the same way (window-restricted FLOPs)
  cpu_baseline  the CPU oracle (restatement of the reference's torch-CPU path) on a bounded sample, all host cores

`--impl reference` times the reference's CPU implementation of the same step (the oracle port: /root/reference is a
Python tree that cannot travel to the GPU box and needs mmcv; oracle/ restates it op for op and is pinned to it
bit-exactly by tests/test_oracle_golden.py) on the host cores.

Multi-GPU: one process per GPU (torchrun), clips are independent -> sharded with no data-path collective
("weak" scaling: every rank processes its own 8 clips); NCCL is only used for the barrier and the max-over-ranks of
the timings.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS, FRAMES, SIZE = 8, 2, 256
CV = 4
TEST_CFG = dict(precede_frames=20, topk=10, temperature=0.07, strides=(1, 2, 1, 1), out_indices=(2, ),
                neighbor_range=36, with_first=True, with_first_neighbor=True, output_dir='eval_results',
                batch_step=CLIPS * FRAMES)   # reference key (vanilla_tracker.py:58): frames per backbone pass
BACKBONE_CFG = dict(type='ResNet', pretrained=None, depth=50, out_indices=(2, ), strides=(1, 2, 1, 1),
                    norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False, zero_init_residual=True)
WORKLOAD = 'r50_res4_feat+affinity_topk10 8clips x 2frames x 256x256 (BASELINE configs[1] shape, DAVIS test_cfg)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='vfs_b200', choices=['vfs_b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the train / affinity_480p sub-benches')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------- helpers
def seg_input(gen, torch):
    """First-frame label map [1,H,W] with CV-1 rectangular objects."""
    seg = torch.zeros(1, SIZE, SIZE)
    for o in range(1, CV):
        y0, x0 = 30 * o, 40 * o
        seg[0, y0:y0 + 80, x0:x0 + 90] = o
    return seg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        loaded = [s for s in sm if s > 0.6 * max(sm)] if sm else []
        return dict(sm_mhz=statistics.median(loaded) if loaded else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def oracle_pair_step(torch, oracle, sd, frames, seg_onehot, mask):
    """The reference's CPU path for ONE frame pair (oracle restatement): features of both frames, then
    masked_attention_efficient from frame 0 to frame 1."""
    with torch.no_grad():
        feats = oracle.resnet_forward(sd, frames, 50, strides=(1, 2, 1, 1), out_indices=(2, ))
        q = feats[1:2]
        k = feats[0:1].unsqueeze(2)
        return oracle.masked_attention_efficient(q, k, seg_onehot.unsqueeze(2), mask, temperature=0.07, topk=10)


def usable_cores():
    """Host threads this process may actually use: CPU affinity capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        with open('/sys/fs/cgroup/cpu.max') as fh:
            quota, period = fh.read().split()
            if quota != 'max':
                n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def pick_cpu_threads(torch, run_once):
    """The reference arm gets 'all the host threads it can use': try the usable core count and a few smaller
    pool sizes (oversubscribed intra-op pools make torch CPU convs pathologically slow) and keep the fastest."""
    limit = usable_cores()
    cands = sorted({c for c in (8, 16, 32, 64, limit) if c <= limit} | {limit})
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        run_once()
        t0 = time.perf_counter()
        run_once()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_reference_setup(torch):
    import oracle
    from vfs_b200.backbones import ResNet
    net = ResNet(50, norm_cfg=dict(type='SyncBN', requires_grad=True), strides=(1, 2, 1, 1), out_indices=(2, ))
    sd = oracle.seeded_state_dict(net, seed=0)
    g = torch.Generator().manual_seed(1234)
    frames = torch.randn(FRAMES, 3, SIZE, SIZE, generator=g)
    fh = SIZE // 8
    lab = torch.randint(0, CV, (1, fh, fh), generator=g)
    seg_onehot = torch.nn.functional.one_hot(lab, CV).permute(0, 3, 1, 2).float()
    mask = oracle.spatial_neighbor(fh, fh, TEST_CFG['neighbor_range'])
    return oracle, sd, frames, seg_onehot, mask


def run_reference_arm(a):
    """--impl reference: the reference's CPU implementation (oracle port) on the host cores; rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    oracle, sd, frames, seg_onehot, mask = cpu_reference_setup(torch)
    cores = pick_cpu_threads(torch, lambda: oracle_pair_step(torch, oracle, sd, frames, seg_onehot, mask))
    for _ in range(max(a.warmup, 1)):
        oracle_pair_step(torch, oracle, sd, frames, seg_onehot, mask)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        oracle_pair_step(torch, oracle, sd, frames, seg_onehot, mask)
    dt = time.perf_counter() - t0
    value = a.steps / dt
    line = dict(metric='frame-pairs/sec (R50 res4 feat+affinity)', value=value, unit='frame-pairs/s', n_gpus=a.gpus,
                steps=a.steps, warmup=a.warmup, ms_per_step=dt / a.steps * 1e3, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=WORKLOAD, step='1 frame pair per step (bounded sample of the 8-clip batch)',
                            device='cpu'),
                cpu_baseline=dict(value=value, unit='frame-pairs/s', cores=cores, kind='port',
                                  sample=f'{a.steps} frame pairs, torch CPU fp32, {cores} threads (fastest pool size, {usable_cores()} usable cores)'),
                e2e=dict(value=value, unit='frame-pairs/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------- train step
TRAIN_MODEL = dict(
    type='SimSiamBaseTracker',                       # model dict of configs/r50_nc_sgd_cos_100e_r5_1xNx2_k400.py:2-24
    backbone=dict(type='ResNet', pretrained=None, depth=50, out_indices=(3, ),
                  norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False, zero_init_residual=True),
    img_head=dict(type='SimSiamHead', in_channels=2048, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                  projection_mid_channels=2048, projection_out_channels=2048, num_predictor_fcs=2,
                  predictor_mid_channels=512, predictor_out_channels=2048, with_norm=True,
                  loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg'))


def bench_train(torch, dist, dev, rank, world, clips, size, steps, peak_tf, label):
    """SimSiam pre-training step of the r50_nc config (SURVEY cfg-2 / cfg-4): imgs [clips,2,3,1,size,size] per rank,
    forward (2 backbone + 2 head passes, cosine loss) + native backward + gradient all-reduce (world > 1) + SGD, SyncBN
    in backbone and head, replayed from ONE CUDA graph per step (vfs_b200.GraphedTrainStep).  Multi-rank: every
    collective is a peer-memory kernel inside the graph.  Returns the sub-object for the JSON line."""
    import vfs_b200
    from vfs_b200 import ops
    from vfs_b200.optim import build_optimizer
    from vfs_b200.synthetic import seeded_state_dict
    model = vfs_b200.build_model(TRAIN_MODEL, train_cfg=vfs_b200.ConfigDict(dict(intra_video=False)), test_cfg=None)
    model.load_state_dict(seeded_state_dict(model, seed=0))
    model = model.to(dev)
    model.train()
    opt = build_optimizer(model, dict(type='SGD', lr=0.05, momentum=0.9, weight_decay=1e-4))   # configs/*:134
    g = torch.Generator().manual_seed(4321 + rank)
    host = torch.randn(clips, 2, 3, 1, size, size, generator=g).pin_memory()
    imgs = host.to(dev)
    l0 = ops.LAUNCHES[0]
    step = vfs_b200.GraphedTrainStep(model, opt, dict(imgs=imgs), warmup=2)
    launches_per_step = (ops.LAUNCHES[0] - l0) // 3          # 2 warm-up steps + the captured one
    for _ in range(3):
        step(dict(imgs=imgs), log=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step(dict(imgs=imgs), log=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    # e2e: the batch comes from pinned host memory every step and the logged loss is read back (runner contract)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step(dict(imgs=host), log=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    loss = float(out['log_vars']['loss'])
    layers = model.backbone.engine.conv_layer_list((2 * clips, 3, size, size), 3)
    fwd = sum(l['flops'] for l in layers) + 2.0 * 2 * clips * (size // 2)**2 * 64 * 147
    achieved = 3 * fwd / (ms * 1e-3) / 1e12
    res = dict(workload=label, clips_per_gpu=clips, size=size, ms_per_step=ms,
               pairs_per_s=world * clips / (ms * 1e-3), unit='frame-pairs/s',
               e2e=dict(pairs_per_s=world * clips * steps / float(dt), h2d_bytes_per_step=int(host.numel() * 4),
                        d2h_bytes_per_step=8, api='GraphedTrainStep(model, optimizer)(data_batch) with a pinned host '
                                                  'batch; logged loss read back every step'),
               steps=steps, loss=loss, mode='one CUDA graph per step', launches_per_step=int(launches_per_step),
               collectives=('none (1 rank)' if world == 1 else
                            'peer-memory kernels over NVLink inside the graph: SyncBN statistics per layer (fwd+bwd), '
                            'two-shot gradient all-reduce of %.1f MB, logged scalars' % (step.flat.numel * 4 / 1e6)),
               roofline=dict(bound='tensor', achieved=achieved, peak=peak_tf, unit='TFLOP/s', frac=achieved / peak_tf,
                             flops_per_step=3 * fwd,
                             note='algorithmic fp32-equivalent conv FLOPs of fwd + dgrad + wgrad (3 x forward) / '
                                  'device step time; 3 fp16 MMAs are issued per product'),
               overflow=ops.overflow_count())
    step.flat.close()
    del step, model, opt
    torch.cuda.empty_cache()
    return res


def dp_rank_check(torch, dist, dev, rank, world):
    """Cross-rank correctness of the data-parallel step on the box the bench runs on (world > 1): every rank runs the
    SimSiam step of a small R18 model on ITS shard with SyncBN statistics exchanged and gradients averaged over the
    peer-memory communicator; rank 0 also runs the whole batch alone (cross-rank exchange switched off).  Both must
    give the same loss and gradients (the definition of SyncBN + DDP).  Criterion: loss to 1e-5, global relative L2 error
    of the gradients < 3e-2 and every tensor < 0.1 (a wrong rank scaling or a missing shard is O(1)): the forward pass differs run to run by fp32 summation order (~2e-6), which
    occasionally flips the ReLU mask of an activation sitting at zero -- on this small model one flipped element moves
    every upstream gradient by ~5e-3 even between two single-process runs (tools/repeat_check.py), so the global L2
    error is bimodal by construction."""
    import vfs_b200
    from vfs_b200 import ops, peer
    from vfs_b200.synthetic import seeded_state_dict
    cfg = dict(type='SimSiamBaseTracker',
               backbone=dict(type='ResNet', pretrained=None, depth=18, out_indices=(3, ),
                             norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False,
                             zero_init_residual=True),
               img_head=dict(type='SimSiamHead', in_channels=512, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                             projection_mid_channels=512, projection_out_channels=512, num_predictor_fcs=2,
                             predictor_mid_channels=128, predictor_out_channels=512, with_norm=True,
                             loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg'))
    per = 4
    g = torch.Generator().manual_seed(99)
    full = torch.randn(per * world, 2, 3, 1, 96, 96, generator=g)

    def grads_of(imgs, cross):
        m = vfs_b200.build_model(cfg, train_cfg=vfs_b200.ConfigDict(dict(intra_video=False)), test_cfg=None)
        m.load_state_dict(seeded_state_dict(m, seed=1))
        m = m.to(dev)
        m.train()
        ops.CROSS_RANK_SYNCBN[0] = cross
        try:
            losses = m(imgs=imgs.to(dev))
            loss = sum(v.mean() for k, v in losses.items() if 'loss' in k)
            loss.backward()
        finally:
            ops.CROSS_RANK_SYNCBN[0] = True
        return loss.detach(), [p.grad.reshape(-1) for n_, p in m.named_parameters()
                               if p.grad is not None and not (n_.endswith('.bias') and 'fcs' in n_)]
        # (Linear biases in front of a BatchNorm have mathematically zero gradients: pure rounding noise, left out)

    loss, grads = grads_of(full[rank * per:(rank + 1) * per], True)
    sizes = [g_.numel() for g_ in grads]
    flat = torch.cat(grads)
    n = flat.numel() // 4 * 4
    comm = peer.active()
    buf = comm.data()[:n * 4].view(torch.float32)
    buf.copy_(flat[:n])
    comm.allreduce_(buf, 1.0 / world)
    packed = torch.stack([loss.float() / world])
    ops.cross_rank_sum_(packed)
    comm.check()
    res = None
    if rank == 0:
        ref_loss, ref_grads = grads_of(full, False)
        ref = torch.cat(ref_grads)[:n]
        err = float((buf - ref).norm() / ref.norm())
        per_tensor, off = [], 0
        for sz in sizes:
            if off + sz <= n:
                per_tensor.append(float((buf[off:off + sz] - ref[off:off + sz]).norm() /
                                        ref[off:off + sz].norm().clamp_min(1e-30)))
            off += sz
        per_tensor.sort()
        med = per_tensor[len(per_tensor) // 2]
        loss_err = abs(float(packed[0]) - float(ref_loss))
        res = dict(model='R18 SimSiam, %d clips x 2 views x 96^2 per rank' % per, grad_rel_l2_err=err,
                   median_tensor_rel_err=med, loss_abs_err=loss_err,
                   max_tensor_rel_err=per_tensor[-1],
                   ok=bool(loss_err < 1e-5 and err < 3e-2 and per_tensor[-1] < 0.1))
    dist.barrier()
    return res


def bench_siamfc(torch, dev, steps=30):
    """SURVEY cfg-5: SiamFC object-level tracking, R18 backbone with the default_config_base.py overrides (dilations
    (1,1,2,4), strides (1,2,1,1), frozen, eval BN), 127 px exemplar / three 255 px search crops, SiamConvFC head.
    ``updates_per_s``: TrackerSiamFC.update(img) end to end (cv2 crops on the host, H2D, backbone, head, fused response
    peak, 12-byte D2H); ``device_us_per_update``: backbone + head + peak on pre-staged crops (CUDA events)."""
    import numpy as np
    from vfs_b200 import ops
    from vfs_b200.siamfc import TrackerSiamFC, build_cfg
    from vfs_b200.synthetic import seeded_state_dict
    cfg = build_cfg(dict(type='ResNet', depth=18, pretrained=None, norm_cfg=dict(type='BN', requires_grad=True)),
                    exemplar_sz=127, out_scale=1e-3)
    trk = TrackerSiamFC(cfg, device=dev)
    trk.net.backbone.load_state_dict(seeded_state_dict(trk.net.backbone, seed=3))
    trk.net.head.load_state_dict(seeded_state_dict(trk.net.head, seed=4))
    trk.net.to(dev)
    rng = np.random.RandomState(0)
    frames = [rng.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(4)]
    box = [300, 200, 80, 60]
    trk.init(frames[0], box)
    for i in range(5):
        trk.update(frames[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        trk.update(frames[i % 4])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    crops = np.stack([rng.randint(0, 256, (255, 255, 3)).astype(np.uint8) for _ in range(3)])
    x = trk._to_device(crops)

    def device_part():
        feats = trk.net.backbone(x)
        r = trk.net.head(trk.kernel, feats).squeeze(1)
        return ops.siamfc_response_peak(r, trk._hann_dev, trk.upscale_sz, cfg.scale_penalty, cfg.window_influence)

    with torch.no_grad():
        for _ in range(5):
            device_part()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            device_part()
        e1.record()
        torch.cuda.synchronize()
    return dict(workload='SiamFC R18 (dilated, frozen) 127/255 exemplar/search, 3 scales, SiamConvFC head',
                updates_per_s=1.0 / wall, ms_per_update=wall * 1e3,
                device_us_per_update=e0.elapsed_time(e1) / steps * 1e3,
                api='TrackerSiamFC.update(img): host cv2 crops + H2D + backbone + x-corr head + fused response peak')


def bench_affinity_480p(torch, dev, T, peaks, peak_tf, flush):
    """SURVEY cfg-3: DAVIS-style propagation of one 480p query frame (R50 res4 map 60x107, C = 1024, Cv = 4, radius 18,
    top-k 10) against T key frames; T = 21 is the steady state of VanillaTracker.forward_test with frame 0 in the key
    set twice (vanilla_tracker.py:133-149).  The fused kernels run from a CUDA graph; L2 is flushed between
    iterations."""
    from vfs_b200 import ops
    from vfs_b200.common import spatial_neighbor
    H, W, C, Cv = 60, 107, 1024, 4
    hw = H * W
    gen = torch.Generator(device='cuda').manual_seed(T)
    F = T + 1
    bank = torch.empty((2, F, H, W, C), dtype=torch.float16, device=dev)
    for f in range(F):      # frame by frame: the fp32 NCHW staging copy of 21 frames would be 550 MB
        bank[:, f:f + 1] = ops.features_to_split(torch.relu(torch.randn(1, C, H, W, device=dev, generator=gen)), True)
    vals = torch.rand(F, Cv, hw, device=dev, generator=gen)
    mask = spatial_neighbor(1, H, W, 36)
    keys = list(range(T)) if T < 21 else [0] + list(range(T - 1))      # frame 0 twice, like the tracker's key set
    ids = [keys]
    out = torch.empty((1, Cv, hw), dtype=torch.float32, device=dev)

    def call():
        out.copy_(ops.attention_bank_batched(bank, [T], bank, ids, vals, ids, 0, Cv * hw, hw, Cv, mask, 0.07, 10))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            call()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call()
    times = []
    for _ in range(12):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times[2:])
    alg_bytes = 4.0 * ((1 + T) * C * hw + T * Cv * hw + Cv * hw)                # SURVEY 8d compulsory bytes
    alg_flops = 2.0 * C * T * float(mask.dense().sum())                          # window-restricted products
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'affinity_traffic.json')) as fh_:
            traffic = json.load(fh_).get(f'T{T}_dram_bytes')
    except Exception:
        pass
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    tf = alg_flops / (ms * 1e-3) / 1e12
    return dict(workload=f'480p 60x107 C=1024 Cv=4 radius 18 top-k 10, T={T} key frames', T=T, us_per_frame=ms * 1e3,
                key_frames_per_s=T / (ms * 1e-3),
                roofline=dict(bound='hbm', achieved=gbs, peak=hbm_peak, unit='GB/s', frac=gbs / hbm_peak,
                              traffic=traffic, bytes_per_launch=alg_bytes,
                              note='compulsory bytes 4*[(1+T)*C*HW + T*Cv*HW + Cv*HW] / CUDA-event time of the fused '
                                   'kernels (scores+top-k, merge+propagate); L2 flushed'),
                roofline_tensor=dict(bound='tensor', achieved=tf, peak=peak_tf, unit='TFLOP/s', frac=tf / peak_tf,
                                     flops_per_launch=alg_flops,
                                     note='window-restricted fp32-equivalent FLOPs 2*C*T*sum_q|N(q)|'))


# ----------------------------------------------------------------------------------------------------- main arm
def main():
    a = parse()
    if a.impl == 'reference':
        run_reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the B200 path has no CPU fallback); '
                         'use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    import vfs_b200
    from vfs_b200 import ops
    from vfs_b200.common import spatial_neighbor
    from vfs_b200.synthetic import seeded_state_dict

    model = vfs_b200.build_model(dict(type='VanillaTracker', backbone=BACKBONE_CFG), train_cfg=None,
                                 test_cfg=vfs_b200.ConfigDict(TEST_CFG))
    model.backbone.load_state_dict(seeded_state_dict(model.backbone, seed=0))
    model = model.to(dev)
    model.eval()
    eng = model.backbone.engine
    eng.check_versions = False  # weights are frozen for the whole run

    g = torch.Generator().manual_seed(1234 + rank)
    imgs_host = torch.randn(CLIPS, 1, 3, FRAMES, SIZE, SIZE, generator=g).pin_memory()   # forward_test layout
    seg_host = seg_input(g, torch).pin_memory()
    # device-resident copy in the batched layout [clip][frame] -> frames [16,3,256,256]
    frames_dev = imgs_host.to(dev)[:, 0].permute(0, 2, 1, 3, 4).reshape(CLIPS * FRAMES, 3, SIZE, SIZE).contiguous()
    fh = fw = SIZE // 8
    hw = fh * fw
    lab = torch.randint(0, CV, (CLIPS, hw), generator=g)
    seg_bank = torch.zeros(CLIPS * FRAMES, CV, hw, device=dev)
    seg_bank[0::2] = torch.nn.functional.one_hot(lab, CV).permute(0, 2, 1).float().to(dev)
    mask = spatial_neighbor(1, fh, fw, TEST_CFG['neighbor_range'])
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    bank = torch.empty((2, CLIPS * FRAMES, fh, fw, 1024), dtype=torch.float16, device=dev)
    out = torch.empty((CLIPS, CV, hw), dtype=torch.float32, device=dev)

    # The step is four CUDA graphs (stem | tcgen05 conv stages | normalise | attention) captured once over static
    # buffers: replaying them removes the Python/ctypes issue cost of the ~60 launches and lets CUDA events between
    # the graphs time each segment on the device.
    state = {}

    def seg_stem():
        state['stem'] = eng.stem(frames_dev)

    def seg_convs():
        state['feat'] = eng.run_stages(state['stem'], 2)      # 16 frames -> res4, split NHWC

    def seg_norm():
        ops.normalize_split(state['feat'], out=bank)

    def seg_attn():
        # frame 2c = key (labels known), 2c+1 = query; the 8 clips are 8 problems of ONE launch
        ids = [[2 * c] for c in range(CLIPS)]
        out.copy_(ops.attention_bank_batched(bank, [2 * c + 1 for c in range(CLIPS)], bank, ids, seg_bank, ids, 0,
                                             CV * hw, hw, CV, mask, TEST_CFG['temperature'], TEST_CFG['topk']))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (eager: builds plans, sets kernel attributes), then capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(a.warmup, 3)):
            seg_stem(); seg_convs(); seg_norm(); seg_attn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    launches0 = ops.LAUNCHES[0]
    graphs = []
    for seg in (seg_stem, seg_convs, seg_norm, seg_attn):
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            seg()
        graphs.append(g_)
    launches_per_step = ops.LAUNCHES[0] - launches0           # kernels of libvfs_b200 captured per step
    for _ in range(2):
        for g_ in graphs:
            g_.replay()
    barrier()

    # ---------------- timed: K steps, device time per step (CUDA events on the launching stream), L2 flushed between
    sampler = ClockSampler(local_rank)
    sampler.start()
    layer_list = eng.conv_layer_list((CLIPS * FRAMES, 3, SIZE, SIZE), 2)
    conv_flops = sum(l['flops'] for l in layer_list)
    n_conv_launches = len(layer_list)
    marks = []
    barrier()
    cuprof = bool(os.environ.get('VFS_BENCH_CUPROFILE'))   # ncu --profile-from-start off: only the timed steps
    if cuprof:
        torch.cuda.profiler.start()
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        flush.fill_(1)                                        # evict L2 (512 MiB > 126 MB); outside the timed span
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        for gi, g_ in enumerate(graphs):
            g_.replay()
            ev[gi + 1].record()
        marks.append(ev)
    barrier()
    wall = time.perf_counter() - wall0
    if cuprof:
        torch.cuda.profiler.stop()
    launches = launches_per_step * a.steps
    step_ms = [e[0].elapsed_time(e[4]) for e in marks]
    stem_ms = [e[0].elapsed_time(e[1]) for e in marks]
    conv_ms = [e[1].elapsed_time(e[2]) for e in marks]
    norm_ms = [e[2].elapsed_time(e[3]) for e in marks]
    attn_ms = [e[3].elapsed_time(e[4]) for e in marks]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    clocks = sampler.stop()
    value = world * CLIPS * a.steps / (total_ms / 1e3)

    # ---------------- e2e through the public API: forward_test(imgs, ref_seg_map, img_meta) with host buffers
    meta = [dict(original_shape=(SIZE, SIZE, 3))]
    seg8_host = seg_host.expand(CLIPS, SIZE, SIZE).contiguous().pin_memory()

    def step_e2e():
        imgs = imgs_host.to(dev, non_blocking=True)                  # [8,1,3,2,H,W]  H2D
        # the label maps go in as the pinned host tensor: forward_test copies them itself and reads the class count
        # on the host instead of synchronising on the device maximum
        return model.forward_test(imgs, seg8_host, meta * CLIPS)        # 8 numpy arrays on the host (D2H inside)

    def step_e2e_per_video():
        res = []
        for c in range(CLIPS):
            imgs = imgs_host[c:c + 1].to(dev, non_blocking=True)        # [1,1,3,2,H,W]  H2D
            seg = seg_host.to(dev, non_blocking=True)
            res.append(model.forward_test(imgs, seg, meta)[0])
        return res

    def time_e2e(fn, steps):
        for _ in range(3):
            r_ = fn()
        barrier()
        t0_ = time.perf_counter()
        for _ in range(steps):
            r_ = fn()
        torch.cuda.synchronize()
        dt_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
        return world * CLIPS * steps / float(dt_), r_

    # the same call fed the way the reference's loader would feed it if Normalize + FormatShape ran on the device:
    # uint8 HWC frames cross PCIe (a quarter of the bytes), vfs_b200.DeviceNormalizeFormat builds the fp32 clip tensor
    u8_host = torch.randint(0, 256, (CLIPS, FRAMES, SIZE, SIZE, 3), dtype=torch.uint8, generator=g).pin_memory()
    feed = vfs_b200.DeviceNormalizeFormat(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_bgr=False)

    def step_e2e_u8():
        imgs = feed(u8_host).unsqueeze(1)                               # H2D of uint8 frames, [8,1,3,2,H,W] fp32 in HBM
        return model.forward_test(imgs, seg8_host, meta * CLIPS)

    # the same feed double-buffered: the pinned -> device copy of step i+1 is issued (on a side stream) before step i's
    # kernels, so it overlaps them; every step still moves its own frames H2D and its own predictions D2H
    ring = vfs_b200.PinnedRing(slots=2)

    def step_e2e_ring():
        if ring._count == 0:
            ring.put(u8_host)
        cur = ring.get()
        ring.put(u8_host)                                               # next step's frames
        return model.forward_test(feed(cur).unsqueeze(1), seg8_host, meta * CLIPS)

    # the reference's evaluation entry point (tools/test.py -> mmaction/apis/test.py:15 single_gpu_test) over a loader
    # of pinned host batches: label maps in the loader's dtype (uint8: RawFrameDecode reads the palette PNG,
    # loading.py:1048-1053, ToTensor keeps it), frames as the fp32 NCTHW tensor Normalize + FormatShape produce.
    # Every step copies its own frames H2D and its own predictions D2H; the driver keeps two calls in flight.
    from vfs_b200.apis import single_gpu_test
    seg8_u8_host = seg8_host.to(torch.uint8).pin_memory()

    def loader_of(n):
        return [dict(imgs=imgs_host, ref_seg_map=seg8_u8_host, img_meta=meta * CLIPS) for _ in range(n)]

    def time_e2e_driver(steps, depth):
        single_gpu_test(model, loader_of(4), pipeline_depth=depth)
        barrier()
        t0_ = time.perf_counter()
        r_ = single_gpu_test(model, loader_of(steps), pipeline_depth=depth)
        torch.cuda.synchronize()
        dt_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
        assert len(r_) == steps * CLIPS
        return world * CLIPS * steps / float(dt_), r_[-CLIPS:]

    e2e_steps = max(3, min(a.steps, 20))
    e2e_value, r = time_e2e_driver(max(e2e_steps, 20), 2)
    e2e_blocking, _ = time_e2e_driver(max(e2e_steps, 20), 1)
    e2e_call, r_call = time_e2e(step_e2e, e2e_steps)

    # the reference's own evaluation flow: videos_per_gpu = 1 (configs/*:106), one video per forward_test call, driven
    # by single_gpu_test -- here with two calls in flight
    def time_e2e_per_video_driver(steps):
        loader = [dict(imgs=imgs_host[c:c + 1], ref_seg_map=seg8_u8_host[c:c + 1], img_meta=meta)
                  for _ in range(steps) for c in range(CLIPS)]
        single_gpu_test(model, loader[:2 * CLIPS])
        barrier()
        t0_ = time.perf_counter()
        r_ = single_gpu_test(model, loader)
        torch.cuda.synchronize()
        dt_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
        return world * CLIPS * steps / float(dt_), r_[-CLIPS:]

    e2e_single_driver, r1d = time_e2e_per_video_driver(max(3, min(a.steps, 10)))
    assert all((x == y).all() for x, y in zip(r, r1d)), 'per-video driver and batched driver disagree'
    assert all((x == y).all() for x, y in zip(r, r_call)), 'pipelined driver and single forward_test call disagree'
    e2e_u8, _ = time_e2e(step_e2e_u8, e2e_steps)
    e2e_ring, r_ring = time_e2e(step_e2e_ring, e2e_steps)
    e2e_single, r1 = time_e2e(step_e2e_per_video, max(3, min(a.steps, 10)))
    assert all((x == y).all() for x, y in zip(r, r1)), 'batched and per-video forward_test disagree'
    h2d = imgs_host.numel() * 4 + seg8_u8_host.numel()
    d2h = sum(int(x.nbytes) for x in r)

    # ---------------- roofline of the dominant kernel (tcgen05 conv): algorithmic FLOPs / CUDA-event time
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh_:
            peaks = json.load(fh_)
    except Exception:
        pass
    # the timed region is tens of milliseconds at full clocks -> the BURST peak is the fair denominator
    peak_tf = peaks.get('bf16_tflops') or 1590.0
    peak_src = 'measured burst (MEASURED_PEAKS.json bf16_tflops)' if peaks else 'fallback 1.59 PFLOP/s'
    traffic = None
    try:   # DRAM bytes of the 42 conv launches of one step, from the committed ncu --set full capture of this command
        with open(os.path.join(ROOT, 'profiles', 'conv_traffic.json')) as fh_:
            traffic = json.load(fh_).get('dram_bytes_per_step')
    except Exception:
        pass
    conv_med = statistics.median(conv_ms)
    achieved = conv_flops / (conv_med * 1e-3) / 1e12
    roofline = dict(bound='tensor', kernel='conv_tc_kernel', achieved=achieved, peak=peak_tf, unit='TFLOP/s',
                    frac=achieved / peak_tf, traffic=traffic, peak_source=peak_src,
                    note='achieved = algorithmic fp32-equivalent conv FLOPs; the kernel issues 3 fp16 MMAs per '
                         'product (split-fp16), so tensor-pipe work is 3x: frac_of_issued = %.3f' %
                         (3 * achieved / peak_tf),
                    launches_per_step=n_conv_launches, segment_ms_median=conv_med,
                    flops_per_step=conv_flops)

    # second named kernel (north_star): fused affinity + top-k + softmax propagation.  Algorithmic work = the
    # window-restricted products 2*C*sum_q |N(q)| per problem / CUDA-event time of the attention graph.
    pairs_in_window = int(mask.dense().sum())
    attn_flops = 2.0 * 1024 * pairs_in_window * CLIPS
    attn_med = statistics.median(attn_ms)
    attn_tf = attn_flops / (attn_med * 1e-3) / 1e12
    attn_bytes = 4.0 * CLIPS * (2 * 1024 * hw + 2 * CV * hw)          # q + k features, values in, labels out
    attn_traffic = None        # ncu dram__bytes of the two attention kernels of one step (profiles/r02_ncu_step.json)
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_ncu_step.json')) as fh_:
            attn_traffic = sum(k_.get('dram_read_bytes', 0.0) + k_.get('dram_write_bytes', 0.0)
                               for k_ in json.load(fh_)['kernels'] if k_['kernel'].startswith('attn_'))
    except Exception:
        pass
    roofline_affinity = dict(bound='tensor', kernel='attn_scores_topk_kernel + attn_merge_propagate_kernel',
                             achieved=attn_tf, peak=peak_tf, unit='TFLOP/s', frac=attn_tf / peak_tf,
                             traffic=attn_traffic,
                             segment_ms_median=attn_med, flops_per_step=attn_flops,
                             hbm_gbs_algorithmic=attn_bytes / (attn_med * 1e-3) / 1e9,
                             note='window-restricted fp32-equivalent FLOPs (3 MMAs issued per product, and whole '
                                  '128x128 key tiles are multiplied: the dense-tile work is larger); the fused kernel '
                                  'is tensor/shared-memory bound, its compulsory HBM bytes would take %.1f us at the '
                                  'measured copy bandwidth' % (attn_bytes / (peaks.get('hbm_gbs', 6650.0) * 1e3)))

    line = dict(metric='frame-pairs/sec (R50 res4 feat+affinity)', value=value, unit='frame-pairs/s', n_gpus=world,
                steps=a.steps, warmup=max(a.warmup, 3), ms_per_step=total_ms / a.steps, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='fp16x3 (split-fp16 operands, fp32 accumulate)',
                data='synthetic',
                config=dict(workload=WORKLOAD, clips_per_gpu=CLIPS, l2='flushed (512 MiB write) between steps',
                            timing='sum of per-step CUDA-event spans (4 CUDA graphs per step), max over ranks',
                            parallelism=f'dp{world} (clips sharded, no collective)'),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit='frame-pairs/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         api='vfs_b200.apis.single_gpu_test(build_model(VanillaTracker), loader) -- the reference\'s '
                             'evaluation driver (apis/test.py:15) -- over a loader of pinned host batches of 8 two-frame '
                             'videos (fp32 NCTHW frames, uint8 label maps); two forward_test calls in flight, each '
                             'step moves its own frames H2D and predictions D2H', steps=max(e2e_steps, 20),
                         blocking_driver=dict(value=e2e_blocking, unit='frame-pairs/s',
                                              api='same driver, pipeline_depth=1 (collect every call before the next)'),
                         single_call=dict(value=e2e_call, unit='frame-pairs/s',
                                          h2d_bytes_per_step=imgs_host.numel() * 4 + seg8_host.numel() * 4,
                                          d2h_bytes_per_step=sum(int(x.nbytes) for x in r_call),
                                          api='one blocking forward_test call per step, float32 label maps '
                                              '(the round-1 definition of e2e)'),
                         per_video_calls=dict(value=e2e_single, unit='frame-pairs/s',
                                              api='one blocking forward_test call per video (reference calling '
                                                  'convention)'),
                         per_video_driver=dict(value=e2e_single_driver, unit='frame-pairs/s',
                                               api='single_gpu_test over a loader of single-video batches '
                                                   '(videos_per_gpu=1 like the reference configs), two calls in flight'),
                         from_uint8_frames=dict(value=e2e_u8, unit='frame-pairs/s', h2d_bytes_per_step=int(u8_host.numel()),
                                                api='uint8 HWC host frames -> DeviceNormalizeFormat (Normalize + '
                                                    'FormatShape on the device) -> forward_test'),
                         from_uint8_frames_ring=dict(value=e2e_ring, unit='frame-pairs/s',
                                                     h2d_bytes_per_step=int(u8_host.numel()),
                                                     api='same through vfs_b200.PinnedRing: the H2D copy of step i+1 '
                                                         'overlaps the kernels of step i')),
                gpu_launches=launches,
                roofline=roofline,
                roofline_affinity=roofline_affinity,
                breakdown_ms=dict(step_median=statistics.median(step_ms), stem_median=statistics.median(stem_ms),
                                  convs_median=conv_med, normalize_median=statistics.median(norm_ms),
                                  attention_median=attn_med),
                wall_s_timed_region=wall)

    # ---------------- the north_star's own configs as sub-objects (each with its roofline)
    if not a.no_extra:
        if world == 1:
            line['affinity_480p'] = dict(T1=bench_affinity_480p(torch, dev, 1, peaks, peak_tf, flush),
                                         T21=bench_affinity_480p(torch, dev, 21, peaks, peak_tf, flush))
            line['siamfc'] = bench_siamfc(torch, dev)
            line['train_cfg2'] = bench_train(torch, dist, dev, rank, world, 8, 256, max(3, min(a.steps, 10)), peak_tf,
                                             'SURVEY cfg-2: R50 SimSiam train step, 8 clips x 2 views x 256^2, 1 GPU')
        if world > 1:
            from vfs_b200 import peer
            peer.install(peer.PeerComm(data_bytes=160 * 1024 * 1024))      # flat R50 SimSiam gradients: 152.8 MB
            line['dp_rank_check'] = dp_rank_check(torch, dist, dev, rank, world)
        line['train'] = bench_train(torch, dist, dev, rank, world, 32, 224, max(3, min(a.steps, 10)), peak_tf,
                                    'SURVEY cfg-4: r50_nc_sgd_cos_100e_r5_1xNx2_k400 train step, 32 clips x 2 views x '
                                    '224^2 per GPU, SyncBN + gradient all-reduce, data-parallel over %d GPU(s)' % world)

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        o, sd, frames, seg_onehot, m = cpu_reference_setup(torch)
        cores = pick_cpu_threads(torch, lambda: oracle_pair_step(torch, o, sd, frames, seg_onehot, m))
        n, t0 = 0, time.perf_counter()
        while n < 2 or time.perf_counter() - t0 < 10.0:
            oracle_pair_step(torch, o, sd, frames, seg_onehot, m)
            n += 1
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = dict(value=n / dt, unit='frame-pairs/s', cores=cores, kind='port',
                                    sample=f'{n} frame pairs (2 frames 256x256 -> res4 + attention), torch CPU fp32, {cores} threads')
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
