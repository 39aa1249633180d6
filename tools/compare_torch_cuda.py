"""Comparator (not product, not bench): the bench workload executed by stock torch-CUDA kernels (cuDNN / cuBLAS /
ATen top-k on sm_100) through the oracle restatement of the reference modules -- "the existing Blackwell kernels" the
hand-written path has to beat (CPU-baseline row).  Same shapes as bench.py: 16 frames 256x256, R50 to res4
with strides (1,2,1,1), then masked_attention_efficient (radius 18, top-k 10) per clip.  fp32 with TF32 convs on / off."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import oracle  # noqa: E402
from vfs_b200.backbones import ResNet  # noqa: E402  (state-dict names only)


def train_comparator(dev):
    """SimSiam R50 train step (SURVEY cfg-2 / cfg-4 per-GPU shape) on stock torch-CUDA: forward + loss + autograd
    backward (cuDNN / cuBLAS) through the oracle functions + torch.optim.SGD, eager launches, TF32 on / off."""
    import vfs_b200
    out = {}
    model = vfs_b200.build_model(bench.TRAIN_MODEL, train_cfg=vfs_b200.ConfigDict(dict(intra_video=False)), test_cfg=None)
    sd0 = oracle.seeded_state_dict(model, seed=0)
    del model
    for label, clips, size in (('cfg2_8x2x256', 8, 256), ('cfg4_32x2x224', 32, 224)):
        g = torch.Generator().manual_seed(4321)
        imgs = torch.randn(clips, 2, 3, 1, size, size, generator=g).to(dev)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            params = {k: v.clone().to(dev).requires_grad_(v.dtype.is_floating_point and 'running' not in k)
                      for k, v in sd0.items()}
            opt = torch.optim.SGD([p for p in params.values() if p.requires_grad], lr=0.05, momentum=0.9,
                                  weight_decay=1e-4)

            def step():
                opt.zero_grad(set_to_none=True)
                losses = oracle.simsiam_forward_train(params, imgs, 50, intra_video=False, bn_training=True)
                loss = sum(v.mean() for v in losses.values())
                loss.backward()
                opt.step()
                return loss

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[f'{label}_{"tf32" if tf32 else "fp32"}'] = dict(ms_per_step=ms, frame_pairs_per_s=clips / (ms * 1e-3),
                                                               loss=float(loss))
            del params, opt
            torch.cuda.empty_cache()
    print(json.dumps(dict(what='stock torch-CUDA SimSiam R50 train step (cuDNN/cuBLAS autograd through the oracle modules '
                          '+ torch.optim.SGD), eager, 1 GPU, device-timed over 10 steps; channels-first fp32 tensors, '
                          'BN running statistics not updated (the oracle clones them)', **out)))


def main():
    dev = torch.device('cuda', 0)
    if '--train' in sys.argv:
        return train_comparator(dev)
    net = ResNet(50, norm_cfg=dict(type='SyncBN', requires_grad=True), strides=(1, 2, 1, 1), out_indices=(2, ))
    sd = {k: v.to(dev) for k, v in oracle.seeded_state_dict(net, seed=0).items()}
    g = torch.Generator().manual_seed(1234)
    frames = torch.randn(bench.CLIPS * bench.FRAMES, 3, bench.SIZE, bench.SIZE, generator=g).to(dev)
    fh = bench.SIZE // 8
    lab = torch.randint(0, bench.CV, (bench.CLIPS, fh, fh), generator=g)
    seg = torch.nn.functional.one_hot(lab, bench.CV).permute(0, 3, 1, 2).float().to(dev)
    mask = oracle.spatial_neighbor(fh, fh, bench.TEST_CFG['neighbor_range']).to(dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32

        def step():
            with torch.no_grad():
                feats = oracle.resnet_forward(sd, frames, 50, strides=(1, 2, 1, 1), out_indices=(2, ))
                res = []
                for c in range(bench.CLIPS):
                    res.append(oracle.masked_attention_efficient(feats[2 * c + 1:2 * c + 2],
                                                                 feats[2 * c:2 * c + 1].unsqueeze(2),
                                                                 seg[c:c + 1].unsqueeze(2), mask, temperature=0.07,
                                                                 topk=10))
                return feats, res

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        times, bb = [], []
        for _ in range(10):
            flush.fill_(1)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            with torch.no_grad():
                feats = oracle.resnet_forward(sd, frames, 50, strides=(1, 2, 1, 1), out_indices=(2, ))
            e[1].record()
            with torch.no_grad():
                for c in range(bench.CLIPS):
                    oracle.masked_attention_efficient(feats[2 * c + 1:2 * c + 2], feats[2 * c:2 * c + 1].unsqueeze(2),
                                                      seg[c:c + 1].unsqueeze(2), mask, temperature=0.07, topk=10)
            e[2].record()
            torch.cuda.synchronize()
            times.append(e[0].elapsed_time(e[2]))
            bb.append(e[0].elapsed_time(e[1]))
        times.sort()
        bb.sort()
        med = times[len(times) // 2]
        out['tf32' if tf32 else 'fp32'] = dict(ms_per_step=med, backbone_ms=bb[len(bb) // 2],
                                               frame_pairs_per_s=bench.CLIPS / (med * 1e-3))
    print(json.dumps(dict(workload=bench.WORKLOAD, what='stock torch-CUDA (cuDNN/cuBLAS/ATen) through the oracle modules, '
                          'device-timed, eager launches, all 4 ResNet stages like the reference', **out)))


if __name__ == '__main__':
    main()
