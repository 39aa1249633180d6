"""Cross-rank correctness probe of the data-parallel step (run under torchrun with N ranks):

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/dp_check.py [--nccl]

Every rank runs the SimSiam step of a small R18 model on its shard (SyncBN statistics exchanged, gradients averaged),
rank 0 also runs the whole batch alone; prints the per-tensor relative errors, worst first.  --nccl uses
torch.distributed collectives instead of the peer-memory communicator."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vfs_b200  # noqa: E402
from vfs_b200 import ops, peer  # noqa: E402
from vfs_b200.synthetic import seeded_state_dict  # noqa: E402


def main():
    use_nccl = '--nccl' in sys.argv
    size = 64
    for a in sys.argv[1:]:
        if a.startswith('--size='):
            size = int(a.split('=')[1])
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    if not use_nccl:
        peer.install(peer.PeerComm(data_bytes=64 * 1024 * 1024))
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
    full = torch.randn(per * world, 2, 3, 1, size, size, generator=g)

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
        return loss.detach(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}

    loss, grads = grads_of(full[rank * per:(rank + 1) * per], True)
    for k in grads:
        dist.all_reduce(grads[k])
        grads[k] /= world
    if rank == 0:
        ref_loss, ref = grads_of(full, False)
        ref2_loss, ref2 = grads_of(full, False)     # run-to-run noise of the single-process step (fp32 atomics)
        rows = []
        for k in ref:
            d = float(ref[k].norm()) + 1e-30
            rows.append((float((grads[k] - ref[k]).norm()) / d, float((ref2[k] - ref[k]).norm()) / d, k))
        rows.sort(reverse=True)
        print(f'world {world} {"nccl" if use_nccl else "peer"} size {size}: loss {float(loss):.7f} (shard 0) ref '
              f'{float(ref_loss):.7f}')
        for e, n, k in rows[:12]:
            print(f'  {e:.3e}  (run-to-run {n:.3e})  {k}')
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
