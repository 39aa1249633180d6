"""Multi-GPU (NCCL) parity of the data-parallel training step: with SyncBN in the backbone and the head, two ranks
each holding half of the batch must produce -- after the gradient all-reduce -- the gradients of a single process on
the whole batch, which is what the CPU oracle computes.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os

import pytest
import torch

import oracle
from tests.golden import cases

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, q):
    import torch.distributed as dist
    import vfs_b200
    from vfs_b200 import ops
    from vfs_b200.optim import allreduce_grads
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    c = cases.TRACKER_TRAIN_CASES[name]
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    model.load_state_dict(oracle.seeded_state_dict(model, seed=c['seed']))
    model = model.cuda()
    model.train()
    shape = (8, ) + tuple(c['shape'][1:])
    imgs = torch.randn(shape, generator=torch.Generator().manual_seed(900 + c['seed']))
    per = shape[0] // world
    mine = imgs[rank * per:(rank + 1) * per].cuda()
    out = model.train_step(dict(imgs=mine), None)
    out['loss'].backward()
    params = [p for p in model.parameters() if p.grad is not None]
    allreduce_grads(params, average=True)
    overflow = ops.overflow_count()
    if rank == 0:
        q.put((out['log_vars']['loss'], {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None},
               {k: v.cpu() for k, v in model.state_dict().items() if 'running_mean' in k}, overflow))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['r18_intra'])
def test_two_rank_syncbn_training_step_equals_single_process(name):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from tests.test_gpu_parity import _oracle_train_reference
    c = cases.TRACKER_TRAIN_CASES[name]
    import vfs_b200
    model = vfs_b200.build_model(c['model'], train_cfg=vfs_b200.ConfigDict(c['train_cfg']), test_cfg=None)
    sd = oracle.seeded_state_dict(model, seed=c['seed'])
    shape = (8, ) + tuple(c['shape'][1:])
    imgs = torch.randn(shape, generator=torch.Generator().manual_seed(900 + c['seed']))
    ref_loss, ref = _oracle_train_reference(c, sd, imgs)
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    _, ref64 = _oracle_train_reference(c, sd64, imgs.double())

    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss, grads, running, overflow = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert overflow == 0
    assert abs(loss - ref_loss) < 1e-3 * max(1.0, abs(ref_loss))     # logged loss = average over ranks
    gnorm = max(float(r.norm()) for r in ref64.values())
    failures = []
    for k, g in grads.items():
        r64 = ref64[k]
        denom = max(float(r64.norm()), 1e-6 * gnorm)
        mine = float((g.double() - r64).norm()) / denom
        base = float((ref[k].double() - r64).norm()) / denom
        if mine > max(20 * base, 3e-3):
            failures.append((k, mine, base))
    assert not failures, failures[:8]
