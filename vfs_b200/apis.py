"""Evaluation drivers with the reference's call signatures (mmaction/apis/test.py:15-194): run a model over a data
loader with ``return_loss=False`` and, across ranks, return the predictions in dataset order on rank 0.

Videos are independent units, so multi-GPU inference is plain sharding (DistributedSampler: sample i goes to rank
i % world) with ONE host-side exchange at the end -- no data-path collective.  The reference pickles every rank's
list into a padded uint8 CUDA tensor and all-gathers it (or goes through a shared tmp directory); here the lists are
gathered as Python objects to rank 0 (``dist.gather_object``: the process group's own byte transport -- NCCL on the
box, gloo in the CPU tests), and interleaved exactly like the reference (``zip(*parts)``, truncated to the dataset
size because the sampler pads)."""
import torch
import torch.distributed as dist


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def single_gpu_test(model, data_loader):
    """mmaction/apis/test.py:15-45: list of per-sample results."""
    model.eval()
    results = []
    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, **data)
        if isinstance(result, list):        # reference test.py:36-39: lists are flattened, anything else is one result
            results.extend(result)
        else:
            results.append(result)
    return results


def collect_results(result_part, size):
    """Per-rank result lists -> ordered list of ``size`` results on rank 0, ``None`` elsewhere (collect_results_gpu /
    collect_results_cpu, test.py:98-194)."""
    rank, world = _dist_info()
    if world == 1:
        return list(result_part)[:size]
    parts = [None] * world if rank == 0 else None
    dist.gather_object(list(result_part), parts, dst=0)
    if rank != 0:
        return None
    ordered = []
    for group in zip(*parts):          # sample order of DistributedSampler: rank-major inside each round
        ordered.extend(group)
    return ordered[:size]


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=True):
    """mmaction/apis/test.py:47-95.  ``tmpdir`` / ``gpu_collect`` select the reference's transport and are accepted for
    signature compatibility; the gather goes through the process group either way."""
    results = single_gpu_test(model, data_loader)
    return collect_results(results, len(data_loader.dataset))
