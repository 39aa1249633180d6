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


def _model_device(model):
    for t in model.parameters():
        return t.device
    return torch.device('cpu')


class _DeviceFeed:
    """What MMDataParallel.scatter does for the reference (mmaction/apis/test.py:31 runs the model through it): move
    the tensors of a loader batch to the model's device.  Here the copy of batch i+1 is issued on a side stream before
    batch i's kernels are enqueued (vfs_b200.PinnedRing), so it overlaps them; inputs a model lists in
    ``host_inputs`` stay on the host."""

    def __init__(self, model, device):
        from .pipelines import PinnedRing
        self.keep = tuple(getattr(model, 'host_inputs', ()))
        self.device = device
        self.rings = {}
        self.PinnedRing = PinnedRing

    def _movable(self, key, v):
        return torch.is_tensor(v) and not v.is_cuda and key not in self.keep

    def put(self, data):
        for k, v in data.items():
            if self._movable(k, v):
                ring = self.rings.get(k)
                if ring is None:
                    ring = self.rings[k] = self.PinnedRing(slots=3, device=self.device)
                ring.put(v)

    def get(self, data):
        return {k: (self.rings[k].get() if self._movable(k, v) else v) for k, v in data.items()}


def _batches_on_device(model, data_loader):
    """Yield the loader's batches with their tensors on the model's device, one batch prefetched."""
    dev = _model_device(model)
    if dev.type != 'cuda':
        yield from data_loader
        return
    feed = _DeviceFeed(model, dev)
    it = iter(data_loader)
    try:
        cur = next(it)
    except StopIteration:
        return
    feed.put(cur)
    while cur is not None:
        nxt = next(it, None)
        on_dev = feed.get(cur)
        if nxt is not None:
            feed.put(nxt)            # H2D of the next batch runs beside this batch's kernels
        yield on_dev
        cur = nxt


def single_gpu_test(model, data_loader, pipeline_depth=2):
    """mmaction/apis/test.py:15-45: list of per-sample results.

    Batches are moved to the model's device like the reference's MMDataParallel does, and a model that offers
    ``forward_test_async`` (VanillaTracker) is driven ``pipeline_depth`` calls deep: call i+1 is enqueued before the
    predictions of call i are collected, so the host-side waits (H2D, D2H, launch latency) overlap device work.  Results
    and their order are those of the plain loop."""
    model.eval()
    results, pending = [], []

    def collect(result):
        if isinstance(result, list):        # reference test.py:36-39: lists are flattened, anything else is one result
            results.extend(result)
        else:
            results.append(result)

    enqueue = getattr(model, 'forward_test_async', None) if pipeline_depth > 1 else None
    for data in _batches_on_device(model, data_loader):
        with torch.no_grad():
            if enqueue is None:
                collect(model(return_loss=False, **data))
                continue
            pending.append(enqueue(**data))
        while len(pending) >= pipeline_depth:
            collect(pending.pop(0).result())
    for handle in pending:
        collect(handle.result())
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
