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


def _signature(data):
    """What must agree for two loader batches to be merged along the batch axis."""
    sig = []
    for k in sorted(data):
        v = data[k]
        if torch.is_tensor(v):
            sig.append((k, 'tensor', tuple(v.shape[1:]), v.dtype, v.device))
        elif isinstance(v, list):
            sig.append((k, 'list'))
        else:
            return None
    return tuple(sig)


def _merge(group):
    """Concatenate loader batches along the batch axis.  Device tensors are copied piece by piece into one buffer as
    they arrive (the feed's ring slots are recycled after a few batches), host tensors and lists are concatenated."""
    if len(group) == 1:
        return group[0]
    out = {}
    for k, v in group[0].items():
        if torch.is_tensor(v):
            out[k] = torch.cat([d[k] for d in group], dim=0)
        else:
            out[k] = [x for d in group for x in d[k]]
    return out


def _coalesced(batches, max_videos):
    """Group consecutive batches of identical layout into one call of at most ``max_videos`` samples.  The reference
    configs evaluate with videos_per_gpu = 1 (configs/*:106), i.e. one two-to-hundred-frame video per forward_test call;
    VanillaTracker.forward_test takes B equal-length videos per call with results identical to B separate calls
    (tests: test_vanilla_tracker_batched_videos_equal_single_video_calls), and small launches are what a per-video call
    loses on 148 SMs.  Videos of different length or resolution are never merged."""
    group, sig, count = [], None, 0
    for data in batches:
        s = _signature(data)
        n = len(next((v for v in data.values() if torch.is_tensor(v)), [0]))
        if group and (s is None or s != sig or count + n > max_videos):
            yield _merge(group)
            group, count = [], 0
        if s is None or (not group and n >= max_videos):
            yield data
            continue
        # device tensors are cloned out of the feed's ring right away: the slot is recycled a few batches later
        group.append({k: (v.clone() if (torch.is_tensor(v) and v.is_cuda) else v) for k, v in data.items()})
        sig, count = s, count + n
        if count >= max_videos:
            yield _merge(group)
            group, count = [], 0
    if group:
        yield _merge(group)


def single_gpu_test(model, data_loader, pipeline_depth=2, coalesce=None):
    """mmaction/apis/test.py:15-45: list of per-sample results.

    Batches are moved to the model's device like the reference's MMDataParallel does, and a model that offers
    ``forward_test_async`` (VanillaTracker) is driven ``pipeline_depth`` calls deep: call i+1 is enqueued before the
    predictions of call i are collected, so the host-side waits (H2D, D2H, launch latency) overlap device work.
    ``coalesce`` (default: the model's ``coalesce_videos`` attribute, 8 for VanillaTracker, 1 = off) merges consecutive
    loader batches of identical layout into one call.  Results and their order are those of the plain loop."""
    model.eval()
    results, pending = [], []
    if coalesce is None:
        coalesce = int(getattr(model, 'coalesce_videos', 1)) if pipeline_depth > 1 else 1

    def collect(result):
        if isinstance(result, list):        # reference test.py:36-39: lists are flattened, anything else is one result
            results.extend(result)
        else:
            results.append(result)

    enqueue = getattr(model, 'forward_test_async', None) if pipeline_depth > 1 else None
    batches = _batches_on_device(model, data_loader)
    if coalesce > 1:
        batches = _coalesced(batches, coalesce)
    for data in batches:
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
