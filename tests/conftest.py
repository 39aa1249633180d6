import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (sm_100a) GPU; run with `pytest -m gpu` on the GPU box')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    path = os.path.join(ROOT, 'tests', 'golden', 'vfs_golden.npz')
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope='session', autouse=True)
def _torch_cpu_threads():
    """The oracle runs on torch CPU: keep its intra-op pool within the cores this process may use (GPU boxes expose
    many more logical CPUs than the container's affinity mask / quota allows)."""
    try:
        import torch
        torch.set_num_threads(max(1, min(16, len(os.sched_getaffinity(0)))))
    except Exception:
        pass
    yield
