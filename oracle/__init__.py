"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch CPU fp32 / numpy) of the reference algorithms on the VFS hot path, each function citing
the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the checker or the reported
CPU baseline -- never as a fallback of ``vfs_b200`` (which fails loudly when its CUDA library is missing).

Pinning: the reference publishes no golden vectors for this path, so the oracle is pinned against
outputs of the *reference itself*, imported unchanged from /root/reference through ``oracle/ref_shim.py`` in the
authoring container; the resulting fixtures are committed under ``tests/golden/`` with the generating script
(``tests/golden/make_golden.py``).  The arithmetic of the reference lives in torch ATen (conv2d, batch_norm,
einsum, topk, softmax ...), which is why the restatement calls the same ATen ops on CPU.
"""
from .attention import (compute_affinity, masked_attention_efficient, propagate, spatial_neighbor)  # noqa: F401
from .head import cosine_sim_loss, simsiam_head_forward, simsiam_loss  # noqa: F401
from .resnet import resnet_forward, seeded_state_dict  # noqa: F401
from .siamfc import siam_conv_fc, xcorr  # noqa: F401
from .pipeline import normalize_format_ncthw, sample_train_augment, train_augment_ncthw  # noqa: F401
from .tracker import simsiam_forward_train, vanilla_forward_test  # noqa: F401
