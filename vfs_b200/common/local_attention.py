"""Restricted-attention label propagation behind the reference's name and signature
(mmaction/models/common/local_attention.py:237-348), executed by the fused tcgen05 affinity / top-k kernel
(csrc/affinity.cu) instead of the reference's 201-chunk einsum/topk/softmax loop: the HW x T*HW affinity never
touches HBM."""
from .. import ops
from .affinity_utils import NeighborMask


def masked_attention_efficient(query, key, value, mask, temperature=1, topk=None, normalize=True, step=32,
                               non_mask_len=0, mode='softmax'):
    """query [N,C,H,W], key [N,C,T,H,W] (or [N,C,H,W]), value [N,Cv,T,H,W]; ``mask`` is None, the ``NeighborMask``
    returned by ``spatial_neighbor`` (fused window kernel) or any boolean tensor [HWk,HWq] / [N,HWk,HWq] (general
    kernel).  ``step`` (the reference's memory chunking) is accepted and ignored.  Returns [N,Cv,H,W] fp32."""
    assert mode in ['softmax', 'cosine']
    assert query.size(0) == key.size(0) == value.size(0)
    assert value.shape[2:] == key.shape[2:], f'{value.shape} {key.shape}'
    if key.ndim == 4:
        key = key.unsqueeze(2)
        value = value.unsqueeze(2)
    assert value.ndim == key.ndim == 5
    clip_len = key.size(2)
    assert 0 <= non_mask_len < clip_len
    window = mask is None or isinstance(mask, NeighborMask)
    if mask is not None:
        if mask.ndim == 2:
            assert tuple(mask.shape) == (key.shape[3] * key.shape[4], query.shape[2] * query.shape[3])
        else:
            assert clip_len == 1
            assert non_mask_len == 0
    if window and topk is not None and query.shape[2:] == key.shape[3:]:
        return ops.masked_attention(query, key, value, mask, temperature, topk, normalize, non_mask_len, mode)
    # arbitrary boolean mask tensors, topk=None (softmax over every key) and query / key maps of different size:
    # the affinity is materialised by the tcgen05 GEMM and finished by the general kernel (csrc/dense.cu)
    if isinstance(mask, NeighborMask):
        mask = mask.dense(device=query.device)
    return ops.masked_attention_generic(query, key, value, mask, temperature, topk, normalize, non_mask_len, mode)
