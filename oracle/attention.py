"""ORACLE -- test infrastructure only (see oracle/__init__.py).

Restatement of the DAVIS label-propagation arithmetic:
    masked_attention_efficient   mmaction/models/common/local_attention.py:237-348
    spatial_neighbor             mmaction/models/common/affinity_utils.py:119-156
    compute_affinity             affinity_utils.py:6-30
    propagate                    affinity_utils.py:33-50
"""
import torch
import torch.nn.functional as F


def spatial_neighbor(height, width, neighbor_range, mode='circle', batches=1):
    """circle: bool [HW, HW], mask[(y,x),(y',x')] = dist < neighbor_range // 2 (affinity_utils.py:145-156);
    square: bool [batches, HW, HW], |dy| <= r//2 and |dx| <= r//2 (:130-144; note the extra batch dim, which
    makes masked_attention_efficient take its ``mask.ndim != 2`` branch that requires clip_len == 1)."""
    radius = neighbor_range // 2
    ys = torch.arange(height, dtype=torch.float32).view(height, 1, 1, 1)
    xs = torch.arange(width, dtype=torch.float32).view(1, width, 1, 1)
    dy = ys - torch.arange(height, dtype=torch.float32).view(1, 1, height, 1)
    dx = xs - torch.arange(width, dtype=torch.float32).view(1, 1, 1, width)
    if mode == 'circle':
        mask = (dy**2 + dx**2)**0.5 < radius
    else:
        mask = (dy.abs() <= radius) & (dx.abs() <= radius)
        mask = mask.expand(height, width, height, width)
        return mask.reshape(1, height * width, height * width).expand(batches, -1, -1).bool()
    return mask.reshape(height * width, height * width).bool()


def masked_attention_efficient(query, key, value, mask, temperature=1, topk=None, normalize=True, step=32,
                               non_mask_len=0, mode='softmax', return_topk=False):
    """query [N,C,H,W], key [N,C,T,H,W], value [N,Cv,T,H,W], mask bool [HW_key, HW_query] (or None).
    Follows the reference chunk loop (local_attention.py:287-342): per ``step`` query columns
    A = (K^T Q)/temperature, A[~mask] = -inf, top-k over keys, softmax over the k, weighted sum of values.
    ``return_topk`` additionally returns (values [N,k,HW], indices [N,k,HW]) for index-parity tests."""
    assert mode in ('softmax', 'cosine')
    batches = query.size(0)
    if key.ndim == 4:
        key, value = key.unsqueeze(2), value.unsqueeze(2)
    clip_len = key.size(2)
    assert 0 <= non_mask_len < clip_len
    att_channels, qh, qw = query.shape[1:]
    kh, kw = key.shape[3:]
    out_channels = value.size(1)
    if normalize:
        query = F.normalize(query, p=2, dim=1)
        key = F.normalize(key, p=2, dim=1)
    q = query.reshape(batches, att_channels, qh * qw)
    k = key.reshape(batches, att_channels, clip_len * kh * kw)
    v = value.reshape(batches, out_channels, clip_len * kh * kw)
    output = torch.zeros(batches, out_channels, qh * qw, dtype=query.dtype)
    all_val, all_idx = [], []
    if step is None:
        step = qh * qw
    for ptr in range(0, qh * qw, step):
        aff = torch.einsum('bci,bcj->bij', k, q[..., ptr:ptr + step]) / temperature
        if mask is not None and mask.ndim != 2:
            assert clip_len == 1 and non_mask_len == 0  # local_attention.py:303-305
            aff.masked_fill_(~mask[..., ptr:ptr + step].bool(), float('-inf'))
        elif mask is not None:
            assert mask.shape == (kh * kw, qh * qw)
            cur = mask.view(1, 1, kh * kw, qh * qw)[..., ptr:ptr + step].expand(
                batches, clip_len - non_mask_len, -1, -1).reshape(batches, -1, aff.size(2))
            if non_mask_len > 0:
                cur = torch.cat([torch.ones(batches, non_mask_len * kh * kw, aff.size(2), dtype=cur.dtype), cur], 1)
            aff.masked_fill_(~cur.bool(), float('-inf'))
        if topk is not None:
            val, idx = aff.topk(k=topk, dim=1)
            tv = v.transpose(0, 1).reshape(out_channels, -1).index_select(1, idx.reshape(-1))
            tv = tv.reshape(out_channels, *idx.shape).transpose(0, 1)
            w = val.softmax(dim=1) if mode == 'softmax' else val.clamp(min=0)**2
            cur_out = torch.einsum('bcks,bks->bcs', tv, w)
            all_val.append(val)
            all_idx.append(idx)
        else:
            w = aff.softmax(dim=1) if mode == 'softmax' else aff.clamp(min=0)**2
            cur_out = torch.einsum('bck,bks->bcs', v, w)
        output[..., ptr:ptr + step] = cur_out
    output = output.reshape(batches, out_channels, qh, qw)
    if return_topk:
        return output, torch.cat(all_val, dim=2), torch.cat(all_idx, dim=2)
    return output


def compute_affinity(src_img, dst_img, temperature=1., normalize=True, softmax_dim=None, mask=None):
    """Dense [B, HW_src, HW_dst] affinity (affinity_utils.py:6-30)."""
    batches, channels = src_img.shape[:2]
    src = src_img.reshape(batches, channels, -1)
    dst = dst_img.reshape(batches, channels, -1)
    if normalize:
        src = F.normalize(src, p=2, dim=1)
        dst = F.normalize(dst, p=2, dim=1)
    affinity = torch.bmm(src.permute(0, 2, 1).contiguous(), dst.contiguous()) / temperature
    if mask is not None:
        affinity.masked_fill_(~mask.bool(), float('-inf'))
    if softmax_dim is not None:
        affinity = affinity.softmax(dim=softmax_dim)
    if mask is not None:
        affinity[affinity.isnan()] = 0
    return affinity


def propagate(img, affinity, topk=None):
    """new_img = img @ affinity, optionally thresholded at the k-th value per column and re-normalised
    (affinity_utils.py:33-50).  Does not modify ``affinity`` (the reference works in place)."""
    batches, channels, height, width = img.size()
    affinity = affinity.clone()
    if topk is not None:
        kth = affinity.topk(dim=1, k=topk)[0][:, topk - 1].view(batches, 1, height * width)
        affinity -= kth
        affinity.clamp_(min=0)
        affinity /= affinity.sum(keepdim=True, dim=1).clamp(min=1e-12)
    new_img = torch.bmm(img.reshape(batches, channels, -1).contiguous(), affinity.contiguous())
    return new_img.reshape(batches, channels, height, width)
