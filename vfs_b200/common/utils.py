"""Layout helpers of the hot path (reference mmaction/models/common/utils.py)."""
from typing import List

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair, _single, _triple


def change_stride(conv, stride):
    """In-place stride change of a conv layer (reference utils.py:10-22)."""
    if isinstance(conv, nn.Conv1d):
        conv.stride = _single(stride)
    if isinstance(conv, nn.Conv2d):
        conv.stride = _pair(stride)
    if isinstance(conv, nn.Conv3d):
        conv.stride = _triple(stride)


def pil_nearest_index_map(src_size, dst_size):
    """Source index picked by Pillow's NEAREST resize along one axis: floor((i + 0.5) * src / dst).

    Pillow (Resample.c, ImagingTransform nearest filter on an axis-aligned box) samples the source at the
    centre of each destination pixel.  Used to reproduce ``mmcv.imresize(..., 'nearest', backend='pillow')``
    (reference utils.py:25-42) without a host round trip."""
    i = np.arange(dst_size, dtype=np.float64)
    idx = np.floor((i + 0.5) * (float(src_size) / float(dst_size))).astype(np.int64)
    return np.clip(idx, 0, src_size - 1)


_INDEX_MAP_CACHE = {}


def _device_index_map(src_size, dst_size, device):
    """Index map of one axis as a device tensor, built once per (sizes, device): a pageable host->device copy per
    call would synchronise the stream in the middle of the tracker's per-video loop."""
    key = (int(src_size), int(dst_size), str(device))
    t = _INDEX_MAP_CACHE.get(key)
    if t is None:
        t = torch.from_numpy(pil_nearest_index_map(src_size, dst_size)).to(device)
        _INDEX_MAP_CACHE[key] = t
    return t


def pil_nearest_interpolate(input, size):
    """Nearest resize with Pillow semantics; ``input`` [N,1,H,W], ``size`` (h, w) -> [N,1,h,w]."""
    assert input.ndim == 4 and input.size(1) == 1
    rows = _device_index_map(input.size(2), size[0], input.device)
    cols = _device_index_map(input.size(3), size[1], input.device)
    return input.index_select(2, rows).index_select(3, cols)


def video2images(imgs):
    """[B,C,T,H,W] -> [B*T,C,H,W] (reference utils.py:45-53)."""
    batches, channels, clip_len = imgs.shape[:3]
    if clip_len == 1:
        return imgs.squeeze(2).reshape(batches, channels, *imgs.shape[3:])
    return imgs.transpose(1, 2).contiguous().reshape(batches * clip_len, channels, *imgs.shape[3:])


def images2video(imgs, clip_len):
    """[B*T,C,...] -> [B,C,T,...] (reference utils.py:56-64)."""
    batches, channels = imgs.shape[:2]
    if clip_len == 1:
        return imgs.unsqueeze(2)
    return imgs.reshape(batches // clip_len, clip_len, channels, *imgs.shape[2:]).transpose(1, 2).contiguous()


class StrideContext(object):
    """``with StrideContext(backbone, strides, out_indices):`` temporary stride/out-index switch
    (reference utils.py:84-101)."""

    def __init__(self, backbone, strides, out_indices=None):
        self.backbone, self.strides, self.out_indices = backbone, strides, out_indices

    def __enter__(self):
        if self.strides is not None:
            self.backbone.switch_strides(self.strides)
        if self.out_indices is not None:
            self.backbone.switch_out_indices(self.out_indices)

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.strides is not None:
            self.backbone.switch_strides()
        if self.out_indices is not None:
            self.backbone.switch_out_indices()


def cat(tensors: List[torch.Tensor], dim: int = 0):
    """``torch.cat`` that returns the single element un-copied (reference utils.py:188-194)."""
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def add_prefix(inputs, prefix):
    """``{k: v} -> {prefix.k: v}`` (reference mmaction/utils/misc.py:45-61)."""
    return {f'{prefix}.{name}': value for name, value in inputs.items()}
