"""Affinity helpers behind the reference's names (mmaction/models/common/affinity_utils.py).

``spatial_neighbor`` does NOT materialise the reference's O(HW^2) bool matrix (41 MB at 480p): it returns a
``NeighborMask`` describing the window analytically, which the fused attention kernel evaluates per (query, key)
pair and uses to skip key tiles outside the window (csrc/affinity.cu).  ``NeighborMask.dense()`` rebuilds the
reference's tensor for callers/tests that want it."""
import torch
from torch.nn.modules.utils import _pair

from .. import ops


class NeighborMask:
    """mask[(y,x) key, (y',x') query] = inside the window centred on the query (circle: dist < radius;
    square: |dy| <= ry and |dx| <= rx).  Symmetric, so key-major vs query-major does not matter."""

    def __init__(self, batches, height, width, neighbor_range, mode='circle', dim=1):
        assert dim in [1, 2]
        assert mode in ['circle', 'square']
        self.batches, self.height, self.width, self.mode, self.dim = batches, height, width, mode, dim
        if mode == 'square':
            rng = _pair(neighbor_range)
            self.radius_y, self.radius_x = rng[0] // 2, rng[1] // 2
        else:
            self.radius_y = self.radius_x = neighbor_range // 2

    @property
    def ndim(self):
        return 3 if self.mode == 'square' else 2

    @property
    def shape(self):
        hw = self.height * self.width
        return (self.batches, hw, hw) if self.mode == 'square' else (hw, hw)

    def dense(self, device='cpu'):
        """The reference's bool tensor ([HW,HW] for circle, [B,HW,HW] for square)."""
        h, w = self.height, self.width
        ys = torch.arange(h, device=device).view(h, 1, 1, 1)
        xs = torch.arange(w, device=device).view(1, w, 1, 1)
        dy = ys - torch.arange(h, device=device).view(1, 1, h, 1)
        dx = xs - torch.arange(w, device=device).view(1, 1, 1, w)
        if self.mode == 'circle':
            m = (dy * dy + dx * dx) < self.radius_y * self.radius_y
            return m.reshape(h * w, h * w)
        m = (dy.abs() <= self.radius_y) & (dx.abs() <= self.radius_x)
        return m.reshape(1, h * w, h * w).expand(self.batches, -1, -1)

    def bool(self):
        return self


def spatial_neighbor(batches, height, width, neighbor_range, device=None, dtype=None, dim=1, mode='circle'):
    """Signature of affinity_utils.py:119-126; ``device``/``dtype`` are accepted and unused (nothing is
    materialised)."""
    return NeighborMask(batches, height, width, neighbor_range, mode=mode, dim=dim)


def compute_affinity(src_img, dst_img, temperature=1., normalize=True, softmax_dim=None, mask=None):
    """Dense [B, HW_src, HW_dst] affinity (affinity_utils.py:6-30): tcgen05 GEMM of the (L2-normalised) features
    scaled by 1/temperature, optional mask fill and softmax along ``softmax_dim``."""
    return ops.dense_affinity(src_img, dst_img, temperature, normalize, softmax_dim, mask)


def propagate(img, affinity, topk=None):
    """``img @ affinity`` with optional k-th-value thresholding and re-normalisation per column
    (affinity_utils.py:33-50).  Unlike the reference this does not modify ``affinity`` in place."""
    return ops.propagate_dense(img, affinity, topk)
