from .affinity_utils import NeighborMask, compute_affinity, propagate, spatial_neighbor
from .local_attention import masked_attention_efficient
from .utils import (StrideContext, add_prefix, cat, change_stride, images2video, pil_nearest_interpolate,
                    video2images)

__all__ = ['change_stride', 'pil_nearest_interpolate', 'compute_affinity', 'propagate', 'images2video',
           'video2images', 'spatial_neighbor', 'StrideContext', 'cat', 'masked_attention_efficient',
           'NeighborMask', 'add_prefix']
