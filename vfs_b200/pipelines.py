"""Device-side data feed: the last two CPU steps of the reference's loading pipeline --
``Normalize(mean, std, to_bgr)`` (mmaction/datasets/pipelines/augmentations.py:711-757) and ``FormatShape('NCTHW')``
(formating.py:248-258) -- as one CUDA kernel, so that decoded frames travel host -> device as uint8 HWC (a quarter of
the bytes of the fp32 NCTHW tensor the reference ships) and the fp32 clip tensor is produced in HBM.

    feed = DeviceNormalizeFormat(**cfg.img_norm_cfg)          # mean, std, to_bgr of configs/*:85-86
    imgs = feed(frames_u8)                                    # uint8 [B, T, H, W, 3] (pinned host or CUDA) -> fp32 [B, 3, T, H, W]
"""
import ctypes

import numpy as np
import torch

from . import _native as nat
from .ops import check, current_stream, ptr


class DeviceNormalizeFormat:

    def __init__(self, mean, std, to_bgr=False, device='cuda'):
        if not isinstance(mean, (list, tuple, np.ndarray)):
            raise TypeError(f'Mean must be list, tuple or np.ndarray, but got {type(mean)}')
        if not isinstance(std, (list, tuple, np.ndarray)):
            raise TypeError(f'Std must be list, tuple or np.ndarray, but got {type(std)}')
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)
        assert self.mean.shape == self.std.shape == (3, ), 'RGB frames: three channel statistics'
        self.to_bgr = bool(to_bgr)
        self.device = torch.device(device)
        # mmcv.imnormalize_: mean and 1/std as float64 row vectors; cv2.subtract rounds to fp32, cv2.multiply by the
        # float64 scalar multiplies in fp64 and rounds once (csrc/layout.cu does the same)
        self._mean3 = (ctypes.c_float * 3)(*[float(m) for m in self.mean])
        self._stdinv3 = (ctypes.c_double * 3)(*[float(1 / np.float64(s)) for s in self.std])

    def __call__(self, frames):
        """``frames`` uint8 [B, T, H, W, 3] (or [T, H, W, 3] = one clip) -> fp32 [B, 3, T, H, W] on the device."""
        if frames.dtype != torch.uint8:
            raise TypeError('DeviceNormalizeFormat expects uint8 frames (the decoder output)')
        if frames.ndim == 4:
            frames = frames.unsqueeze(0)
        assert frames.ndim == 5 and frames.shape[-1] == 3, 'expected [B, T, H, W, 3]'
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200.DeviceNormalizeFormat needs a CUDA device (no CPU fallback)')
        dev = frames.to(self.device, non_blocking=True).contiguous()
        B, T, H, W, _ = dev.shape
        out = torch.empty((B, 3, T, H, W), dtype=torch.float32, device=dev.device)
        check(nat.lib().vfs_frames_u8_to_ncthw_f32(ptr(dev), ptr(out), B, T, H, W, self._mean3, self._stdinv3,
                                                   int(self.to_bgr), current_stream()), 'frames_u8_to_ncthw_f32')
        return out
