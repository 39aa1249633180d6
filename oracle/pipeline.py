"""ORACLE -- test infrastructure only (see oracle/__init__.py).

    Normalize.__call__            mmaction/datasets/pipelines/augmentations.py:711-757
    mmcv.imnormalize_             mmcv-full 1.2.1 (mmcv/image/photometric.py; third-party, absent here): float64 mean and
                                  1/std row vectors, optional cv2.cvtColor(BGR2RGB), cv2.subtract, cv2.multiply in place
    FormatShape('NCTHW')          mmaction/datasets/pipelines/formating.py:248-258

Restated with the same cv2 calls mmcv makes; mmcv itself is not installed, so this piece is "parity unpinned"
(anchored on the reference's call sites above and on cv2, which IS the arithmetic).
"""
import cv2
import numpy as np


def normalize_format_ncthw(frames_u8, mean, std, to_bgr=False, num_clips=1):
    """frames uint8 [M, H, W, 3] with M = num_clips * clip_len -> float32 [num_clips, 3, clip_len, H, W]."""
    mean = np.array(mean, dtype=np.float32)
    std = np.array(std, dtype=np.float32)
    n, h, w, c = frames_u8.shape
    imgs = np.empty((n, h, w, c), dtype=np.float32)
    for i, img in enumerate(frames_u8):
        imgs[i] = img
    m64 = np.float64(mean.reshape(1, -1))
    stdinv = 1 / np.float64(std.reshape(1, -1))
    for img in imgs:
        if to_bgr:
            cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
        cv2.subtract(img, m64, img)
        cv2.multiply(img, stdinv, img)
    clip_len = n // num_clips
    imgs = imgs.reshape((-1, num_clips, clip_len) + imgs.shape[1:])
    imgs = np.transpose(imgs, (0, 1, 5, 2, 3, 4))
    return imgs.reshape((-1, ) + imgs.shape[2:])
