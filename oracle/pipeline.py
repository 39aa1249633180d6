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


# ---------------------------------------------------------------------------------------------------------------------
# Training pipeline of the configs (configs/*:48-92): RandomResizedCrop -> Resize(keep_ratio=False) -> Flip -> Normalize
# -> FormatShape('NCTHW').
#     RandomResizedCrop.get_crop_bbox / __call__   augmentations.py:214-262, 264-306
#     Resize.__call__ (keep_ratio=False)            augmentations.py:487-597 -> mmcv.imresize = cv2.resize(INTER_LINEAR)
#     Flip.__call__                                 augmentations.py:600-711 -> mmcv.imflip_ = cv2.flip(img, 1, img)
# PINNED: oracle/ref_shim.py::load_reference_pipelines imports the reference's pipeline classes unchanged (mmcv's three
# image helpers restated with the cv2 calls mmcv makes); tests/golden/train_pipeline_golden.npz holds their output for
# seeded runs and tests/test_oracle_golden.py checks this restatement against it bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
def random_crop_bbox(img_shape, area_range, aspect_ratio_range, max_attempts=10):
    import random
    img_h, img_w = img_shape
    area = img_h * img_w
    min_ar, max_ar = aspect_ratio_range
    aspect_ratios = np.exp(np.random.uniform(np.log(min_ar), np.log(max_ar), size=max_attempts))
    target_areas = np.random.uniform(*area_range, size=max_attempts) * area
    candidate_crop_w = np.round(np.sqrt(target_areas * aspect_ratios)).astype(np.int32)
    candidate_crop_h = np.round(np.sqrt(target_areas / aspect_ratios)).astype(np.int32)
    for i in range(max_attempts):
        crop_w, crop_h = candidate_crop_w[i], candidate_crop_h[i]
        if crop_h <= img_h and crop_w <= img_w:
            x_offset = random.randint(0, img_w - crop_w)
            y_offset = random.randint(0, img_h - crop_h)
            return x_offset, y_offset, x_offset + crop_w, y_offset + crop_h
    crop_size = min(img_h, img_w)
    x_offset, y_offset = (img_w - crop_size) // 2, (img_h - crop_size) // 2
    return x_offset, y_offset, x_offset + crop_size, y_offset + crop_size


def sample_train_augment(img_shape, num_frames, clip_len, area_range, aspect_ratio_range, flip_ratio, same_on_clip,
                         same_across_clip):
    """Crop boxes and flip flags in the order the two pipeline steps consume the global RNGs."""
    boxes = []
    box = random_crop_bbox(img_shape, area_range, aspect_ratio_range)
    for i in range(num_frames):
        is_new_clip = not same_across_clip and i % clip_len == 0 and i > 0
        if not same_on_clip or is_new_clip:
            box = random_crop_bbox(img_shape, area_range, aspect_ratio_range)
        boxes.append(tuple(int(v) for v in box))
    flips = []
    flip = bool(np.random.rand() < flip_ratio)
    for i in range(num_frames):
        is_new_clip = not same_across_clip and i % clip_len == 0 and i > 0
        if not same_on_clip or is_new_clip:
            flip = bool(np.random.rand() < flip_ratio)
        flips.append(flip)
    return boxes, flips


def train_augment_ncthw(frames_u8, boxes, flips, scale, mean, std, to_bgr=False, num_clips=1):
    """frames: list of uint8 [H,W,3]; boxes (x0,y0,x1,y1); scale (w, h) -> float32 [num_clips, 3, clip_len, h, w]."""
    out = []
    for img, (x0, y0, x1, y1), flip in zip(frames_u8, boxes, flips):
        crop = img[y0:y1, x0:x1]
        res = cv2.resize(crop, tuple(scale), dst=None, interpolation=cv2.INTER_LINEAR)
        if flip:
            res = np.ascontiguousarray(res)
            cv2.flip(res, 1, res)
        out.append(res)
    return normalize_format_ncthw(np.stack(out), mean, std, to_bgr, num_clips)
