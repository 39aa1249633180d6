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


# ---------------------------------------------------------------------------------------------------------------------
# Training feed: RandomResizedCrop -> Resize -> Flip -> Normalize -> FormatShape of the reference's train_pipeline
# (configs/*:48-92).  The random decisions (crop boxes, flips) are drawn on the host with the reference's exact sequence of
# numpy / random calls, so a seeded run reproduces the reference's augmentation; the pixel work is one CUDA kernel per
# batch (csrc/layout.cu: vfs_augment_u8_to_ncthw_f32 -- cv2-exact bilinear resize of the crop, mirror, normalise).
# ---------------------------------------------------------------------------------------------------------------------
def random_crop_bbox(img_shape, area_range=(0.08, 1.0), aspect_ratio_range=(3 / 4, 4 / 3), max_attempts=10):
    """RandomResizedCrop.get_crop_bbox (augmentations.py:214-262): same draws, same arithmetic."""
    import random
    assert 0 < area_range[0] <= area_range[1] <= 1
    assert 0 < aspect_ratio_range[0] <= aspect_ratio_range[1]
    img_h, img_w = img_shape
    area = img_h * img_w
    min_ar, max_ar = aspect_ratio_range
    ratios = np.exp(np.random.uniform(np.log(min_ar), np.log(max_ar), size=max_attempts))
    areas = np.random.uniform(*area_range, size=max_attempts) * area
    cand_w = np.round(np.sqrt(areas * ratios)).astype(np.int32)
    cand_h = np.round(np.sqrt(areas / ratios)).astype(np.int32)
    for crop_w, crop_h in zip(cand_w, cand_h):
        if crop_h <= img_h and crop_w <= img_w:
            x0 = random.randint(0, img_w - crop_w)
            y0 = random.randint(0, img_h - crop_h)
            return x0, y0, x0 + int(crop_w), y0 + int(crop_h)
    side = min(img_h, img_w)
    x0, y0 = (img_w - side) // 2, (img_h - side) // 2
    return x0, y0, x0 + side, y0 + side


class DeviceTrainAugment:
    """``RandomResizedCrop(area_range, aspect_ratio_range, same_on_clip, same_across_clip)`` +
    ``Resize(scale, keep_ratio=False)`` + ``Flip(flip_ratio, same_on_clip, same_across_clip)`` + ``Normalize`` +
    ``FormatShape('NCTHW')``: ``sample()`` draws the per-frame crop boxes and flip flags like the reference pipeline
    would for one video, ``__call__`` runs the pixels on the device."""

    def __init__(self, mean, std, to_bgr=False, scale=(224, 224), area_range=(0.08, 1.0),
                 aspect_ratio_range=(3 / 4, 4 / 3), flip_ratio=0.5, same_on_clip=True, same_across_clip=True,
                 device='cuda'):
        self.norm = DeviceNormalizeFormat(mean, std, to_bgr, device)
        self.scale = (int(scale[0]), int(scale[1]))            # (w, h) like mmcv
        self.area_range, self.aspect_ratio_range = tuple(area_range), tuple(aspect_ratio_range)
        self.flip_ratio, self.same_on_clip, self.same_across_clip = flip_ratio, same_on_clip, same_across_clip
        self.device = torch.device(device)

    def sample(self, img_shape, num_frames, clip_len):
        """Crop boxes [num_frames, 4] (x0, y0, x1, y1) and flip flags [num_frames] of one video's frames, consuming
        numpy's and random's global generators in the reference's order (all crops, then all flips)."""
        boxes = []
        box = random_crop_bbox(img_shape, self.area_range, self.aspect_ratio_range)
        for i in range(num_frames):
            is_new_clip = not self.same_across_clip and i % clip_len == 0 and i > 0
            if not self.same_on_clip or is_new_clip:
                box = random_crop_bbox(img_shape, self.area_range, self.aspect_ratio_range)
            boxes.append(box)
        flips = []
        flip = bool(np.random.rand() < self.flip_ratio)
        for i in range(num_frames):
            is_new_clip = not self.same_across_clip and i % clip_len == 0 and i > 0
            if not self.same_on_clip or is_new_clip:
                flip = bool(np.random.rand() < self.flip_ratio)
            flips.append(flip)
        return np.asarray(boxes, dtype=np.int32), np.asarray(flips, dtype=bool)

    def __call__(self, frames, boxes, flips, clip_len=1):
        """``frames``: uint8 tensor [F, H, W, 3] (pinned host or CUDA; F = clips * clip_len) or a list of F uint8
        [H_i, W_i, 3] tensors of different sizes; ``boxes`` [F, 4], ``flips`` [F] -> fp32 [clips, 3, clip_len, h, w]."""
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200.DeviceTrainAugment needs a CUDA device (no CPU fallback)')
        if torch.is_tensor(frames):
            frames = list(frames.to(self.device, non_blocking=True).contiguous().unbind(0))
        else:
            frames = [f.to(self.device, non_blocking=True).contiguous() for f in frames]
        F = len(frames)
        assert F % clip_len == 0 and len(boxes) == F and len(flips) == F
        items = (nat.VfsAugItem * F)()
        for i, (f, b, fl) in enumerate(zip(frames, boxes, flips)):
            if f.dtype != torch.uint8 or f.ndim != 3 or f.shape[2] != 3:
                raise TypeError('DeviceTrainAugment expects uint8 HWC frames')
            x0, y0, x1, y1 = (int(v) for v in b)
            if not (0 <= x0 < x1 <= f.shape[1] and 0 <= y0 < y1 <= f.shape[0]):
                raise ValueError(f'crop box {tuple(b)} outside a {f.shape[0]}x{f.shape[1]} frame')
            items[i] = nat.VfsAugItem(src=f.data_ptr(), H=f.shape[0], W=f.shape[1], crop_x0=x0, crop_y0=y0,
                                      crop_w=x1 - x0, crop_h=y1 - y0, flip=int(bool(fl)), reserved=0)
        table = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).pin_memory().to(self.device,
                                                                                             non_blocking=True)
        w, h = self.scale
        out = torch.empty((F // clip_len, 3, clip_len, h, w), dtype=torch.float32, device=self.device)
        check(nat.lib().vfs_augment_u8_to_ncthw_f32(ptr(table), ptr(out), F // clip_len, clip_len, h, w,
                                                    self.norm._mean3, self.norm._stdinv3, int(self.norm.to_bgr),
                                                    current_stream()), 'augment_u8_to_ncthw_f32')
        self._keep = (frames, table)      # the kernel reads them asynchronously
        return out


class PinnedRing:
    """Double-buffered pinned host -> device feed: ``put(host_tensor)`` copies the tensor into the next pinned slot and
    starts its H2D copy on a side stream; ``get()`` returns the oldest device tensor after making the current stream
    wait for its copy.  Calling ``put`` for batch i+1 right after ``get`` of batch i -- before enqueuing batch i's
    kernels -- overlaps the PCIe transfer of i+1 with the compute of i (the reference relies on
    DataLoader(pin_memory=True) + scatter: one blocking copy per step).

    Slot life cycle: the pinned buffer may be rewritten once its own H2D has finished (``h2d`` event, host wait); the
    device buffer may be overwritten once the kernels that consumed it have finished (``free`` event, recorded on the
    consumer stream at the NEXT ``get`` -- by then the consumers of the previously returned slot have been enqueued --
    and waited for on the copy stream, not on the host)."""

    def __init__(self, slots=2, device='cuda'):
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200.PinnedRing needs a CUDA device')
        if slots < 2:
            raise ValueError('PinnedRing needs at least two slots')
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(host=None, dev=None, h2d=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(slots)]
        self._head = self._tail = self._count = 0
        self._last = None

    def put(self, t):
        if self._count == len(self.slots):
            raise RuntimeError('PinnedRing is full: call get() first')
        s = self.slots[self._head]
        fresh = s['dev'] is None or s['dev'].shape != t.shape or s['dev'].dtype != t.dtype
        if fresh:
            s['host'] = None
        src = t
        if not t.is_pinned():
            # pageable source: stage through this slot's pinned buffer (a loader that decodes straight into pinned
            # memory -- DataLoader(pin_memory=True) -- skips this copy)
            if s['host'] is None:
                s['host'] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            s['h2d'].synchronize()           # the previous copy out of this pinned buffer has finished
            s['host'].copy_(t)
            src = s['host']
        self.stream.wait_event(s['free'])    # the kernels that read the old device contents are done
        with torch.cuda.stream(self.stream):
            if fresh:
                # allocated under the copy stream: the caching allocator orders block reuse per stream, and a block of
                # the caller's stream may still be referenced by kernels queued there when the copy stream writes it
                s['dev'] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            s['dev'].copy_(src, non_blocking=True)
            s['h2d'].record(self.stream)
        s['src'] = src                       # keep the pinned source alive until its copy has run
        self._head = (self._head + 1) % len(self.slots)
        self._count += 1

    def get(self):
        if self._count == 0:
            raise RuntimeError('PinnedRing is empty')
        cur = torch.cuda.current_stream(self.device)
        if self._last is not None:
            self._last['free'].record(cur)   # everything enqueued since the last get() consumed that slot
        s = self.slots[self._tail]
        cur.wait_event(s['h2d'])
        s['dev'].record_stream(cur)          # consumed on the caller's stream: keep the block until that work is done
        self._tail = (self._tail + 1) % len(self.slots)
        self._count -= 1
        self._last = s
        return s['dev']
