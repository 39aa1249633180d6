"""SiamFC object tracker: interface of projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py (``Net`` :75-85,
``TrackerSiamFC.init`` :200-241, ``update`` :243-319, ``track`` :321-349) on the B200 path.

Per frame the reference crops three scaled search windows on the host (cv2), runs backbone + head on the GPU, moves
the three 17 x 17 response maps back to the host and finishes with cv2.resize / numpy.  Here the crops still come from
cv2 (data-format side, a few hundred KB), but everything after the host->device copy stays on the device: input
normalisation, the tcgen05 backbone, the 1x1 adapters + cross-correlation head, and the response post-processing
(bicubic upsample, scale penalty, Hann window, arg-max -- csrc/post.cu); 12 bytes come back per frame.  The got10k
``Tracker`` base class of the reference only supplies the benchmark loop and is not needed for the path."""
import time

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..builder import build_backbone
from ..mmcv_lite import ConfigDict
from . import image_ops
from .heads import SiamConvFC, SiamFC
from .train import TrainingMixin

# projects/siamfc-pytorch/siamfc/default_config_base.py:2-51 (inference-relevant keys)
DEFAULT_CFG = dict(
    out_scale=0.001, exemplar_sz=120, instance_sz=255, context=0.5, scale_num=3, scale_step=1.0375, scale_lr=0.59,
    scale_penalty=0.9745, window_influence=0.176, response_sz=17, response_up=16, total_stride=8, extra_conv=True,
    out_channels=512, reduction=1, out_block_index=None,
    # training keys (default_config_base.py:18-38)
    epoch_num=50, batch_size=8, initial_lr=1e-3, ultimate_lr=1e-5, weight_decay=5e-4, momentum=0.9, r_pos=16, r_neg=0,
    optimizer='Adam', loss='focal', lr_schedule='exp', lr_step_size=10, force_wd=False,
    model=dict(backbone=dict(frozen_stages=4, dilations=(1, 1, 2, 4), strides=(1, 2, 1, 1), out_indices=(3, ),
                             with_cp=False, norm_eval=True)))

_MEAN = (123.675, 116.28, 103.53)
_STD = (58.395, 57.12, 57.375)


class Net(nn.Module):
    """backbone on both branches, then the correlation head (siamfc_tracker_base.py:75-85)."""

    def __init__(self, backbone, head):
        super().__init__()
        self.backbone = backbone
        self.head = head

    def forward(self, z, x):
        return self.head(self.backbone(z), self.backbone(x))


def build_cfg(backbone, **overrides):
    """default_cfg merged with a backbone config dict (``type='ResNet', depth=...``) and keyword overrides."""
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in DEFAULT_CFG.items()}
    model_backbone = dict(DEFAULT_CFG['model']['backbone'])
    model_backbone.update(backbone)
    cfg['model'] = dict(backbone=model_backbone)
    cfg.update(overrides)
    return ConfigDict(cfg)


class TrackerSiamFC(TrainingMixin):
    """``init(img, box)`` / ``update(img)`` / ``track(frames, box)`` with the reference's state variables
    (``center``, ``target_sz``, ``z_sz``, ``x_sz``, ``scale_factors``, ``hann_window``, ``kernel``).  ``img`` is an
    RGB uint8 array [H,W,3]; ``box`` is 1-indexed (x, y, w, h) like the OTB / GOT-10k annotations."""

    def __init__(self, cfg, device='cuda'):
        self.cfg = cfg if isinstance(cfg, ConfigDict) else ConfigDict(cfg)
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200 TrackerSiamFC needs a CUDA device (no CPU fallback)')
        self.device = torch.device(device)
        backbone = build_backbone(dict(self.cfg.model.backbone))
        backbone.init_weights()
        if self.cfg.out_block_index is not None:
            index = self.cfg.out_block_index
            forward_block = backbone.forward_block
            backbone.forward = lambda x: forward_block(x, index=index)      # siamfc_tracker_base.py:104-109
        if self.cfg.extra_conv:
            head = SiamConvFC(self.cfg.out_channels, self.cfg.out_channels // self.cfg.reduction,
                              out_scale=self.cfg.out_scale)
        else:
            head = SiamFC(out_scale=self.cfg.out_scale)
        self.net = Net(backbone, head).to(self.device)
        self._mean = torch.tensor(_MEAN, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
        self._std = torch.tensor(_STD, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)

    # ------------------------------------------------------------------ helpers
    def normalize(self, x):
        """torchvision Normalize(mean, std) on a [N,3,H,W] batch (siamfc_tracker_base.py:168-169)."""
        return (x - self._mean) / self._std

    def _to_device(self, crops):
        """uint8 / float HWC crops [N,h,w,3] -> normalised fp32 NCHW on the device (one pinned copy)."""
        host = torch.from_numpy(np.ascontiguousarray(crops)).pin_memory()
        x = host.to(self.device, non_blocking=True).permute(0, 3, 1, 2).float()
        return self.normalize(x)

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def init(self, img, box):
        self.net.eval()
        cfg = self.cfg
        box = np.array([box[1] - 1 + (box[3] - 1) / 2, box[0] - 1 + (box[2] - 1) / 2, box[3], box[2]],
                       dtype=np.float32)
        self.center, self.target_sz = box[:2], box[2:]
        self.upscale_sz = cfg.response_up * cfg.response_sz
        self.hann_window = np.outer(np.hanning(self.upscale_sz), np.hanning(self.upscale_sz))
        self.hann_window /= self.hann_window.sum()
        self._hann_dev = torch.from_numpy(self.hann_window).to(self.device)
        self.scale_factors = cfg.scale_step**np.linspace(-(cfg.scale_num // 2), cfg.scale_num // 2, cfg.scale_num)
        context = cfg.context * np.sum(self.target_sz)
        self.z_sz = np.sqrt(np.prod(self.target_sz + context))
        self.x_sz = self.z_sz * cfg.instance_sz / cfg.exemplar_sz
        self.avg_color = np.mean(img, axis=(0, 1))
        z = image_ops.crop_and_resize(img, self.center, self.z_sz, out_size=cfg.exemplar_sz,
                                      border_value=self.avg_color)
        self.kernel = self.net.backbone(self._to_device(z[None]))

    @torch.no_grad()
    def responses(self, img):
        """Backbone + head on the scaled search windows of ``img`` -> [scale_num, R, R] on the device."""
        cfg = self.cfg
        mean = image_ops.mean_colour(img)          # once per frame (the reference recomputes it inside every crop)
        x = np.stack([image_ops.crop_and_resize(img, self.center, self.x_sz * f, out_size=cfg.instance_sz,
                                                border_value=self.avg_color, mean=mean)
                      for f in self.scale_factors], axis=0)
        feats = self.net.backbone(self._to_device(x))
        return self.net.head(self.kernel, feats).squeeze(1)

    @torch.no_grad()
    def update(self, img):
        if self.net.training:                      # (module.eval() walks every sub-module: 1.6 ms per frame)
            self.net.eval()
        cfg = self.cfg
        responses = self.responses(img)
        peak = ops.siamfc_response_peak(responses, self._hann_dev, self.upscale_sz, cfg.scale_penalty,
                                        cfg.window_influence).cpu().numpy()           # the frame's only D2H
        scale_id, loc = int(peak[0]), (int(peak[1]), int(peak[2]))
        return self._apply_peak(scale_id, loc)

    def _apply_peak(self, scale_id, loc):
        """State update from the peak (siamfc_tracker_base.py:293-319)."""
        cfg = self.cfg
        disp_in_response = np.array(loc) - (self.upscale_sz - 1) / 2
        disp_in_instance = disp_in_response * cfg.total_stride / cfg.response_up
        disp_in_image = disp_in_instance * self.x_sz * self.scale_factors[scale_id] / cfg.instance_sz
        self.center += disp_in_image
        scale = (1 - cfg.scale_lr) * 1.0 + cfg.scale_lr * self.scale_factors[scale_id]
        self.target_sz *= scale
        self.z_sz *= scale
        self.x_sz *= scale
        return np.array([self.center[1] + 1 - (self.target_sz[1] - 1) / 2,
                         self.center[0] + 1 - (self.target_sz[0] - 1) / 2, self.target_sz[1], self.target_sz[0]])

    def track(self, frames, box):
        """``frames``: iterable of RGB uint8 arrays (the reference reads files with cv2 and converts BGR->RGB)."""
        frames = list(frames)
        boxes = np.zeros((len(frames), 4))
        boxes[0] = box
        times = np.zeros(len(frames))
        for f, img in enumerate(frames):
            begin = time.time()
            if f == 0:
                self.init(img, box)
            else:
                boxes[f, :] = self.update(img)
            times[f] = time.time() - begin
        return boxes, times
