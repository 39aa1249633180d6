"""SimSiam pre-training wrapper behind the reference's class name and config keys
(mmaction/models/trackers/sim_siam_base_tracker.py).

Input contract: ``imgs`` [B, 2 views, C, T, H, W].  Each view is a separate backbone pass -- and therefore a separate
BatchNorm batch, which matters for parity -- followed by the shared projector/predictor head and the symmetric
stop-gradient cosine loss.  With ``train_cfg.intra_video`` the loss is additionally evaluated against the second
view rolled along the clip axis (every other frame of the same video is a positive), each term weighted 1/T.
"""
import torch

from .. import builder
from ..common import add_prefix, images2video, video2images
from ..registry import TRACKERS
from .base import BaseTracker


@TRACKERS.register_module()
class SimSiamBaseTracker(BaseTracker):

    #: train_cfg switches and their defaults when the key (or the whole train_cfg) is absent
    _TRAIN_FLAGS = dict(intra_video=False, transpose_temporal=False)

    def __init__(self, *args, backbone, img_head=None, **kwargs):
        BaseTracker.__init__(self, *args, backbone=backbone, **kwargs)
        self.img_head = None if img_head is None else builder.build_head(img_head)
        self.init_extra_weights()
        flags = self.train_cfg or {}
        for key, default in self._TRAIN_FLAGS.items():
            setattr(self, key, flags.get(key, default))

    @property
    def with_img_head(self):
        return self.img_head is not None

    def init_extra_weights(self):
        """Initialise what BaseTracker.init_weights does not know about (the SimSiam head)."""
        if self.img_head is not None:
            self.img_head.init_weights()

    # -- one training step --------------------------------------------------------------------------------
    def _embed(self, feat):
        """Backbone output (tuple of levels or a tensor) -> (projection z, prediction p) of its last level."""
        last = feat[-1] if isinstance(feat, tuple) else feat
        return self.img_head(last)

    def forward_img_head(self, x1, x2, clip_len):
        """Loss dict ``{'<shift>.<name>': value}``: shift 0 pairs frame t of view 1 with frame t of view 2; with
        ``intra_video`` every cyclic shift 1..T-1 of view 2 along the clip axis adds a term, all weighted 1/T."""
        (z1, p1), (z2, p2) = self._embed(x1), self._embed(x2)
        shifts = range(clip_len) if self.intra_video else range(1)
        weight = 1.0 / len(shifts)
        z2_clip = p2_clip = None
        if len(shifts) > 1:
            z2_clip, p2_clip = images2video(z2, clip_len), images2video(p2, clip_len)
        out = {}
        for k in shifts:
            zk = z2 if k == 0 else video2images(torch.roll(z2_clip, k, dims=2))
            pk = p2 if k == 0 else video2images(torch.roll(p2_clip, k, dims=2))
            out.update(add_prefix(self.img_head.loss(p1, z1, pk, zk, weight=weight), prefix=str(k)))
        return out

    def forward_train(self, imgs, grids=None, label=None):
        if self.transpose_temporal:                           # loader delivered [B, T, C, 2, H, W]
            imgs = imgs.transpose(1, 3).contiguous()
        assert imgs.ndim == 6 and imgs.size(1) == 2, 'expected [B, 2 views, C, T, H, W]'
        clip_len = imgs.size(3)
        per_view = []
        for view in imgs.unbind(dim=1):                       # separate passes = separate BN batches per view
            per_view.append(self.backbone(video2images(view.contiguous())))
        if not self.with_img_head:
            return {}
        return add_prefix(self.forward_img_head(*per_view, clip_len), prefix='img_head')

    def forward_test(self, imgs, **kwargs):
        raise NotImplementedError
