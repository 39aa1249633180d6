"""ORACLE -- test infrastructure only (see oracle/__init__.py).

    SimSiamBaseTracker.forward_train / forward_img_head   mmaction/models/trackers/sim_siam_base_tracker.py:31-76
    VanillaTracker.forward_test / get_feats               mmaction/models/trackers/vanilla_tracker.py:55-206
    video2images / images2video / pil_nearest_interpolate mmaction/models/common/utils.py:25-64

Composed from the other oracle pieces (resnet_forward, simsiam_head_forward, simsiam_loss, masked_attention_efficient,
spatial_neighbor) with the reference's own tensor bookkeeping; pinned to the fixtures `tracker_train/*` and
`tracker_test/*` of tests/golden/vfs_golden.npz (outputs of the unmodified reference trackers).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .attention import masked_attention_efficient, spatial_neighbor
from .head import simsiam_head_forward, simsiam_loss
from .resnet import resnet_forward


def video2images(imgs):
    """[B,C,T,H,W] -> [B*T,C,H,W] (utils.py:45-53)."""
    batches, channels, clip_len = imgs.shape[:3]
    if clip_len == 1:
        return imgs.squeeze(2).reshape(batches, channels, *imgs.shape[3:])
    return imgs.transpose(1, 2).contiguous().reshape(batches * clip_len, channels, *imgs.shape[3:])


def images2video(imgs, clip_len):
    """[B*T,C,...] -> [B,C,T,...] (utils.py:56-64)."""
    batches, channels = imgs.shape[:2]
    if clip_len == 1:
        return imgs.unsqueeze(2)
    return imgs.reshape(batches // clip_len, clip_len, channels, *imgs.shape[2:]).transpose(1, 2).contiguous()


def pil_nearest_interpolate(input, size):
    """Per-image Pillow NEAREST resize of [N,1,H,W] to ``size`` = (h, w) (utils.py:25-42 -> mmcv.imresize(...,
    interpolation='nearest', backend='pillow'))."""
    from PIL import Image
    out = []
    for img in input.permute(0, 2, 3, 1):
        arr = img.squeeze(-1).detach().cpu().numpy()
        pil = Image.fromarray(arr)
        res = np.array(pil.resize((size[1], size[0]), Image.NEAREST))
        out.append(torch.from_numpy(res).to(input).unsqueeze(2).permute(2, 0, 1))
    return torch.stack(out, dim=0)


def simsiam_forward_train(sd, imgs, depth, intra_video=False, bn_training=True):
    """Loss dict of SimSiamBaseTracker.forward_train for ``imgs`` [B,2,C,T,H,W]; ``sd`` = the tracker's state dict
    (keys ``backbone.*``, ``img_head.*``).  Differentiable (plain torch ops)."""
    assert imgs.size(1) == 2 and imgs.ndim == 6
    bsd = {k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')}
    hsd = {k[len('img_head.'):]: v for k, v in sd.items() if k.startswith('img_head.')}
    clip_len = imgs.size(3)
    i1 = video2images(imgs[:, 0].contiguous().reshape(-1, *imgs.shape[2:]))
    i2 = video2images(imgs[:, 1].contiguous().reshape(-1, *imgs.shape[2:]))
    z1, p1 = simsiam_head_forward(hsd, resnet_forward(bsd, i1, depth, bn_training=bn_training), bn_training=bn_training)
    z2, p2 = simsiam_head_forward(hsd, resnet_forward(bsd, i2, depth, bn_training=bn_training), bn_training=bn_training)
    w = 1. / clip_len if intra_video else 1.
    losses = {'img_head.0.loss_feat': simsiam_loss(p1, z1, p2, z2, weight=w)}
    if intra_video:
        z2v, p2v = images2video(z2, clip_len), images2video(p2, clip_len)
        for i in range(1, clip_len):
            losses[f'img_head.{i}.loss_feat'] = simsiam_loss(p1, z1, video2images(p2v.roll(i, dims=2)),
                                                             video2images(z2v.roll(i, dims=2)), weight=w)
    return losses


def vanilla_forward_test(bsd, imgs, ref_seg_map, original_shape, depth, strides, out_indices, test_cfg,
                         dilations=(1, 1, 1, 1), bn_training=False):
    """VanillaTracker.forward_test for label-id input (``ref_seg_map`` [1,H,W]): list with one array [T,H,W]
    (vanilla_tracker.py:80-206; feature extraction :55-75).  ``bsd`` = backbone state dict."""
    imgs = imgs.reshape((-1, ) + imgs.shape[2:])
    assert imgs.shape[0] == 1
    clip_len = imgs.size(2)
    frames = video2images(imgs)
    step = test_cfg.get('batch_step', 10)
    with torch.no_grad():
        feats = torch.cat([resnet_forward(bsd, frames[p:p + step], depth, strides, dilations, out_indices,
                                          bn_training=bn_training) for p in range(0, clip_len, step)], dim=0)
    feat_bank = images2video(feats, clip_len)                                   # [1,C,T,h,w]
    fh, fw = feats.shape[2:]
    resized = pil_nearest_interpolate(ref_seg_map.unsqueeze(1), size=(fh, fw)).squeeze(1).long()
    resized = F.one_hot(resized).permute(0, 3, 1, 2).float()
    ref_full = F.interpolate(ref_seg_map.unsqueeze(1), size=original_shape[:2], mode='nearest').squeeze(1)
    seg_bank = [resized]
    seg_preds = [ref_full.numpy()]
    rng = test_cfg.get('neighbor_range', None)
    mask = spatial_neighbor(fh, fw, rng, mode='circle') if rng is not None else None
    for f in range(1, clip_len):
        key_start = max(0, f - test_cfg['precede_frames'])
        query = feat_bank[:, :, f]
        key = feat_bank[:, :, key_start:f]
        value = torch.stack(seg_bank[key_start:f], dim=2)
        if test_cfg.get('with_first', True):
            key = torch.cat([feat_bank[:, :, 0:1], key], dim=2)
            value = torch.cat([seg_bank[0].unsqueeze(2), value], dim=2)
        seg_logit = masked_attention_efficient(query, key, value, mask, temperature=test_cfg['temperature'],
                                               topk=test_cfg['topk'], normalize=test_cfg.get('with_norm', True),
                                               non_mask_len=0 if test_cfg.get('with_first_neighbor', True) else 1)
        seg_bank.append(seg_logit)
        seg_pred = F.interpolate(seg_logit, size=original_shape[:2], mode='bilinear', align_corners=False)
        lo = seg_pred.view(*seg_pred.shape[:2], -1).min(dim=-1)[0].view(*seg_pred.shape[:2], 1, 1)
        hi = seg_pred.view(*seg_pred.shape[:2], -1).max(dim=-1)[0].view(*seg_pred.shape[:2], 1, 1)
        seg_pred = torch.where(hi > 0, (seg_pred - lo) / (hi - lo + 1e-12), seg_pred).argmax(dim=1)
        seg_pred = F.interpolate(seg_pred.byte().unsqueeze(1), size=original_shape[:2], mode='nearest').squeeze(1)
        seg_preds.append(seg_pred.numpy())
    return list(np.stack(seg_preds, axis=1))
