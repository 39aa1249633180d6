"""Generates tests/golden/vfs_golden.npz from the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shim.py).  Run in the authoring container only:

    python tests/golden/make_golden.py

Inputs and weights are reproducible from seeds alone (``oracle.seeded_state_dict`` fills parameters by
state-dict name), so only the reference's OUTPUTS are stored.  The fixtures pin the oracle (tests/test_oracle_*.py,
CPU) and the CUDA path (tests/test_gpu_*.py, on the B200 box where /root/reference does not exist).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, seeded_state_dict  # noqa: E402
from tests.golden import cases  # noqa: E402


def reference_outputs():
    """Runs every golden case through the unmodified reference and returns {key: ndarray}.  Also used by
    tests/test_oracle_golden.py::test_oracle_bit_exact_vs_live_reference when /root/reference is present."""
    ref = ref_shim.load_reference()
    out = {}

    # ---- backbone (eval and train-mode BN), reference ResNet.forward
    for name, c in cases.BACKBONE_CASES.items():
        m = ref.ResNet(c['depth'], norm_cfg=dict(type='SyncBN', requires_grad=True), strides=c['strides'],
                       dilations=c['dilations'], out_indices=c['out_indices'])
        m.load_state_dict(seeded_state_dict(m, seed=c['seed']))
        x = cases.backbone_input(c)
        with torch.no_grad():
            m.train(False)
            y = m(x)
            out[f'backbone/{name}/eval'] = y.numpy()
            m.train(True)
            out[f'backbone/{name}/train'] = m(x).numpy()

    # ---- SimSiam head + loss
    for name, c in cases.HEAD_CASES.items():
        h = ref.SimSiamHead(**c['cfg'])
        h.load_state_dict(seeded_state_dict(h, seed=c['seed']))
        x1, x2 = cases.head_inputs(c)
        for mode in ('eval', 'train'):
            h.train(mode == 'train')
            with torch.no_grad():
                z1, p1 = h(x1)
                z2, p2 = h(x2)
                loss = h.loss(p1, z1, p2, z2)['loss_feat']
            out[f'head/{name}/{mode}/z1'] = z1.numpy()
            out[f'head/{name}/{mode}/p1'] = p1.numpy()
            out[f'head/{name}/{mode}/loss'] = loss.numpy()
    p, z = cases.loss_inputs()
    for neg in (False, True):
        out[f'loss/cosine/neg{int(neg)}'] = ref.CosineSimLoss(negative=neg)(p, z).numpy()

    # ---- full SimSiam forward_train through the reference tracker + config dict
    for name, c in cases.TRACKER_TRAIN_CASES.items():
        model = ref.build_model(c['model'], train_cfg=c['train_cfg'], test_cfg=None)
        model.load_state_dict(seeded_state_dict(model, seed=c['seed']))
        model.train()
        imgs = cases.tracker_train_input(c)
        with torch.no_grad():
            losses = model.forward_train(imgs)
        for k, v in losses.items():
            out[f'tracker_train/{name}/{k}'] = v.numpy()

    # ---- restricted attention / affinity
    for name, c in cases.ATTENTION_CASES.items():
        q, k, v = cases.attention_inputs(c)
        mask = ref.spatial_neighbor(1, c['H'], c['W'], neighbor_range=c['range'], device='cpu', dtype=torch.float32,
                                    mode=c.get('mask_mode', 'circle')) if c['range'] else None
        if mask is not None:
            out[f'attention/{name}/mask_packed'] = np.packbits(mask.numpy())
        o = ref.masked_attention_efficient(q, k, v, mask, temperature=c['temperature'], topk=c['topk'],
                                           non_mask_len=c.get('non_mask_len', 0), mode=c.get('mode', 'softmax'))
        out[f'attention/{name}/out'] = o.numpy()
    for name, c in cases.AFFINITY_CASES.items():
        a, b, img = cases.affinity_inputs(c)
        aff = ref.compute_affinity(a, b, temperature=c['temperature'], softmax_dim=c['softmax_dim'])
        out[f'affinity/{name}/aff'] = aff.numpy()
        out[f'affinity/{name}/prop'] = ref.propagate(img, aff.clone(), topk=c['topk']).numpy()

    # ---- SiamFC heads
    for name, c in cases.XCORR_CASES.items():
        z, x = cases.xcorr_inputs(c)
        out[f'xcorr/{name}/siamfc'] = ref.siamfc_heads.SiamFC(out_scale=c['out_scale'])(z, x).numpy()
        m = ref.siamfc_heads.SiamConvFC(c['C'], c['C'], out_scale=c['out_scale'])
        m.load_state_dict(seeded_state_dict(m, seed=c['seed']))
        with torch.no_grad():
            out[f'xcorr/{name}/siamconvfc'] = m(z, x).numpy()

    # ---- DAVIS-style inference through the reference VanillaTracker
    for name, c in cases.TRACKER_TEST_CASES.items():
        tr = ref.VanillaTracker(backbone=c['backbone'], test_cfg=ref_shim.sys.modules['mmcv'].ConfigDict(c['test_cfg']))
        tr.backbone.load_state_dict(seeded_state_dict(tr.backbone, seed=c['seed']))
        tr.eval()
        imgs, seg = cases.tracker_test_inputs(c)
        with torch.no_grad():
            preds = tr.forward_test(imgs, seg, [dict(original_shape=(c['H'], c['W'], 3))])
        out[f'tracker_test/{name}/preds'] = np.asarray(preds[0]).astype(np.uint8)

    return out


def siamfc_crop_outputs():
    """Crops of the reference's own siamfc/ops.py::crop_and_resize (cv2) -> tests/golden/siamfc_crop_golden.npz."""
    refops = ref_shim.load_reference_siamfc_ops()
    img = cases.siamfc_image()
    out = {}
    for name, (cy, cx, size, out_size) in cases.SIAMFC_CROP_CASES.items():
        out[name] = refops.crop_and_resize(img, np.array([cy, cx], dtype=np.float32), size, out_size=out_size,
                                           border_value=np.mean(img, axis=(0, 1)))
    return out


def siamfc_tracker_outputs():
    """The UNMODIFIED reference TrackerSiamFC (siamfc_tracker_base.py:88-319, got10k base class stubbed) run on the
    synthetic sequence: exemplar kernel, per-frame raw responses and boxes -> tests/golden/siamfc_tracker_golden.npz."""
    import logging
    import oracle
    from vfs_b200.mmcv_lite import ConfigDict
    mod = ref_shim.load_reference_siamfc_tracker()
    frames, box0 = cases.siamfc_tracker_frames()
    out = {}
    for name, c in cases.SIAMFC_TRACKER_CASES.items():
        trk = mod.TrackerSiamFC(ConfigDict(cases.siamfc_tracker_cfg(c)), logging.getLogger('ref_siamfc'))
        trk.net.backbone.load_state_dict(oracle.seeded_state_dict(trk.net.backbone, seed=c['seed']))
        if c['extra_conv']:
            trk.net.head.load_state_dict(oracle.seeded_state_dict(trk.net.head, seed=c['seed'] + 1))
        trk.net.cpu()
        trk.device = torch.device('cpu')
        # record the raw responses the reference computes inside update(): same net, same crops, no state change
        trk.init(frames[0], box0)
        out[f'{name}/kernel'] = trk.kernel.numpy().copy()
        boxes, responses = [], []
        for img in frames[1:]:
            x = np.stack([mod.ops.crop_and_resize(img, trk.center, trk.x_sz * f, out_size=trk.cfg.instance_sz,
                                                  border_value=trk.avg_color) for f in trk.scale_factors], axis=0)
            with torch.no_grad():
                xt = trk.normalize(torch.from_numpy(x).permute(0, 3, 1, 2).float())
                responses.append(trk.net.head(trk.kernel, trk.net.backbone(xt)).squeeze(1).numpy().copy())
            boxes.append(trk.update(img).copy())
        out[f'{name}/responses'] = np.stack(responses)
        out[f'{name}/boxes'] = np.stack(boxes)
        out[f'{name}/state'] = np.concatenate([trk.center, trk.target_sz, [trk.z_sz, trk.x_sz]]).astype(np.float64)
    return out


def train_pipeline_outputs():
    """The UNMODIFIED reference pipeline classes RandomResizedCrop -> Resize -> Flip -> Normalize -> FormatShape
    (configs/*:48-92) on seeded synthetic frames -> tests/golden/train_pipeline_golden.npz."""
    import random
    ns = ref_shim.load_reference_pipelines()
    A, F = ns.augmentations, ns.formating
    out = {}
    for name, c in cases.TRAIN_PIPELINE_CASES.items():
        frames = cases.train_pipeline_frames(c)
        np.random.seed(c['seed'])
        random.seed(c['seed'])
        results = dict(imgs=[f.copy() for f in frames], img_shape=(c['H'], c['W']), original_shape=(c['H'], c['W']),
                       clip_len=c['clip_len'], num_clips=c['num_clips'], modality='RGB')
        steps = [A.RandomResizedCrop(area_range=c['area_range'], same_across_clip=c['same_across_clip'],
                                     same_on_clip=c['same_on_clip']),
                 A.Resize(scale=c['scale'], keep_ratio=False),
                 A.Flip(flip_ratio=c['flip_ratio'], same_across_clip=c['same_across_clip'],
                        same_on_clip=c['same_on_clip']),
                 A.Normalize(to_bgr=c['to_bgr'], **cases.NORM_CFG),
                 F.FormatShape(input_format='NCTHW')]
        for step in steps:
            results = step(results)
        out[name] = np.ascontiguousarray(results['imgs'])
    return out


def siamfc_train_outputs():
    """The UNMODIFIED reference ``TrackerSiamFC.train_step`` (siamfc_tracker_base.py:364-386: frozen R18 backbone,
    SiamConvFC head, Focal / Balanced loss, Adam / SGD) for two steps -> tests/golden/siamfc_train_golden.npz: losses,
    head gradients of the first step (recovered from the optimiser state) and head parameters after the second."""
    import logging
    import oracle
    from vfs_b200.mmcv_lite import ConfigDict
    mod = ref_shim.load_reference_siamfc_tracker()
    out = {}
    for name, c in cases.SIAMFC_TRAIN_CASES.items():
        trk = mod.TrackerSiamFC(ConfigDict(cases.siamfc_train_cfg(c)), logging.getLogger('ref_siamfc'))
        trk.net.backbone.load_state_dict(oracle.seeded_state_dict(trk.net.backbone, seed=c['seed']))
        trk.net.head.load_state_dict(oracle.seeded_state_dict(trk.net.head, seed=c['seed'] + 1))
        trk.net.cpu()
        trk.device, trk.cuda = torch.device('cpu'), False
        head_params = dict(trk.net.head.named_parameters())
        losses = []
        for i, batch in enumerate(cases.siamfc_train_batches(c)):
            losses.append(trk.train_step(batch, backward=True))
            if i == 0:
                for k, p in head_params.items():
                    st = trk.optimizer.state[p]
                    g = st['exp_avg'] / 0.1 if c['optimizer'] == 'Adam' else st['momentum_buffer']
                    out[f'{name}/grad/{k}'] = g.detach().numpy().copy()
        out[f'{name}/losses'] = np.asarray(losses, dtype=np.float64)
        for k, p in head_params.items():
            out[f'{name}/param/{k}'] = p.detach().numpy().copy()
    return out


def attention_extra_outputs():
    """masked_attention_efficient of the unmodified reference for arbitrary bool masks / topk=None / rectangular maps
    -> tests/golden/attention_extra_golden.npz."""
    ref = ref_shim.load_reference()
    out = {}
    for name, c in cases.ATTENTION_EXTRA_CASES.items():
        q, k, v, mask = cases.attention_extra_inputs(c)
        if isinstance(mask, tuple):
            mask = ref.spatial_neighbor(1, c['k'][0], c['k'][1], neighbor_range=mask[1], device='cpu',
                                        dtype=torch.float32, mode='circle')
        with torch.no_grad():
            out[name] = ref.masked_attention_efficient(q, k, v, mask, temperature=c['temperature'], topk=c['topk'],
                                                       non_mask_len=c['non_mask_len'], mode=c['mode']).numpy()
    return out


def main():
    extra = attention_extra_outputs()
    extra_path = os.path.join(ROOT, 'tests', 'golden', 'attention_extra_golden.npz')
    np.savez_compressed(extra_path, **extra)
    print(f'wrote {extra_path}: {len(extra)} arrays')
    crops = siamfc_crop_outputs()
    crop_path = os.path.join(ROOT, 'tests', 'golden', 'siamfc_crop_golden.npz')
    np.savez_compressed(crop_path, **crops)
    print(f'wrote {crop_path}: {len(crops)} arrays')
    tp = train_pipeline_outputs()
    tp_path = os.path.join(ROOT, 'tests', 'golden', 'train_pipeline_golden.npz')
    np.savez_compressed(tp_path, **tp)
    print(f'wrote {tp_path}: {len(tp)} arrays', {k: v.shape for k, v in tp.items()})
    st = siamfc_train_outputs()
    st_path = os.path.join(ROOT, 'tests', 'golden', 'siamfc_train_golden.npz')
    np.savez_compressed(st_path, **st)
    print(f'wrote {st_path}: {len(st)} arrays')
    trk = siamfc_tracker_outputs()
    trk_path = os.path.join(ROOT, 'tests', 'golden', 'siamfc_tracker_golden.npz')
    np.savez_compressed(trk_path, **trk)
    print(f'wrote {trk_path}: {len(trk)} arrays')
    if '--only-siamfc' in sys.argv or '--only-small' in sys.argv:
        return
    torch.set_num_threads(8)
    out = reference_outputs()
    path = os.path.join(ROOT, 'tests', 'golden', 'vfs_golden.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB')
    for k in sorted(out):
        print(f'  {k:48s} {out[k].dtype} {out[k].shape}')


if __name__ == '__main__':
    main()
