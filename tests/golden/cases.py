"""Seeded test cases shared by make_golden.py (reference outputs), the oracle tests (CPU) and the GPU parity
tests.  Everything is reproducible from seeds; only reference OUTPUTS live in vfs_golden.npz."""
import torch


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------ backbone
BACKBONE_CASES = {
    # name: reference-config analogue
    'r18_default': dict(depth=18, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(3, ), seed=1,
                        shape=(2, 3, 64, 64)),                       # tests/test_models/test_backbone.py:109-113
    'r50_default': dict(depth=50, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(3, ), seed=2,
                        shape=(2, 3, 64, 64)),                       # test_backbone.py:116-120
    'r50_davis': dict(depth=50, strides=(1, 2, 1, 1), dilations=(1, 1, 1, 1), out_indices=(2, ), seed=3,
                      shape=(2, 3, 72, 104)),                        # configs/r50_nc...:27-36 test_cfg
    'r18_davis': dict(depth=18, strides=(1, 2, 1, 1), dilations=(1, 1, 1, 1), out_indices=(2, ), seed=4,
                      shape=(1, 3, 75, 91)),                         # odd sizes
    'r18_siamfc': dict(depth=18, strides=(1, 2, 1, 1), dilations=(1, 1, 2, 4), out_indices=(3, ), seed=5,
                       shape=(2, 3, 95, 95)),                        # siamfc/default_config_base.py:40-49
    'r50_siamfc': dict(depth=50, strides=(1, 2, 1, 1), dilations=(1, 1, 2, 4), out_indices=(3, ), seed=6,
                       shape=(1, 3, 63, 63)),
    'r50_multi_out': dict(depth=50, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(2, ), seed=7,
                          shape=(3, 3, 96, 64)),
}


def backbone_input(c):
    return torch.randn(c['shape'], generator=_gen(100 + c['seed']))


# ------------------------------------------------------------------ SimSiam head / loss
HEAD_CASES = {
    'r18_head': dict(seed=11, B=4, hw=(2, 2),
                     cfg=dict(in_channels=512, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                              projection_mid_channels=512, projection_out_channels=512, num_predictor_fcs=2,
                              predictor_mid_channels=128, predictor_out_channels=512, with_norm=True,
                              loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg')),
    'r50_head': dict(seed=12, B=8, hw=(2, 3),
                     cfg=dict(in_channels=2048, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                              projection_mid_channels=2048, projection_out_channels=2048, num_predictor_fcs=2,
                              predictor_mid_channels=512, predictor_out_channels=2048, with_norm=True,
                              loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg')),
}


def head_inputs(c):
    g = _gen(200 + c['seed'])
    shape = (c['B'], c['cfg']['in_channels']) + tuple(c['hw'])
    return torch.relu(torch.randn(shape, generator=g)), torch.relu(torch.randn(shape, generator=g))


def loss_inputs():
    g = _gen(300)
    return torch.randn(6, 96, generator=g), torch.randn(6, 96, generator=g)


# ------------------------------------------------------------------ full SimSiam forward_train (config dicts)
def _simsiam_model(depth, in_ch, mid, pred_mid):
    return dict(
        type='SimSiamBaseTracker',
        backbone=dict(type='ResNet', pretrained=None, depth=depth, out_indices=(3, ),
                      norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False, zero_init_residual=True),
        img_head=dict(type='SimSiamHead', in_channels=in_ch, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
                      projection_mid_channels=mid, projection_out_channels=mid, num_predictor_fcs=2,
                      predictor_mid_channels=pred_mid, predictor_out_channels=mid, with_norm=True,
                      loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg'))


TRACKER_TRAIN_CASES = {
    # configs/r18_nc_sgd_cos_100e_r2_1xNx8_k400.py model dict, intra_video=True, clip_len 2
    'r18_intra': dict(seed=21, model=_simsiam_model(18, 512, 512, 128), train_cfg=dict(intra_video=True),
                      shape=(2, 2, 3, 2, 64, 64)),
    # configs/r50_nc_sgd_cos_100e_r5_1xNx2_k400.py model dict
    'r50': dict(seed=22, model=_simsiam_model(50, 2048, 2048, 512), train_cfg=dict(intra_video=False),
                shape=(3, 2, 3, 1, 64, 64)),
}


def tracker_train_input(c):
    return torch.randn(c['shape'], generator=_gen(400 + c['seed']))


# ------------------------------------------------------------------ restricted attention (DAVIS propagation)
ATTENTION_CASES = {
    'small_T1': dict(seed=31, C=32, Cv=3, T=1, H=9, W=11, range=8, temperature=0.07, topk=10),
    'small_T3': dict(seed=32, C=64, Cv=4, T=3, H=12, W=17, range=10, temperature=0.07, topk=10),
    'small_T3_first_free': dict(seed=33, C=64, Cv=4, T=3, H=12, W=17, range=10, temperature=0.07, topk=10,
                                non_mask_len=1),
    'square_mask': dict(seed=34, C=32, Cv=2, T=1, H=10, W=10, range=6, temperature=0.05, topk=5,
                        mask_mode='square'),
    'no_mask': dict(seed=35, C=32, Cv=2, T=2, H=8, W=9, range=None, temperature=0.1, topk=10),
    'cosine_mode': dict(seed=36, C=32, Cv=3, T=2, H=9, W=11, range=8, temperature=1.0, topk=10, mode='cosine'),
    'dup_first': dict(seed=37, C=64, Cv=4, T=3, H=12, W=17, range=10, temperature=0.07, topk=10, dup_first=True),
    'mid_T2': dict(seed=38, C=256, Cv=5, T=2, H=30, W=54, range=24, temperature=0.07, topk=10),
}


def attention_inputs(c):
    g = _gen(500 + c['seed'])
    q = torch.relu(torch.randn(1, c['C'], c['H'], c['W'], generator=g))
    k = torch.relu(torch.randn(1, c['C'], c['T'], c['H'], c['W'], generator=g))
    v = torch.rand(1, c['Cv'], c['T'], c['H'], c['W'], generator=g)
    if c.get('dup_first'):  # frame 0 appears twice in the key set while frame_idx <= precede_frames
        k[:, :, 1] = k[:, :, 0]
        v[:, :, 1] = v[:, :, 0]
    return q, k, v


AFFINITY_CASES = {
    'dense': dict(seed=41, C=32, Cv=3, B=2, H=7, W=9, temperature=0.07, softmax_dim=1, topk=None),
    'dense_topk': dict(seed=42, C=32, Cv=3, B=1, H=8, W=8, temperature=0.07, softmax_dim=1, topk=5),
}


def affinity_inputs(c):
    g = _gen(600 + c['seed'])
    a = torch.randn(c['B'], c['C'], c['H'], c['W'], generator=g)
    b = torch.randn(c['B'], c['C'], c['H'], c['W'], generator=g)
    img = torch.rand(c['B'], c['Cv'], c['H'], c['W'], generator=g)
    return a, b, img


# ------------------------------------------------------------------ SiamFC cross-correlation
XCORR_CASES = {
    'z15_x32': dict(seed=51, C=64, nz=1, nx=3, hz=15, hx=32, out_scale=1e-3),  # exemplar_sz=120 default
    'z16_x32': dict(seed=52, C=64, nz=1, nx=3, hz=16, hx=32, out_scale=1e-3),  # BASELINE cfg-5 (127 px)
    'z6_x11': dict(seed=53, C=512, nz=1, nx=2, hz=6, hx=11, out_scale=1e-5),
}


def xcorr_inputs(c):
    g = _gen(700 + c['seed'])
    z = torch.randn(c['nz'], c['C'], c['hz'], c['hz'], generator=g)
    x = torch.randn(c['nx'], c['C'], c['hx'], c['hx'], generator=g)
    return z, x


# ------------------------------------------------------------------ DAVIS-style inference (VanillaTracker)
TRACKER_TEST_CASES = {
    'r18_clip5': dict(seed=61, T=5, H=64, W=96, num_objs=3,
                      backbone=dict(type='ResNet', pretrained=None, depth=18, out_indices=(2, ),
                                    strides=(1, 2, 1, 1), norm_cfg=dict(type='SyncBN', requires_grad=True),
                                    norm_eval=False, zero_init_residual=True),
                      test_cfg=dict(precede_frames=2, topk=10, temperature=0.07, strides=(1, 2, 1, 1),
                                    out_indices=(2, ), neighbor_range=24, with_first=True,
                                    with_first_neighbor=True, output_dir='eval_results')),
}


def tracker_test_inputs(c):
    g = _gen(800 + c['seed'])
    imgs = torch.randn(1, 1, 3, c['T'], c['H'], c['W'], generator=g)
    seg = torch.zeros(1, c['H'], c['W'])
    # a few rectangles as first-frame objects
    for o in range(1, c['num_objs']):
        y0, x0 = 8 * o, 12 * o
        seg[0, y0:y0 + 24, x0:x0 + 30] = o
    return imgs, seg


# ------------------------------------------------------------------ SiamFC tracker crops (host side, cv2)
SIAMFC_CROP_CASES = {
    # name: (center y, center x, crop size, output size) on a seeded 180 x 240 RGB image
    'inside_z': (90.0, 120.0, 80.3, 120),
    'top_left_x': (10.0, 20.0, 100.7, 255),
    'bottom_right_x': (170.0, 230.0, 150.2, 255),
    'larger_than_image': (90.0, 120.0, 400.0, 255),
    'tiny': (50.0, 60.0, 1.0, 120),
}


# ------------------------------------------------------------------ training data pipeline (crop / resize / flip / normalise)
NORM_CFG = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375])      # configs/*:46-47
TRAIN_PIPELINE_CASES = {
    # the configs' setting: 2 clips of 1 frame, every frame its own crop and flip, 224 x 224
    'k400_2x1': dict(seed=11, H=180, W=240, num_clips=2, clip_len=1, scale=(224, 224), area_range=(0.2, 1.),
                     flip_ratio=0.5, same_on_clip=False, same_across_clip=False, to_bgr=False),
    'clips_2x2_shared': dict(seed=12, H=97, W=131, num_clips=2, clip_len=2, scale=(64, 48), area_range=(0.08, 1.),
                             flip_ratio=0.5, same_on_clip=True, same_across_clip=False, to_bgr=False),
    'upscale_bgr': dict(seed=13, H=40, W=56, num_clips=3, clip_len=1, scale=(96, 96), area_range=(0.2, 1.),
                        flip_ratio=0.7, same_on_clip=False, same_across_clip=False, to_bgr=True),
}


def train_pipeline_frames(c):
    import numpy as np
    rng = np.random.RandomState(1000 + c['seed'])
    n = c['num_clips'] * c['clip_len']
    yy, xx = np.mgrid[0:c['H'], 0:c['W']]
    frames = []
    for i in range(n):
        smooth = 127 + 90 * np.sin(yy[..., None] / (7.0 + i) + np.arange(3)) * np.cos(xx[..., None] / (11.0 + 2 * i))
        frames.append(np.clip(smooth + rng.randint(-30, 31, (c['H'], c['W'], 3)), 0, 255).astype(np.uint8))
    return frames


# ------------------------------------------------------------------ SiamFC tracker (init + updates on a moving blob)
SIAMFC_TRACKER_CASES = {
    # reference default exemplar size (default_config_base.py:6) and the 127 px of BASELINE cfg-5
    'r18_convfc_120': dict(depth=18, exemplar_sz=120, extra_conv=True, out_scale=1e-3, seed=71),
    'r18_convfc_127': dict(depth=18, exemplar_sz=127, extra_conv=True, out_scale=1e-3, seed=73),
    'r18_fc_127': dict(depth=18, exemplar_sz=127, extra_conv=False, out_scale=1e-3, seed=75),
}


def siamfc_tracker_frames(num=4):
    """RGB uint8 frames [240,320,3] with a bright blob drifting down-right, and its 1-indexed (x, y, w, h) box."""
    import numpy as np
    rng = np.random.RandomState(5)
    base = (rng.rand(240, 320, 3) * 60 + 60)
    frames = []
    for f in range(num):
        img = base.copy()
        cy, cx = 120 + 6 * f, 150 + 9 * f
        yy, xx = np.mgrid[0:240, 0:320]
        blob = np.exp(-(((yy - cy) / 14.0)**2 + ((xx - cx) / 20.0)**2))
        img += 150 * blob[..., None] * np.array([1.0, 0.6, 0.2])
        frames.append(np.clip(img, 0, 255).astype(np.uint8))
    return frames, [150 - 30 + 1, 120 - 21 + 1, 60, 42]


def siamfc_tracker_cfg(c):
    """default_config_base.py:2-51 with the case's overrides, as a plain nested dict."""
    cfg = dict(out_scale=c['out_scale'], exemplar_sz=c['exemplar_sz'], instance_sz=255, context=0.5, scale_num=3,
               scale_step=1.0375, scale_lr=0.59, scale_penalty=0.9745, window_influence=0.176, response_sz=17,
               response_up=16, total_stride=8, epoch_num=50, batch_size=8, num_workers=8, initial_lr=1e-3,
               ultimate_lr=1e-5, weight_decay=5e-4, momentum=0.9, r_pos=16, r_neg=0, pairs_per_seq=1, optimizer='Adam',
               loss='focal', lr_schedule='exp', lr_step_size=10, extra_conv=c['extra_conv'], out_channels=512,
               reduction=1, auto_resume=False, force_wd=False, out_block_index=None, checkpoint=None,
               work_dir='/tmp', suffix='vfs_golden',
               model=dict(backbone=dict(type='ResNet', depth=c['depth'], pretrained=None, frozen_stages=4,
                                        dilations=(1, 1, 2, 4), strides=(1, 2, 1, 1), out_indices=(3, ), with_cp=False,
                                        norm_eval=True, norm_cfg=dict(type='SyncBN', requires_grad=True))))
    return cfg


SIAMFC_TRAIN_CASES = {
    'focal_adam': dict(depth=18, exemplar_sz=127, extra_conv=True, out_scale=1e-3, seed=81, loss='focal',
                       optimizer='Adam', batch=2),
    'balance_sgd': dict(depth=18, exemplar_sz=127, extra_conv=True, out_scale=1e-3, seed=83, loss='balance',
                        optimizer='SGD', batch=3),
}


def siamfc_train_batches(c, steps=2):
    """(z [B,3,127,127], x [B,3,255,255]) float batches of 0..255 pixel values, like the GOT-10k pair loader yields."""
    g = _gen(1500 + c['seed'])
    out = []
    for _ in range(steps):
        z = torch.rand(c['batch'], 3, c['exemplar_sz'], c['exemplar_sz'], generator=g) * 255
        x = torch.rand(c['batch'], 3, 255, 255, generator=g) * 255
        out.append((z, x))
    return out


def siamfc_train_cfg(c):
    cfg = siamfc_tracker_cfg(c)
    cfg.update(loss=c['loss'], optimizer=c['optimizer'], lr_schedule='fixed', gpus=None)
    return cfg


def siamfc_image():
    import numpy as np
    return np.random.RandomState(900).randint(0, 256, (180, 240, 3)).astype(np.uint8)


# ------------------------------------------------------------------ attention, general forms (bool masks, topk=None)
ATTENTION_EXTRA_CASES = {
    # name: N, C, Cv, T, (Hq, Wq), (Hk, Wk), mask kind, topk, mode, non_mask_len, temperature
    'bool2d_topk': dict(seed=71, N=1, C=64, Cv=3, T=2, q=(9, 11), k=(9, 11), mask='random2d', topk=5, mode='softmax',
                        non_mask_len=0, temperature=0.07),
    'bool2d_first_free': dict(seed=72, N=1, C=64, Cv=4, T=3, q=(8, 8), k=(8, 8), mask='random2d', topk=10,
                              mode='softmax', non_mask_len=1, temperature=0.07),
    'window_dense_softmax': dict(seed=73, N=1, C=32, Cv=3, T=2, q=(9, 11), k=(9, 11), mask='window', topk=None,
                                 mode='softmax', non_mask_len=0, temperature=0.07),
    'nomask_dense_softmax': dict(seed=74, N=2, C=64, Cv=20, T=1, q=(7, 9), k=(7, 9), mask=None, topk=None,
                                 mode='softmax', non_mask_len=0, temperature=0.07),
    'bool3d_topk': dict(seed=75, N=2, C=64, Cv=3, T=1, q=(8, 9), k=(8, 9), mask='random3d', topk=4, mode='softmax',
                        non_mask_len=0, temperature=0.07),
    'dense_cosine': dict(seed=76, N=1, C=64, Cv=2, T=2, q=(6, 7), k=(6, 7), mask='random2d', topk=None, mode='cosine',
                         non_mask_len=0, temperature=1.0),
    'rect_query_vs_key': dict(seed=77, N=1, C=64, Cv=3, T=2, q=(6, 7), k=(8, 9), mask=None, topk=6, mode='softmax',
                              non_mask_len=0, temperature=0.07),
}


def attention_extra_inputs(c):
    """-> (query, key, value, mask tensor | ('window', range) | None)."""
    g = _gen(900 + c['seed'])
    (hq, wq), (hk, wk) = c['q'], c['k']
    q = torch.relu(torch.randn(c['N'], c['C'], hq, wq, generator=g))
    k = torch.relu(torch.randn(c['N'], c['C'], c['T'], hk, wk, generator=g))
    v = torch.rand(c['N'], c['Cv'], c['T'], hk, wk, generator=g)
    mask = None
    if c['mask'] == 'random2d':
        mask = torch.rand(hk * wk, hq * wq, generator=g) > 0.4
        mask[:16] = True                                    # every query keeps at least 16 keys
    elif c['mask'] == 'random3d':
        mask = torch.rand(c['N'], hk * wk, hq * wq, generator=g) > 0.4
        mask[:, :16] = True
    elif c['mask'] == 'window':
        mask = ('window', 8)
    return q, k, v, mask
