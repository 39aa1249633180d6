"""ORACLE -- test infrastructure only (see oracle/__init__.py).

    SiamFC._fast_xcorr / SiamConvFC.forward   projects/siamfc-pytorch/siamfc/heads.py:16-23, 46-58
"""
import torch.nn.functional as F


def xcorr(z, x, out_scale=0.001):
    """Per-pair valid cross-correlation via grouped conv (heads.py:16-23): z [nz,c,hz,wz], x [nx,c,h,w]."""
    nz = z.size(0)
    nx, c, h, w = x.size()
    out = F.conv2d(x.reshape(-1, nz * c, h, w), z, groups=nz)
    return out.reshape(nx, -1, out.size(-2), out.size(-1)) * out_scale


def siam_conv_fc(sd, z, x, out_scale=0.001, num_convs=1):
    """1x1 (biased) conv adapters on both branches then xcorr (heads.py:46-49)."""
    for i in range(num_convs):
        z = F.conv2d(z, sd[f'z_convs.{i}.weight'], sd[f'z_convs.{i}.bias'])
        x = F.conv2d(x, sd[f'x_convs.{i}.weight'], sd[f'x_convs.{i}.bias'])
    return xcorr(z, x, out_scale)


# ----------------------------------------------------------------------------------------------------------------
# TrackerSiamFC inference (projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py:200-319), restated with the same
# cv2 / numpy calls.  PINNED: oracle/ref_shim.py::load_reference_siamfc_tracker imports the unmodified reference file
# (its base class got10k.trackers.Tracker is a trivial stub); tests/golden/siamfc_tracker_golden.npz holds its kernel /
# responses / boxes on a synthetic sequence and tests/test_oracle_golden.py checks this restatement against it -- bit
# for bit where the reference tree is present.
# ----------------------------------------------------------------------------------------------------------------
def crop_and_resize(img, center, size, out_size, border_value=None):
    """ops.py:87-104 (faster=True) -> image_utils.py:7-76: crop around ``center`` (y, x), resize, pad with the mean
    colour."""
    import cv2
    import numpy as np
    size = max(2, size)
    b = np.array([center[1], center[0], size, size]).astype(np.float32)          # bbox_utils.py:27-40 (list input)
    x1, y1, x2, y2 = b[0] - b[2] / 2.0, b[1] - b[3] / 2.0, b[0] + b[2] / 2.0, b[1] + b[3] / 2.0
    avg = np.mean(img, axis=(0, 1), dtype=float)
    w, h = float(x2 - x1), float(y2 - y1)
    xc, yc = float(x1 + x2) / 2, float(y1 + y2) / 2
    box = np.round(np.array([xc - w / 2, yc - h / 2, xc + w / 2, yc + h / 2])).astype(int)
    bw = np.array([box[2] - box[0], box[3] - box[1]])
    H, W = img.shape[:2]
    patch = img[max(box[1], 0):min(box[3], H), max(box[0], 0):min(box[2], W), :]
    bounded = np.clip(box, 0, np.array([W, H, W, H]))
    bwh = np.array([bounded[2] - bounded[0], bounded[3] - bounded[1]])
    if patch.shape[0] == 0 or patch.shape[1] == 0:
        return np.zeros((int(out_size), int(out_size), 3), dtype=patch.dtype)
    patch = cv2.resize(patch, (max(1, int(np.round(out_size * bwh[0] / bw[0]))),
                               max(1, int(np.round(out_size * bwh[1] / bw[1])))), interpolation=cv2.INTER_LINEAR)
    pad = np.zeros(4, dtype=int)
    pad[:2] = np.maximum(0, -box[:2] * out_size / bw)
    pad[2:] = out_size - (pad[:2] + np.array(patch.shape)[[1, 0]])
    if np.any(pad != 0):
        if len(pad[pad < 0]) > 0:
            return np.zeros((int(out_size), int(out_size), 3))
        return cv2.copyMakeBorder(patch, pad[1], pad[3], pad[0], pad[2], cv2.BORDER_CONSTANT, value=avg)
    return patch


def response_peak(responses, hann_window, upscale_sz, scale_num, scale_penalty, window_influence):
    """siamfc_tracker_base.py:263-291: responses float32 [S,R,R] -> (scale_id, (row, col), blended map)."""
    import cv2
    import numpy as np
    responses = np.stack([cv2.resize(u, (upscale_sz, upscale_sz), interpolation=cv2.INTER_CUBIC) for u in responses])
    responses[:scale_num // 2] *= scale_penalty
    responses[scale_num // 2 + 1:] *= scale_penalty
    scale_id = np.argmax(np.amax(responses, axis=(1, 2)))
    response = responses[scale_id]
    response -= response.min()
    response /= response.sum() + 1e-16
    response = (1 - window_influence) * response + window_influence * hann_window
    loc = np.unravel_index(response.argmax(), response.shape)
    return int(scale_id), (int(loc[0]), int(loc[1])), response


class TrackerOracle:
    """init / update of the reference tracker on the CPU: oracle backbone + head, cv2 crops and post-processing."""

    def __init__(self, cfg, backbone_sd, head_sd, depth):
        self.cfg, self.bsd, self.hsd, self.depth = cfg, backbone_sd, head_sd, depth

    def _features(self, crops, exemplar=False):
        import numpy as np
        import torch
        from .resnet import resnet_forward
        if exemplar:
            # siamfc_tracker_base.py:239-241 builds the exemplar batch with permute(2,0,1).unsqueeze(0): the size-1
            # batch stride makes ATen treat it as NCHW, the stacked search batch (:258-260) is genuinely channels-last
            # -- different conv kernels, different fp32 summation order; restated exactly for bit parity
            x = torch.from_numpy(np.ascontiguousarray(crops)).permute(2, 0, 1).unsqueeze(0).float()
        else:
            x = torch.from_numpy(np.ascontiguousarray(crops)).permute(0, 3, 1, 2).float()
        mean = torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1)
        std = torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1)
        b = self.cfg['model']['backbone']
        with torch.no_grad():
            return resnet_forward(self.bsd, (x - mean) / std, self.depth, b['strides'], b['dilations'],
                                  b['out_indices'])

    def init(self, img, box):
        import numpy as np
        cfg = self.cfg
        box = np.array([box[1] - 1 + (box[3] - 1) / 2, box[0] - 1 + (box[2] - 1) / 2, box[3], box[2]],
                       dtype=np.float32)
        self.center, self.target_sz = box[:2], box[2:]
        self.upscale_sz = cfg['response_up'] * cfg['response_sz']
        self.hann_window = np.outer(np.hanning(self.upscale_sz), np.hanning(self.upscale_sz))
        self.hann_window /= self.hann_window.sum()
        self.scale_factors = cfg['scale_step']**np.linspace(-(cfg['scale_num'] // 2), cfg['scale_num'] // 2,
                                                            cfg['scale_num'])
        context = cfg['context'] * np.sum(self.target_sz)
        self.z_sz = np.sqrt(np.prod(self.target_sz + context))
        self.x_sz = self.z_sz * cfg['instance_sz'] / cfg['exemplar_sz']
        z = crop_and_resize(img, self.center, self.z_sz, cfg['exemplar_sz'])
        self.kernel = self._features(z, exemplar=True)

    def responses(self, img):
        import numpy as np
        import torch
        cfg = self.cfg
        x = np.stack([crop_and_resize(img, self.center, self.x_sz * f, cfg['instance_sz'])
                      for f in self.scale_factors], axis=0)
        feats = self._features(x)
        with torch.no_grad():
            if self.hsd is not None:
                r = siam_conv_fc(self.hsd, self.kernel, feats, cfg['out_scale'])
            else:
                r = xcorr(self.kernel, feats, cfg['out_scale'])
        return r.squeeze(1).numpy()

    def update(self, img, responses=None):
        import numpy as np
        cfg = self.cfg
        if responses is None:
            responses = self.responses(img)
        scale_id, loc, _ = response_peak(responses, self.hann_window, self.upscale_sz, cfg['scale_num'],
                                         cfg['scale_penalty'], cfg['window_influence'])
        disp_in_response = np.array(loc) - (self.upscale_sz - 1) / 2
        disp_in_instance = disp_in_response * cfg['total_stride'] / cfg['response_up']
        disp_in_image = disp_in_instance * self.x_sz * self.scale_factors[scale_id] / cfg['instance_sz']
        self.center += disp_in_image
        scale = (1 - cfg['scale_lr']) * 1.0 + cfg['scale_lr'] * self.scale_factors[scale_id]
        self.target_sz *= scale
        self.z_sz *= scale
        self.x_sz *= scale
        return np.array([self.center[1] + 1 - (self.target_sz[1] - 1) / 2,
                         self.center[0] + 1 - (self.target_sz[0] - 1) / 2, self.target_sz[1], self.target_sz[0]])


# ----------------------------------------------------------------------------------------------------------------
# TrackerSiamFC.train_step (siamfc_tracker_base.py:364-386) with the frozen backbone of default_config_base.py:40-49:
# labels (_create_labels :468-500), FocalLoss / BalancedLoss (siamfc/losses.py:27-64), torch.optim.Adam / SGD.
# PINNED by tests/golden/siamfc_train_golden.npz (the unmodified reference class, oracle/ref_shim.py).
# ----------------------------------------------------------------------------------------------------------------
def create_labels(size, r_pos, r_neg, total_stride):
    import numpy as np
    n, c, h, w = size
    x = np.arange(w) - (w - 1) / 2
    y = np.arange(h) - (h - 1) / 2
    x, y = np.meshgrid(x, y)
    dist = np.abs(x) + np.abs(y)
    rp, rn = r_pos / total_stride, r_neg / total_stride
    labels = np.where(dist <= rp, np.ones_like(x), np.where(dist < rn, np.ones_like(x) * 0.5, np.zeros_like(x)))
    return np.tile(labels.reshape((1, 1, h, w)), (n, c, 1, 1))


def focal_loss(input, target, gamma=2):
    import torch
    pos_log_sig = torch.clamp(input, max=0) - torch.log(1 + torch.exp(-torch.abs(input))) + \
        0.5 * torch.clamp(input, min=0, max=0)
    neg_log_sig = torch.clamp(-input, max=0) - torch.log(1 + torch.exp(-torch.abs(input))) + \
        0.5 * torch.clamp(input, min=0, max=0)
    prob = torch.sigmoid(input)
    pos_weight = torch.pow(1 - prob, gamma)
    neg_weight = torch.pow(prob, gamma)
    loss = -(target * pos_weight * pos_log_sig + (1 - target) * neg_weight * neg_log_sig)
    avg_weight = target * pos_weight + (1 - target) * neg_weight
    loss = loss / avg_weight.mean()
    return loss.mean()


def balanced_loss(input, target, neg_weight=1.0):
    import torch
    pos_mask, neg_mask = (target == 1), (target == 0)
    pos_num, neg_num = pos_mask.sum().float(), neg_mask.sum().float()
    weight = target.new_zeros(target.size())
    weight[pos_mask] = 1 / pos_num
    weight[neg_mask] = 1 / neg_num * neg_weight
    weight /= weight.sum()
    return F.binary_cross_entropy_with_logits(input, target, weight, reduction='sum')


def train_steps(cfg, backbone_sd, head_sd, depth, batches):
    """Runs ``len(batches)`` reference train steps on the CPU; returns (losses, gradients of the first step keyed like
    the head's state dict, head state dict after the last step)."""
    import numpy as np
    import torch
    from .resnet import resnet_forward
    params = {k: v.clone().requires_grad_(True) for k, v in head_sd.items()}
    lr = cfg['initial_lr']
    if cfg['optimizer'] == 'Adam':
        opt = torch.optim.Adam(list(params.values()), lr=lr, weight_decay=0)
    else:
        opt = torch.optim.SGD(list(params.values()), lr=lr, weight_decay=0, momentum=cfg['momentum'])
    mean = torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1)
    std = torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1)
    b = cfg['model']['backbone']
    losses, first_grads = [], None
    for z, x in batches:
        with torch.no_grad():
            fz = resnet_forward(backbone_sd, (z.float() - mean) / std, depth, b['strides'], b['dilations'], b['out_indices'])
            fx = resnet_forward(backbone_sd, (x.float() - mean) / std, depth, b['strides'], b['dilations'], b['out_indices'])
        responses = siam_conv_fc(params, fz, fx, cfg['out_scale'])
        labels = torch.from_numpy(create_labels(tuple(responses.shape), cfg['r_pos'], cfg['r_neg'],
                                                cfg['total_stride'])).float()
        loss = focal_loss(responses, labels) if cfg['loss'] == 'focal' else balanced_loss(responses, labels)
        opt.zero_grad()
        loss.backward()
        if first_grads is None:
            first_grads = {k: v.grad.detach().clone() for k, v in params.items()}
        opt.step()
        losses.append(float(loss))
    return np.asarray(losses), first_grads, {k: v.detach().clone() for k, v in params.items()}
