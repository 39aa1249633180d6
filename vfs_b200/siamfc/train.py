"""SiamFC linear-probe training on the B200 path: ``TrackerSiamFC.train_step`` / ``train_over`` of
projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py:364-467 (+ ``_create_labels`` :468-500, the losses of
siamfc/losses.py:27-64 and the optimiser / LR-schedule setup of :129-168).

The backbone is frozen in this setting (default_config_base.py:40-49: ``frozen_stages=4``, ``norm_eval=True``), so a
step is: backbone features of the exemplar / search batches (inference kernels, no tape) -> the two biased 1x1 adapter
convolutions of ``SiamConvFC`` (tcgen05 conv kernel) -> cross-correlation -> Focal / Balanced loss and its gradient
(one launch) -> correlation gradients -> tcgen05 weight gradients + bias sums -> Adam / SGD update.  The plain
``SiamFC`` head has no parameters (the reference's step then updates nothing)."""
import numpy as np
import torch

from .. import _native as nat
from .. import ops
from .._native import current_stream, ptr


def create_labels(size, r_pos, r_neg, total_stride):
    """``_create_labels`` (siamfc_tracker_base.py:468-500): 1 inside the L1 ball of radius r_pos / stride around the
    centre, 0.5 inside r_neg / stride, 0 elsewhere; float32 [n, c, h, w] (numpy)."""
    n, c, h, w = size
    x = np.arange(w) - (w - 1) / 2
    y = np.arange(h) - (h - 1) / 2
    x, y = np.meshgrid(x, y)
    dist = np.abs(x) + np.abs(y)
    rp, rn = r_pos / total_stride, r_neg / total_stride
    labels = np.where(dist <= rp, np.ones_like(x), np.where(dist < rn, np.ones_like(x) * 0.5, np.zeros_like(x)))
    return np.tile(labels.reshape((1, 1, h, w)), (n, c, 1, 1)).astype(np.float32)


def siamfc_loss(responses, labels, kind='focal', gamma=2.0, neg_weight=1.0, want_grad=True):
    """(loss [1], d loss / d responses | None) of FocalLoss (``kind='focal'``) or BalancedLoss (``'balance'``)."""
    r = responses.contiguous().float()
    t = labels.contiguous().float()
    assert r.is_cuda and r.shape == t.shape
    loss = torch.empty((1, ), dtype=torch.float32, device=r.device)
    grad = torch.empty_like(r) if want_grad else None
    mode = {'focal': 0, 'balance': 1}[kind]
    ops.check(nat.lib().vfs_siamfc_loss(ptr(r), ptr(t), ptr(loss), ptr(grad), r.numel(), mode, float(gamma),
                                        float(neg_weight), current_stream()), 'siamfc_loss')
    return loss, grad


def xcorr_backward(dr, z_nhwc, x_nhwc, out_scale):
    """Gradients of ops.xcorr_nhwc for paired batches: dr [n,1,ho,wo] -> (dz [n,hz,wz,C], dx [n,h,w,C])."""
    n, hz, wz, C = z_nhwc.shape
    n2, h, w, C2 = x_nhwc.shape
    if n != n2:
        raise NotImplementedError('vfs_b200 SiamFC training pairs exemplar i with search i (nz == nx)')
    assert C == C2 and tuple(dr.shape) == (n, 1, h - hz + 1, w - wz + 1)
    dz = torch.empty_like(z_nhwc)
    dx = torch.empty_like(x_nhwc)
    ops.check(nat.lib().vfs_xcorr_backward_nhwc(ptr(dr.contiguous()), ptr(z_nhwc), ptr(x_nhwc), ptr(dz), ptr(dx), n, C, hz,
                                                wz, h, w, float(out_scale), current_stream()), 'xcorr_backward')
    ops.LAUNCHES[0] += 1
    return dz, dx


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam's constructor and update rule (no amsgrad) with one fused kernel per parameter."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or not p.is_contiguous() or p.dtype != torch.float32:
                    raise RuntimeError('vfs_b200 Adam needs contiguous float32 CUDA parameters (no CPU fallback)')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                ops.check(nat.lib().vfs_adam_step(ptr(p.data), ptr(p.grad.contiguous()), ptr(st['exp_avg']),
                                                  ptr(st['exp_avg_sq']), p.numel(), float(group['lr']),
                                                  float(group['betas'][0]), float(group['betas'][1]),
                                                  float(group['eps']), float(group['weight_decay']), int(st['step']),
                                                  current_stream()), 'adam_step')
                p.add_(0)      # version bump for cached packed weights
        return loss


class TrainingMixin:
    """Training half of ``TrackerSiamFC`` (the inference half lives in tracker.py)."""

    # ------------------------------------------------------------------ setup (siamfc_tracker_base.py:122-168)
    def setup_training(self):
        cfg = self.cfg
        self.loss_kind = cfg.get('loss', 'focal')
        if self.loss_kind not in ('focal', 'balance'):
            raise NotImplementedError(f'loss {self.loss_kind}')
        params = [p for p in self.net.parameters() if p.requires_grad]
        frozen = cfg.model.backbone.get('frozen_stages', -1)
        wd = cfg.get('weight_decay', 5e-4) if (frozen < 4 or cfg.get('force_wd', False)) else 0
        kind = cfg.get('optimizer', 'Adam')
        lr = cfg.get('initial_lr', 1e-3)
        if not params:
            self.optimizer = None
        elif kind == 'Adam':
            self.optimizer = Adam(params, lr=lr, weight_decay=wd)
        elif kind == 'SGD':
            from ..optim import SGD
            self.optimizer = SGD(params, lr=lr, weight_decay=wd, momentum=cfg.get('momentum', 0.9))
        else:
            raise NotImplementedError(kind)
        self.lr_scheduler = None
        if self.optimizer is not None:
            sched = cfg.get('lr_schedule', 'exp')
            if sched == 'exp':
                gamma = np.power(cfg.get('ultimate_lr', 1e-5) / lr, 1.0 / cfg.get('epoch_num', 50))
                self.lr_scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma)
            elif sched == 'step':
                self.lr_scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, cfg.get('lr_step_size', 10))
            elif sched != 'fixed':
                raise NotImplementedError(sched)
        self._labels = None

    def _create_labels(self, size):
        if self._labels is None or tuple(self._labels.shape) != tuple(size):
            cfg = self.cfg
            self._labels = torch.from_numpy(create_labels(tuple(size), cfg.get('r_pos', 16), cfg.get('r_neg', 0),
                                                          cfg.total_stride)).to(self.device)
        return self._labels

    # ------------------------------------------------------------------ head forward with what the backward needs
    def _head_forward_train(self, fz, fx):
        head = self.net.head
        convs_z, convs_x = list(getattr(head, 'z_convs', [])), list(getattr(head, 'x_convs', []))
        if len(convs_z) > 1 or len(convs_x) > 1:
            raise NotImplementedError('vfs_b200 SiamFC training supports num_convs <= 1 (the reference default)')
        ctx = {}
        branches = []
        for name, feat, convs in (('z', fz, convs_z), ('x', fx, convs_x)):
            if convs:
                m = convs[0]
                assert m.kernel_size == (1, 1) and m.stride == (1, 1)
                xs = ops.to_split(feat.contiguous().float())
                w = ops.pack_conv_weight(m.weight.detach().float().contiguous(), ops.WEIGHT_SCALE)
                ones = ops._const_vec(1.0 / ops.WEIGHT_SCALE, m.out_channels, feat.device)
                shift = m.bias.detach().float().contiguous() if m.bias is not None else \
                    ops._const_vec(0, m.out_channels, feat.device)
                _, a = ops.conv_bn_act(xs, w, ones, shift, 1, 1, 1, relu=False, want_split=False, want_f32=True)
                ctx[name] = (m, xs)
            else:
                a = ops.nchw_to_nhwc(feat)
            branches.append(a)
        ctx['a_z'], ctx['a_x'] = branches
        return ops.xcorr_nhwc(branches[0], branches[1], head.out_scale), ctx

    def _head_backward(self, ctx, dr):
        head = self.net.head
        dz, dx = xcorr_backward(dr, ctx['a_z'], ctx['a_x'], head.out_scale)
        for name, da in (('z', dz), ('x', dx)):
            if name not in ctx:
                continue
            m, xs = ctx[name]
            # the adapter gradients are ~1e-6: scale by a power of two into the range where both fp16 planes of the
            # tensor-core operand are normal numbers, undo it in the weight-gradient epilogue
            amax = float(da.abs().max())
            scale = 2.0 ** np.floor(np.log2(1024.0 / amax)) if amax > 0 else 1.0
            C = m.out_channels
            da_split = ops.bn_apply(da, torch.full((C, ), scale, dtype=torch.float32, device=da.device),
                                    ops._const_vec(0, C, da.device), None, relu=False)
            m.weight.grad = ops.conv_wgrad(xs, da_split, 1, 1, 1, out_scale=1.0 / scale)
            if m.bias is not None:
                m.bias.grad = ops.channel_stats(da)[:C].float()

    # ------------------------------------------------------------------ reference API
    def train_step(self, batch, backward=True):
        """``batch = (z [B,3,ez,ez], x [B,3,sz,sz])`` uint8/float RGB crops (NCHW) -> python float loss; with
        ``backward`` the head parameters are updated (siamfc_tracker_base.py:364-386)."""
        if not hasattr(self, 'optimizer'):
            self.setup_training()
        self.net.train(backward)
        z = batch[0].to(self.device, non_blocking=True).float()
        x = batch[1].to(self.device, non_blocking=True).float()
        with torch.no_grad():
            bb = self.net.backbone
            if any(p.requires_grad for p in bb.parameters()):
                raise NotImplementedError('vfs_b200 SiamFC training: the backbone must be frozen (frozen_stages=4, the '
                                          'reference default); end-to-end fine-tuning goes through SimSiamBaseTracker-'
                                          'style training')
            fz = bb(self.normalize(z))
            fx = bb(self.normalize(x))
            responses, ctx = self._head_forward_train(fz, fx)
            labels = self._create_labels(responses.size())
            loss, dr = siamfc_loss(responses, labels, self.loss_kind, want_grad=backward)
            if backward and self.optimizer is not None:
                self.optimizer.zero_grad()
                self._head_backward(ctx, dr)
                self.optimizer.step()
        return float(loss.item())

    def train_over(self, loader, epochs=None, log=None):
        """Loop of ``train_over`` (:388-467) over any iterable of ``(z, x)`` batches (the reference builds a GOT-10k
        ``Pair`` dataset + DataLoader, host-side data code outside this path); steps the LR schedule per epoch and
        returns the per-epoch mean losses."""
        if not hasattr(self, 'optimizer'):
            self.setup_training()
        history = []
        for epoch in range(epochs if epochs is not None else self.cfg.get('epoch_num', 50)):
            losses = [self.train_step(batch, backward=True) for batch in loader]
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            history.append(float(np.mean(losses)) if losses else float('nan'))
            if log is not None:
                log(f'Epoch: {epoch + 1} loss {history[-1]:.5f}')
        return history
