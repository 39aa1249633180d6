"""Host-side executor of the backbone: owns the device-side plan of a ``ResNet`` module (packed split-fp16
weights, folded BN scale/shift) and walks the block list issuing one fused kernel per ConvModule through the
C ABI.  It never computes anything in torch: tensors here are only device buffers handed to libvfs_b200.so.

Train-mode BatchNorm (batch statistics, SyncBN exchange) runs as conv+stats -> finalise -> normalise kernels.
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm

from . import ops
from .mmcv_lite import ConvModule


class _ConvPlan:
    __slots__ = ('w_split', 'wt_split', 'scale', 'shift', 'ksize', 'w_version', 'bn_version')


def fold_bn(conv_module, device):
    """Eval-mode BN folded to y = conv(x) * scale + shift (computed in fp64, stored fp32)."""
    cout = conv_module.conv.out_channels
    if conv_module.with_norm:
        bn = conv_module.norm
        inv = 1.0 / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        g = bn.weight.detach().double() if bn.weight is not None else torch.ones(cout, dtype=torch.float64)
        b = bn.bias.detach().double() if bn.bias is not None else torch.zeros(cout, dtype=torch.float64)
        scale = g.to(inv.device) * inv
        shift = b.to(inv.device) - bn.running_mean.detach().double() * scale
    else:
        scale = torch.ones(cout, dtype=torch.float64)
        shift = torch.zeros(cout, dtype=torch.float64)
    if conv_module.conv.bias is not None:
        shift = shift + conv_module.conv.bias.detach().double().to(shift.device) * scale
    if conv_module.conv.in_channels % 64 == 0:
        scale = scale / ops.WEIGHT_SCALE      # residual-stage weights are packed scaled by WEIGHT_SCALE (exact: 2^-8)
    return (scale.float().to(device).contiguous(), shift.float().to(device).contiguous())


class BackboneEngine:

    def __init__(self, resnet):
        self.net = resnet
        self._plans = {}
        self.check_versions = True  # re-pack when parameters were modified in place / reloaded
        self.events = None          # bench instrumentation: list collecting (tag, cuda event) at phase boundaries
        self.tape = None            # training: list recording what backward needs (set by the autograd wrapper)
        self._train_pack = None     # persistent packed weights + descriptor table of the multi-tensor pack
        self._graphs = {}           # (input shape, stage, normalize, geometry) -> captured CUDA graph (LRU order)
        self.max_graphs = 8
        self._convs = None
        self._tensors = None
        self._arena = None          # fp64 statistics arena of the running training pass (forward or backward)
        self._arena_off = 0
        self._arena_size = None
        self._nbt = None            # num_batches_tracked tensors to bump at the end of the pass
        self._region = None         # (requires_grad signature, ids of ConvModules inside the gradient region)

    # -------------------------------------------------------------- plans
    @staticmethod
    def _w_version(cm):
        """Validity stamp of the packed weight operands (conv weight only)."""
        return cm.conv.weight._version + cm.conv.weight.data_ptr() + (ops.WEIGHT_EPOCH[0] << 20)

    @staticmethod
    def _bn_version(cm):
        """Validity stamp of the folded eval-mode scale / shift (BN parameters and running statistics, conv bias)."""
        v = ops.WEIGHT_EPOCH[0] << 20
        if cm.conv.bias is not None:
            v += cm.conv.bias._version
        if cm.with_norm:
            bn = cm.norm
            for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked):
                if t is not None:
                    v += t._version + t.data_ptr()
        return v

    @classmethod
    def _version(cls, cm):
        return cls._w_version(cm) + cls._bn_version(cm)

    def plan(self, cm, device):
        """Device-side plan of one ConvModule: packed split-fp16 weights (re-packed when the conv weight changed) and
        the folded eval-mode BN (re-folded when BN tensors changed; skipped while the BN is in batch-statistics mode,
        whose running-stat updates must not invalidate the weight pack between the two views of a training step)."""
        p = self._plans.get(id(cm))
        check = self.check_versions or p is None
        wv = self._w_version(cm) if check else None
        if p is None or (wv is not None and p.w_version != wv) or p.w_split.device != device:
            p = _ConvPlan()
            w = cm.conv.weight.detach().to(device=device, dtype=torch.float32).contiguous()
            p.ksize = w.shape[2]
            # residual-stage convs: [2][Cout][k*k*Cin]; the 7x7 stem: [2][64][192] (K = 147 zero-padded)
            p.w_split = ops.pack_conv_weight(w, ops.WEIGHT_SCALE) if w.shape[1] % 64 == 0 else ops.stem_pack_weight(w)
            p.wt_split = None  # dgrad packing, built on first use by the backward pass
            p.scale = p.shift = None
            p.w_version = wv if wv is not None else self._w_version(cm)
            p.bn_version = None
            self._plans[id(cm)] = p
        train_bn = cm.with_norm and cm.norm.training
        if not train_bn:
            bv = self._bn_version(cm) if (check or p.bn_version is None) else p.bn_version
            if p.scale is None or p.bn_version != bv:
                # (~12 tiny fp64 kernels; a train-mode step never pays for it)
                p.scale, p.shift = fold_bn(cm, device)
                p.bn_version = bv
        return p

    def invalidate(self):
        self._plans.clear()
        self._graphs.clear()
        if self._train_pack is not None:
            self._train_pack['stamp'] = None     # buffers and descriptor table stay (no H2D copy inside a capture)

    def _refresh_train_weights(self, device):
        """Training entry: (re)pack the weights of EVERY residual-stage conv -- forward and data-gradient operand
        layouts -- with ONE launch into persistent buffers (vfs_pack_conv_weights_multi) when any parameter changed
        since the last pack; ``plan()`` then finds every plan current.  Replaces ~210 per-tensor pack launches per
        training step."""
        import ctypes
        from . import _native as nat
        tp = self._train_pack
        # conv weights only: BN running statistics change between the two views of a step without touching the operands
        stamp = (ops.WEIGHT_EPOCH[0] << 20) + (sum(cm.conv.weight._version for cm in tp['cms']) if tp else 0)
        if tp is not None and tp['stamp'] == stamp and tp['device'] == device:
            return
        if tp is None or tp['device'] != device or tp['ptrs'] != [cm.conv.weight.data_ptr() for cm in tp['cms']]:
            cms = [m for m in self.net.modules() if isinstance(m, ConvModule) and m.conv.in_channels % 64 == 0]
            items, bufs, first = [], [], 0
            for cm in cms:
                w = cm.conv.weight
                if w.device != device or w.dtype != torch.float32 or not w.is_contiguous():
                    raise RuntimeError('vfs_b200 training needs contiguous float32 CUDA conv weights')
                cout, cin, k, _ = w.shape
                ws = torch.empty((2, cout, k * k * cin), dtype=torch.float16, device=device)
                wt = torch.empty((2, cin, k * k * cout), dtype=torch.float16, device=device)
                bufs.append((ws, wt))
                nb = int(nat.lib().vfs_pack_blocks(cout, cin, k))
                for mode, dst in ((0, ws), (1, wt)):
                    items.append(nat.VfsPackItem(w=w.data_ptr(), dst_split=dst.data_ptr(), Cout=cout, Cin=cin, ksize=k,
                                                 mode=mode, first_block=first, scale_log2=ops.WEIGHT_SCALE_LOG2))
                    first += nb
            arr = (nat.VfsPackItem * len(items))(*items)
            table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
            tp = dict(cms=cms, bufs=bufs, table=table, n=len(items), blocks=first, device=device,
                      ptrs=[cm.conv.weight.data_ptr() for cm in cms], stamp=None)
            self._train_pack = tp
        ops.check(nat.lib().vfs_pack_conv_weights_multi(nat.ptr(tp['table']), tp['n'], tp['blocks'],
                                                         nat.current_stream()), 'pack_conv_weights_multi')
        for cm, (ws, wt) in zip(tp['cms'], tp['bufs']):
            p = self._plans.get(id(cm))
            if p is None:
                p = _ConvPlan()
                p.scale = p.shift = None
                self._plans[id(cm)] = p
                p.bn_version = None
            p.ksize = cm.conv.kernel_size[0]
            p.w_split, p.wt_split = ws, wt
            p.w_version = self._w_version(cm)
        tp['stamp'] = (ops.WEIGHT_EPOCH[0] << 20) + sum(cm.conv.weight._version for cm in tp['cms'])

    def _stamp(self):
        """Cheap global version stamp of every parameter/buffer (in-place updates and reloads bump it)."""
        if self._tensors is None:
            self._tensors = list(self.net.parameters()) + list(self.net.buffers())
        return sum(t._version for t in self._tensors) + (ops.WEIGHT_EPOCH[0] << 20)

    # -------------------------------------------------------------- single fused layer
    def conv(self, cm, xs, relu, residual=None, want_f32=False):
        """One ConvModule (+residual, +ReLU) on a split NHWC tensor."""
        assert isinstance(cm, ConvModule)
        k = cm.conv.kernel_size[0]
        stride, dil, pad = cm.conv.stride[0], cm.conv.dilation[0], cm.conv.padding[0]
        assert cm.conv.kernel_size[0] == cm.conv.kernel_size[1] and cm.conv.stride[0] == cm.conv.stride[1]
        assert cm.conv.groups == 1
        assert pad == (0 if k == 1 else dil), f'unsupported padding {pad} for k={k}, dilation={dil}'
        p = self.plan(cm, xs.device)
        if cm.with_norm and cm.norm.training:
            # batch statistics: conv (+ per-channel sums in the epilogue) -> finalise (+ SyncBN all-reduce,
            # running-stat update) -> normalise + residual + ReLU
            assert not want_f32
            # (z stays a split tensor: written by the TMA epilogue whose math warps also accumulate the statistics)
            z, stats = ops.conv_stats_split(xs, p.w_split, k, stride, dil,
                                            stats=self.stats_vec(2 * cm.conv.out_channels, xs.device),
                                            wscale=ops.WEIGHT_SCALE)
            scale, shift, mean, invstd = ops.bn_finalize(stats, z.numel() // (2 * z.shape[-1]), cm.norm,
                                                         nbt_list=self._nbt)
            y = ops.bn_apply(z, scale, shift, residual, relu)
            if self.tape is not None:
                self.tape.append(dict(cm=cm, xs=xs, z=z, mean=mean, invstd=invstd, y=y, relu=relu,
                                      residual=residual, k=k, stride=stride, dil=dil))
            return y
        if p.scale is None:   # planned while the BN was in train mode
            p.scale, p.shift = fold_bn(cm, xs.device)
        if self.tape is not None and cm.with_norm and not want_f32 and id(cm) in self._grad_region():
            # Eval-mode BatchNorm inside a taped (training) forward -- norm_eval / frozen-BN fine-tuning, reference
            # resnet.py:645-654: the raw conv output is kept (the BN backward needs xhat, and dgamma cannot be
            # recovered from y when gamma == 0, the zero-init-residual state), BN on the running statistics is the folded
            # scale / shift.  Gradients: dz = gamma * invstd_running * g, dgamma / dbeta from the running-statistics xhat.
            bn = cm.norm
            z, _ = ops.conv_bn_act(xs, p.w_split, ops._const_vec(1.0 / ops.WEIGHT_SCALE, cm.conv.out_channels, xs.device),
                                   ops._const_vec(0, cm.conv.out_channels, xs.device), k, stride, dil, relu=False)
            # (p.scale carries 1 / WEIGHT_SCALE for the fused kernel; z here is the true conv output)
            y = ops.bn_apply(z, p.scale * ops.WEIGHT_SCALE, p.shift, residual, relu)
            invstd = torch.rsqrt(bn.running_var.detach().double() + bn.eps).float()
            self.tape.append(dict(cm=cm, xs=xs, z=z, mean=bn.running_mean.detach().float(), invstd=invstd, y=y,
                                  relu=relu, residual=residual, k=k, stride=stride, dil=dil, eval_bn=True))
            return y
        out, out32 = ops.conv_bn_act(xs, p.w_split, p.scale, p.shift, k, stride, dil, relu, residual,
                                     want_split=not want_f32, want_f32=want_f32)
        return out32 if want_f32 else out

    def stem(self, x):
        cm = self.net.conv1
        assert cm.conv.in_channels == 3 and cm.conv.kernel_size == (7, 7) and cm.conv.stride == (2, 2)
        p = self.plan(cm, x.device)
        if cm.with_norm and cm.norm.training:
            z = ops.stem_conv_raw(x, p.w_split)
            stats = ops.channel_stats(z, stats=self.stats_vec(128, x.device))
            scale, shift, mean, invstd = ops.bn_finalize(stats, z.numel() // 64, cm.norm, nbt_list=self._nbt)
            y = ops.stem_bn_relu_pool(z, scale, shift, x.shape[2:])
            if self.tape is not None:
                self.tape.append(dict(cm=cm, stem=True, x=x, z=z, mean=mean, invstd=invstd, scale=scale,
                                      shift=shift, y=y))
            return y
        if p.scale is None:
            p.scale, p.shift = fold_bn(cm, x.device)
        return ops.stem_forward(x, p.w_split, p.scale, p.shift)

    # -------------------------------------------------------------- whole backbone
    def forward(self, x, out_indices=None, block_index=None):
        """Runs stem + residual stages.  Returns a list of NCHW fp32 tensors: the outputs of the stages in
        ``out_indices`` (reference ResNet.forward, resnet.py:555-575) or of block ``block_index``
        (forward_block :577-587).  Stages after the last requested one are skipped -- the reference computes
        and discards them, the results are identical."""
        if not x.is_cuda:
            raise RuntimeError('vfs_b200.ResNet needs a CUDA tensor: the B200 path has no CPU fallback '
                               '(the CPU restatement lives in oracle/ and is test infrastructure only)')
        x = x.contiguous().float()
        net = self.net
        xs = self.stem(x)
        outs = []
        last_stage = max(out_indices) if out_indices is not None else len(net.res_layers) - 1
        bidx = 0
        for i, name in enumerate(net.res_layers):
            if i > last_stage and not net.training:
                break  # in train mode the reference's later stages still update their BN running statistics
            for block in getattr(net, name):
                xs = block.native_forward(self, xs)
                if block_index is not None and bidx == block_index:
                    return [ops.from_split(xs)]
                bidx += 1
            if out_indices is not None and i in out_indices:
                outs.append(ops.from_split(xs))
        if block_index is not None:
            return [None]
        return outs

    # -------------------------------------------------------------- training: taped forward + native backward
    def forward_taped(self, x, out_indices):
        """Like ``forward`` but also returns the split tensors behind the outputs (identity keys of the gradient
        bookkeeping).  ``self.tape`` must be a list; train-mode ConvModules append what their backward needs."""
        assert self.tape is not None
        x = x.contiguous().float()
        self._refresh_train_weights(x.device)
        self._begin_stats_arena(x.device)
        try:
            xs = self.stem(x)
            outs, out_splits = [], []
            for i, name in enumerate(self.net.res_layers):
                for block in getattr(self.net, name):
                    xs = block.native_forward(self, xs)
                if i in out_indices:
                    outs.append(ops.from_split(xs))
                    out_splits.append(xs)
        finally:
            self._end_stats_arena()
        return outs, out_splits

    def _grad_region(self):
        """ids of the ConvModules a gradient has to flow through: everything from the first residual block that owns a
        trainable parameter onwards (frozen_stages freezes a prefix of the network, resnet.py:593-609; layers in front of
        the first trainable one keep the fused inference kernel and stay off the tape)."""
        key = tuple(p.requires_grad for p in self.net.parameters())
        if self._region is None or self._region[0] != key:
            ids, started = set(), False
            for name in self.net.res_layers:
                for block in getattr(self.net, name):
                    started = started or any(p.requires_grad for p in block.parameters())
                    if started:
                        ids.update(id(m) for m in block.modules() if isinstance(m, ConvModule))
            self._region = (key, ids)
        return self._region[1]

    # -- per-pass arena of zero-initialised fp64 statistics vectors: one fill per pass instead of one per BN layer
    def _arena_total(self):
        if self._arena_size is None:
            self._arena_size = sum(2 * m.num_features for m in self.net.modules() if isinstance(m, _BatchNorm))
        return self._arena_size

    def _begin_stats_arena(self, device):
        self._arena = torch.zeros((self._arena_total(), ), dtype=torch.float64, device=device)
        self._arena_off = 0
        self._nbt = []

    def _end_stats_arena(self):
        self._arena = None
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)      # one launch for every num_batches_tracked of the pass
        self._nbt = None

    def stats_vec(self, n, device):
        """Zeroed fp64 [n] for per-channel sums: a slice of the pass's arena when one is open."""
        if self._arena is not None and self._arena_off + n <= self._arena.numel():
            v = self._arena[self._arena_off:self._arena_off + n]
            self._arena_off += n
            return v
        return torch.zeros((n, ), dtype=torch.float64, device=device)

    def backward(self, tape, out_splits, grad_outs):
        """Reverse walk over the tape.  Gradients of activations are split tensors scaled by autograd.GRAD_SCALE;
        returns {id(parameter): fp32 gradient}.  Per ConvModule: BN(+ReLU) backward (two per-channel sums, all-reduced
        for SyncBN) -> tcgen05 wgrad -> tcgen05 dgrad (+ the gradient already accumulated for that input)."""
        from .autograd import GRAD_SCALE as S
        grad = {}
        for xs, g in zip(out_splits, grad_outs):
            if g is None:
                continue
            gs = ops.to_split_scaled(g.contiguous().float(), S)
            if id(xs) in grad:
                raise NotImplementedError('vfs_b200: the same stage output was requested twice')
            grad[id(xs)] = gs
        pgrads = {}
        inv = 1.0 / S
        if tape:
            self._begin_stats_arena(tape[0]['z'].device)
        try:
            return self._backward_walk(tape, grad, pgrads, inv)
        finally:
            self._end_stats_arena()

    def _backward_walk(self, tape, grad, pgrads, inv):
        for op in reversed(tape):
            dy = grad.pop(id(op['y']), None)
            if dy is None:
                continue  # not on any gradient path (e.g. stages after the last used output)
            cm = op['cm']
            bn = cm.norm
            if op.get('stem'):
                g32 = ops.stem_pool_relu_backward(dy, op['z'], op['scale'], op['shift'], op['x'].shape[2:])
                dz32, _, dgam, dbet = ops.bn_backward(g32, None, op['z'], op['mean'], op['invstd'], bn, dy_is_f32=True,
                                                      want_f32=True, param_scale=inv,
                                                      sums=self.stats_vec(128, g32.device))
                if cm.conv.weight.requires_grad:
                    sink = ops.grad_sink(cm.conv.weight)
                    dw = ops.stem_wgrad(op['x'], dz32, out_scale=inv, out=sink)
                    if sink is None:
                        pgrads[id(cm.conv.weight)] = dw
                if bn.affine and bn.weight.requires_grad and dgam is not None:
                    pgrads[id(bn.weight)], pgrads[id(bn.bias)] = dgam, dbet
                continue
            want_g = op['residual'] is not None
            dz, g, dgam, dbet = ops.bn_backward(dy, op['y'] if op['relu'] else None, op['z'], op['mean'], op['invstd'],
                                                bn, want_g=want_g, param_scale=inv,
                                                sums=self.stats_vec(2 * op['z'].shape[-1], dy.device),
                                                eval_mode=op.get('eval_bn', False))
            if want_g:
                if id(op['residual']) in grad:
                    raise NotImplementedError('vfs_b200: unexpected second gradient for a residual input')
                grad[id(op['residual'])] = g
            if bn.affine and bn.weight.requires_grad and dgam is not None:
                pgrads[id(bn.weight)], pgrads[id(bn.bias)] = dgam, dbet
            if cm.conv.weight.requires_grad:
                sink = ops.grad_sink(cm.conv.weight)   # flat gradient view: accumulate in place (dp.FlatTrainState)
                dw = ops.conv_wgrad(op['xs'], dz, op['k'], op['stride'], op['dil'], out=sink,
                                    accumulate=sink is not None, out_scale=inv)
                if sink is None:
                    pgrads[id(cm.conv.weight)] = dw
            plan = self.plan(cm, dz.device)
            if plan.wt_split is None:
                plan.wt_split = ops.pack_conv_weight_dgrad(
                    cm.conv.weight.detach().to(device=dz.device, dtype=torch.float32).contiguous(), ops.WEIGHT_SCALE)
            prev = grad.get(id(op['xs']))
            grad[id(op['xs'])] = ops.conv_dgrad(dz, plan.wt_split, tuple(op['xs'].shape[2:4]), op['k'], op['stride'],
                                                op['dil'], add=prev, wscale=ops.WEIGHT_SCALE)
        return pgrads

    def _mark(self, tag):
        if self.events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.events.append((tag, ev))

    def forward_split(self, x, stage):
        """Features of stage ``stage`` kept in the library's split NHWC format (for the on-device tracker)."""
        if not x.is_cuda:
            raise RuntimeError('vfs_b200 backbone needs a CUDA tensor: the B200 path has no CPU fallback')
        self._mark('stem_begin')
        xs = self.stem(x.contiguous().float())
        self._mark('convs_begin')
        xs = self.run_stages(xs, stage)
        self._mark('convs_end')
        return xs

    def forward_split_taps(self, x, stages, all_blocks=False):
        """Split-format features of several tap points from one pass: the outputs of the stages in ``stages``, or
        with ``all_blocks`` of every residual block inside those stages (VanillaTracker.extract_feat_test,
        vanilla_tracker.py:30-46)."""
        if not x.is_cuda:
            raise RuntimeError('vfs_b200 backbone needs a CUDA tensor: the B200 path has no CPU fallback')
        xs = self.stem(x.contiguous().float())
        outs = []
        for i, name in enumerate(self.net.res_layers):
            if i > max(stages):
                break
            for block in getattr(self.net, name):
                xs = block.native_forward(self, xs)
                if all_blocks and i in stages:
                    outs.append(xs)
            if not all_blocks and i in stages:
                outs.append(xs)
        return outs

    def run_stages(self, xs, stage):
        """Residual stages 0..``stage`` on a split NHWC tensor (all tcgen05 conv launches)."""
        for i, name in enumerate(self.net.res_layers):
            if i > stage:
                break
            for block in getattr(self.net, name):
                xs = block.native_forward(self, xs)
        return xs

    def features_graphed(self, x, stage, normalize=False):
        """forward_split (+ optional L2 normalisation over channels) replayed from a CUDA graph captured once per
        input shape over static buffers: one graph launch instead of ~45 Python/ctypes kernel launches.  The
        returned split tensor is the graph's static output (valid until the next call with the same shape)."""
        # geometry (switch_strides / StrideContext / change_stride rewrite conv.stride without touching any tensor)
        # is part of the key; parameter / buffer updates are caught by the stamp (tensor versions + ops.WEIGHT_EPOCH,
        # which native optimiser steps and training-graph replays bump)
        if self._convs is None:
            self._convs = [m for m in self.net.modules() if isinstance(m, torch.nn.Conv2d)]
        geom = tuple((m.stride[0], m.dilation[0], m.padding[0]) for m in self._convs)
        key = (tuple(x.shape), int(stage), bool(normalize), x.device.index, geom)
        stamp = self._stamp()
        ent = self._graphs.get(key)
        if ent is not None and ent[3] != stamp:
            self._plans.clear()
            self._graphs.clear()
            ent = None
        if ent is None:
            static_in = x.contiguous().float().clone()
            saved_events, self.events = self.events, None
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):               # eager warm-up: builds plans, sets kernel attributes
                xs = self.forward_split(static_in, stage)
                if normalize:
                    xs = ops.normalize_split(xs)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                xs = self.forward_split(static_in, stage)
                if normalize:
                    xs = ops.normalize_split(xs)
            self.events = saved_events
            ent = (graph, static_in, xs, stamp)
            self._graphs[key] = ent
            while len(self._graphs) > self.max_graphs:          # LRU bound: each entry pins its activation pool
                self._graphs.pop(next(iter(self._graphs)))
        else:
            self._graphs[key] = self._graphs.pop(key)           # mark as most recently used
        graph, static_in, xs, _ = ent
        static_in.copy_(x, non_blocking=True)
        graph.replay()
        return xs

    def conv_layer_list(self, in_shape, stage):
        """Static (Cin, Cout, k, stride, dil, H, W) list of the tcgen05 conv launches of forward_split for an
        NCHW input shape -- used by bench.py to count algorithmic FLOPs."""
        from .ops import conv_out_hw
        N, _, H, W = in_shape
        H, W = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        H, W = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        out = []

        def add(cm, h, w):
            k, s, d = cm.conv.kernel_size[0], cm.conv.stride[0], cm.conv.dilation[0]
            ho, wo = conv_out_hw(h, w, k, s, d)
            out.append(dict(Cin=cm.conv.in_channels, Cout=cm.conv.out_channels, k=k, stride=s, dil=d, H=h, W=w,
                            Ho=ho, Wo=wo, N=N, flops=2.0 * N * ho * wo * cm.conv.out_channels *
                            cm.conv.in_channels * k * k))
            return ho, wo

        for i, name in enumerate(self.net.res_layers):
            if i > stage:
                break
            for block in getattr(self.net, name):
                if block.downsample is not None:
                    add(block.downsample, H, W)
                convs = [block.conv1, block.conv2] + ([block.conv3] if hasattr(block, 'conv3') else [])
                h, w = H, W
                for cm in convs:
                    h, w = add(cm, h, w)
                H, W = h, w
        return out
