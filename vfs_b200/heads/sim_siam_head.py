"""SimSiam projector/predictor head behind the reference's interface
(mmaction/models/heads/sim_siam_head.py:14-174).  The nn.Modules own the parameters with the reference's
Sequential indices (``projection_fcs.{0,1,3,4,6,7}``, ``predictor_fcs.{0,1,3}``); the arithmetic is native:
global average pool, then one fused Linear(+BatchNorm1d)(+ReLU) kernel per layer (csrc/head.cu) -- M = batch is
8..64 rows, so each layer is bound by streaming its fp32 weight matrix once from HBM, not by FLOPs."""
import torch.nn as nn

from .. import ops
from ..builder import build_drop_layer, build_loss
from ..mmcv_lite import build_norm_layer
from ..registry import HEADS


def build_norm1d(cfg, num_features):
    if cfg['type'] == 'BN':
        return nn.BatchNorm1d(num_features=num_features)
    return build_norm_layer(cfg, num_features=num_features)[1]


@HEADS.register_module()
class SimSiamHead(nn.Module):

    def __init__(self, in_channels, conv_mid_channels=2048, conv_out_channles=2048, num_convs=0, kernel_size=1,
                 conv_cfg=dict(type='Conv2d'), norm_cfg=dict(type='BN'), act_cfg=None, drop_layer_cfg=None,
                 order=('pool', 'drop'), num_projection_fcs=3, projection_mid_channels=2048,
                 projection_out_channels=2048, drop_projection_fc=False, num_predictor_fcs=2,
                 predictor_mid_channels=512, predictor_out_channels=2048, drop_predictor_fc=False, with_norm=True,
                 loss_feat=dict(type='CosineSimLoss', negative=False), spatial_type='avg'):
        super().__init__()
        if num_convs != 0:
            raise NotImplementedError('vfs_b200 SimSiamHead: num_convs > 0 is not used by any VFS config')
        if drop_layer_cfg is not None or drop_projection_fc or drop_predictor_fc:
            raise NotImplementedError('vfs_b200 SimSiamHead: dropout is not used by any VFS config')
        self.in_channels, self.num_convs = in_channels, num_convs
        self.conv_cfg, self.norm_cfg, self.act_cfg, self.with_norm = conv_cfg, norm_cfg, act_cfg, with_norm
        self.loss_feat = build_loss(loss_feat)
        self.convs = nn.Identity()
        last = in_channels
        fcs = []
        for i in range(num_projection_fcs):
            is_last = i == num_projection_fcs - 1
            out = projection_out_channels if is_last else projection_mid_channels
            fcs += [nn.Linear(last, out), build_norm1d(norm_cfg, out)]
            if not is_last:
                fcs.append(nn.ReLU())
            last = out
        self.projection_fcs = nn.Sequential(*fcs) if fcs else nn.Identity()
        fcs = []
        for i in range(num_predictor_fcs):
            is_last = i == num_predictor_fcs - 1
            out = predictor_out_channels if is_last else predictor_mid_channels
            fcs.append(nn.Linear(last, out))
            if not is_last:
                fcs += [build_norm1d(norm_cfg, out), nn.ReLU()]
            last = out
        self.predictor_fcs = nn.Sequential(*fcs) if fcs else nn.Identity()
        assert spatial_type in ['avg', 'att', None]
        self.spatial_type = spatial_type
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1)) if spatial_type == 'avg' else nn.Identity()
        self.dropout = nn.Identity()
        assert set(order) == {'pool', 'drop'}
        self.order = order

    def init_weights(self):
        pass

    @staticmethod
    def _run_mlp(seq, x):
        """Walk a Sequential of Linear / BatchNorm1d / ReLU, fusing each Linear with what follows it."""
        mods = list(seq) if isinstance(seq, nn.Sequential) else []
        i = 0
        while i < len(mods):
            lin = mods[i]
            assert isinstance(lin, nn.Linear), type(lin)
            bn = relu = None
            i += 1
            if i < len(mods) and isinstance(mods[i], nn.modules.batchnorm._BatchNorm):
                bn = mods[i]
                i += 1
            if i < len(mods) and isinstance(mods[i], nn.ReLU):
                relu = mods[i]
                i += 1
            x = ops.linear_bn_act(x, lin, bn, relu is not None)
        return x

    def _pool(self, x):
        if self.spatial_type == 'avg':
            return ops.global_avg_pool(x)
        if x.ndim != 2:
            raise NotImplementedError('vfs_b200 SimSiamHead: spatial_type other than "avg" needs 2-D input')
        return x

    def forward_projection(self, x):
        return self._run_mlp(self.projection_fcs, self._pool(x))

    def forward(self, x):
        """x [B, C, h, w] CUDA fp32 -> (z, p)."""
        z = self._run_mlp(self.projection_fcs, self._pool(x))
        p = self._run_mlp(self.predictor_fcs, z)
        return z, p

    def loss(self, p1, z1, p2, z2, mask12=None, mask21=None, weight=1.):
        assert mask12 is None
        assert mask21 is None
        loss_feat = self.loss_feat(p1, z2.detach()) * 0.5 + self.loss_feat(p2, z1.detach()) * 0.5
        return dict(loss_feat=loss_feat * weight)
