"""TSN classification head + average consensus (mirrors codes/models/heads/{base,tsn_clshead}.py and
segmental_consensuses/simple_consensus.py:41-61): spatial average pool -> dropout -> Linear -> mean over
the T segments; `fcn_testing` applies the same Linear as a 1x1x1 convolution over (T, h, w) and averages
(tsn_clshead.py:99-117).  State_dict keys: `new_fc.{weight,bias}`."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .builder import HEADS, SEGMENTAL_CONSENSUSES


@SEGMENTAL_CONSENSUSES.register_module
class SimpleConsensus(nn.Module):
    def __init__(self, consensus_type, dim=1):
        super().__init__()
        assert consensus_type in ['avg']
        self.consensus_type, self.dim = consensus_type, dim

    def init_weights(self):
        pass

    def forward(self, x):
        self.shape = x.size()
        return x.mean(dim=self.dim, keepdim=True)


class BaseHead(nn.Module):
    """Shared head state + the cross-entropy loss (heads/base.py:8-45)."""

    def __init__(self, spatial_size=7, dropout_ratio=0.8, in_channels=1024, num_classes=101, init_std=0.001,
                 extract_feat=False):
        super().__init__()
        self.spatial_size = spatial_size if spatial_size == -1 else (spatial_size, spatial_size)
        self.dropout_ratio, self.in_channels, self.num_classes = dropout_ratio, in_channels, num_classes
        self.init_std, self.extract_feat = init_std, extract_feat
        self.dropout = nn.Dropout(p=dropout_ratio) if dropout_ratio != 0 else None
        self.Logits = None

    def init_weights(self):
        pass

    def loss(self, cls_score, labels):
        if labels.shape == torch.Size([]):
            labels = labels.unsqueeze(0)
        return dict(loss_cls=F.cross_entropy(cls_score, labels))


@HEADS.register_module
class TSNClsHead(BaseHead):
    def __init__(self, spatial_type='avg', spatial_size=7, consensus_cfg=dict(type='avg', dim=1),
                 with_avg_pool=False, temporal_feature_size=1, spatial_feature_size=1, dropout_ratio=0.8,
                 in_channels=1024, num_classes=101, init_std=0.001, fcn_testing=False, extract_feat=False):
        super().__init__(spatial_size, dropout_ratio, in_channels, num_classes, init_std, extract_feat)
        if consensus_cfg['type'] != 'avg':
            raise NotImplementedError('only the average consensus of the MVFNet configs is built')
        if spatial_type not in ('avg', 'max'):
            raise ValueError('spatial_type must be avg or max')
        self.spatial_type, self.consensus_type = spatial_type, consensus_cfg['type']
        self.temporal_feature_size, self.spatial_feature_size = temporal_feature_size, spatial_feature_size
        self.cls_pool_size = (temporal_feature_size, spatial_feature_size, spatial_feature_size)
        self.with_avg_pool = with_avg_pool
        self.segmental_consensus = SimpleConsensus(self.consensus_type, consensus_cfg['dim'])
        if self.spatial_size == -1:
            self.pool_size = (1, 1)
            self.Logits = (nn.AdaptiveAvgPool2d if spatial_type == 'avg' else nn.AdaptiveMaxPool2d)(self.pool_size)
        else:
            self.pool_size = self.spatial_size
            self.Logits = (nn.AvgPool2d if spatial_type == 'avg' else nn.MaxPool2d)(self.pool_size, stride=1, padding=0)
        if with_avg_pool:
            self.avg_pool = nn.AvgPool3d(self.cls_pool_size)
        self.new_fc = nn.Linear(in_channels, num_classes)
        self.fcn_testing = fcn_testing
        self.new_cls = None

    def forward(self, x, num_seg):
        if self.fcn_testing:
            # x: (clips, C, T, h, w).  A 1x1x1 conv followed by a mean over (T,h,w) (tsn_clshead.py:99-117)
            # commutes with the mean, so the class map is never materialised.
            feat = x.float().mean([2, 3, 4])
            return feat if self.extract_feat else F.linear(feat, self.new_fc.weight, self.new_fc.bias)
        x = self.Logits(x)
        if x.ndimension() == 4:
            x = x.unsqueeze(2)
        assert x.shape[1] == self.in_channels and x.shape[2] == self.temporal_feature_size
        assert x.shape[3] == self.spatial_feature_size and x.shape[4] == self.spatial_feature_size
        if self.with_avg_pool:
            x = self.avg_pool(x)
        if self.dropout is not None:
            x = self.dropout(x)
        x = x.view(x.size(0), -1)
        score = x if self.extract_feat else self.new_fc(x)
        score = score.reshape((-1, num_seg) + score.shape[1:])
        return self.segmental_consensus(score).squeeze(1)

    def init_weights(self):
        nn.init.normal_(self.new_fc.weight, 0, self.init_std)
        nn.init.constant_(self.new_fc.bias, 0)
