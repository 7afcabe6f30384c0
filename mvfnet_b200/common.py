"""Small shared layers (mirrors codes/models/common/se_module.py:5-24 and common/norm.py:4-71)."""
import torch.nn as nn
import torch.nn.functional as F


class HardSigmoid(nn.Module):
    """relu6(x + 3) / 6 (se_module.py:5-13).  Inside MVF this is fused into the CUDA kernel."""

    def __init__(self, inplace=True):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return F.relu6(x + 3., inplace=self.inplace) / 6.


class HardSwish(nn.Module):
    """x * hardsigmoid(x) (se_module.py:16-24)."""

    def __init__(self, inplace=True):
        super().__init__()
        self.sigmoid = HardSigmoid(inplace=inplace)

    def forward(self, x):
        return x * self.sigmoid(x)


# the MVFNet configs build every norm layer as dict(type='BN', requires_grad=True) (r50_dense.py:27)
norm_cfg = {'BN': ('bn', nn.BatchNorm2d), 'BN3d': ('bn', nn.BatchNorm3d)}


def get_norm_type(cfg):
    if cfg['type'] not in norm_cfg:
        raise KeyError('Unrecognized norm type {}'.format(cfg['type']))
    return norm_cfg[cfg['type']][1]


def build_norm_layer(cfg, num_features, postfix=''):
    """-> ('bn<postfix>', layer): the sub-module names `bn1..3` are part of the state_dict contract (norm.py:28-71);
    eps defaults to 1e-5, `requires_grad` freezes gamma / beta."""
    kw = dict(cfg)
    abbr, cls = norm_cfg[kw.pop('type')] if kw.get('type') in norm_cfg else (None, None)
    if cls is None:
        raise KeyError('Unrecognized norm type {}'.format(cfg.get('type')))
    requires_grad = kw.pop('requires_grad', True)
    kw.setdefault('eps', 1e-5)
    layer = cls(num_features, **kw)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer
