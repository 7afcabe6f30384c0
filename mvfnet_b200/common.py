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


norm_cfg = {'BN': ('bn', nn.BatchNorm2d), 'BN3d': ('bn', nn.BatchNorm3d), 'GN': ('gn', nn.GroupNorm)}


def get_norm_type(cfg):
    assert isinstance(cfg, dict) and 'type' in cfg
    if cfg['type'] not in norm_cfg:
        raise KeyError('Unrecognized norm type {}'.format(cfg['type']))
    return norm_cfg[cfg['type']][1]


def build_norm_layer(cfg, num_features, postfix=''):
    """-> (name, layer): 'bn1', nn.BatchNorm2d(num_features, eps=1e-5) ... (norm.py:28-71)."""
    assert isinstance(cfg, dict) and 'type' in cfg
    cfg_ = cfg.copy()
    layer_type = cfg_.pop('type')
    if layer_type not in norm_cfg:
        raise KeyError('Unrecognized norm type {}'.format(layer_type))
    abbr, norm_layer = norm_cfg[layer_type]
    assert isinstance(postfix, (int, str))
    name = abbr + str(postfix)
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    if layer_type != 'GN':
        layer = norm_layer(num_features, **cfg_)
    else:
        assert 'num_groups' in cfg_
        layer = norm_layer(num_channels=num_features, **cfg_)
    for param in layer.parameters():
        param.requires_grad = requires_grad
    return name, layer
