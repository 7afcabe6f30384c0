"""2D ResNet-50/101/152 backbone whose bottlenecks host the MVF module.

Host-side mirror of the part of codes/models/backbones/resnet.py the MVFNet configs use
(`Bottleneck` :104-244, `make_res_layer` :247-326, `ResNet` :330-527): same constructor arguments,
sub-module names (`conv1/bn1/conv2/bn2/conv3/bn3/downsample.{0,1}`, `layer1..4`), state_dict keys,
`init_weights` and `train()` semantics.  The unused variants (BasicBlock, avd, deep_stem, avg_down,
caffe style, checkpointing) are rejected loudly instead of silently ignored.
"""
import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from .builder import BACKBONES
from .common import build_norm_layer, get_norm_type


def kaiming_init(module, mode='fan_out', nonlinearity='relu', bias=0):
    """mmcv.cnn.kaiming_init as called by ResNet.init_weights (resnet.py:471-473)."""
    nn.init.kaiming_normal_(module.weight, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    """mmcv.cnn.constant_init (resnet.py:474-475)."""
    nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def _conv1x1_id(conv, x):
    """conv1 of a Bottleneck in the fused configuration: (out, sums, identity) where `identity` aliases x and carries
    its gradient back into the convolution's own backward (ops._Conv1x1); None when the layer is not eligible."""
    from . import ops
    from .mvf import MVF
    if not (x.requires_grad and torch.is_grad_enabled() and ops.add_fusion_enabled()):
        return None
    if isinstance(conv, MVF):
        return conv.forward_with_identity(x)
    if (type(conv) is nn.Conv2d and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.bias is None
            and conv.groups == 1 and ops.eligible(x, conv.in_channels, conv.out_channels)):
        return ops.conv1x1(x, conv.weight, True, True)
    return None


def _conv1x1(conv, x, want_stats=False):
    """Run a 1x1 stride-1 nn.Conv2d (or the MVF wrapper around one) on the tcgen05 GEMM when the activation is
    bf16 channels_last (the training / inference configuration); anything else (strided down-sampling, fp32
    parity runs) is called as the module it is.  Returns (out, sums): `sums` is the (2, Cout) per-channel
    (sum, sum of squares) accumulated by the GEMM epilogue, or None."""
    from . import ops
    from .mvf import MVF
    if isinstance(conv, MVF):
        return conv(x, with_stats=True) if want_stats else (conv(x), None)
    if (type(conv) is nn.Conv2d and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.bias is None
            and conv.groups == 1 and ops.eligible(x, conv.in_channels, conv.out_channels)):
        if want_stats:
            return ops.conv1x1(x, conv.weight, True)
        return ops.conv1x1(x, conv.weight), None
    if (type(conv) is nn.Conv2d and conv.kernel_size == (1, 1) and conv.stride == (2, 2) and conv.bias is None
            and conv.groups == 1 and conv.padding == (0, 0) and ops.conv3x3_enabled()
            and ops.eligible(x, conv.in_channels, conv.out_channels) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0):
        if want_stats:
            return ops.conv1x1_strided(x, conv.weight, 2, True)
        return ops.conv1x1_strided(x, conv.weight, 2), None
    return conv(x), None


def _conv3x3(conv, x, want_stats=False):
    """Bottleneck.conv2 on the tcgen05 implicit GEMM when eligible (bf16 channels_last), else the module itself."""
    from . import ops
    if ops.conv3x3_eligible(x, conv):
        if want_stats:
            return ops.conv3x3(x, conv.weight, conv.stride[0], True)
        return ops.conv3x3(x, conv.weight, conv.stride[0]), None
    return conv(x), None


def _bn_act(bn, x, relu, residual=None, sums=None, relu_module=None):
    """BatchNorm2d (+ residual) (+ ReLU): one fused kernel pair when eligible, the torch modules otherwise."""
    from . import ops
    if ops.bn_eligible(x, bn) and (residual is None or (residual.dtype == x.dtype and residual.shape == x.shape)):
        if residual is not None:
            residual = residual.contiguous(memory_format=torch.channels_last)
        return ops.bn_act(x, bn, relu=relu, residual=residual, sums=sums)
    out = bn(x)
    if residual is not None:
        out = out + residual
    if relu:
        out = relu_module(out) if relu_module is not None else torch.relu(out)
    return out


class Bottleneck(nn.Module):
    """1x1 -> 3x3 (stride) -> 1x1 (x4) residual block, style='pytorch' (resnet.py:104-244)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 norm_cfg=dict(type='BN'), with_cp=False, avd=False, avd_first=False):
        super().__init__()
        if style != 'pytorch' or with_cp or avd or dilation != 1:
            raise NotImplementedError('mvfnet_b200.Bottleneck covers the MVFNet configs only: '
                                      "style='pytorch', dilation=1, no avd / checkpointing")
        self.inplanes, self.planes = inplanes, planes
        self.conv1_stride, self.conv2_stride = 1, stride
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=1, bias=False)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.add_module(self.norm1_name, norm1)
        self.add_module(self.norm2_name, norm2)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.norm3_name, norm3 = build_norm_layer(norm_cfg, planes * self.expansion, postfix=3)
        self.add_module(self.norm3_name, norm3)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride, self.dilation, self.norm_cfg, self.with_cp = stride, dilation, norm_cfg, with_cp

    norm1 = property(lambda self: getattr(self, self.norm1_name))
    norm2 = property(lambda self: getattr(self, self.norm2_name))
    norm3 = property(lambda self: getattr(self, self.norm3_name))

    def _infer(self, x):
        """Inference fast path (model.eval(), no autograd, bf16 channels_last): every BatchNorm, residual add and ReLU
        is the epilogue of the convolution before it -- three or four kernels per block (+ the MVF kernel), one HBM
        round trip per activation.  None when the block is not eligible."""
        from . import ops
        from .mvf import MVF
        bns = [self.norm1, self.norm2, self.norm3] + ([self.downsample[1]] if self.downsample is not None else [])
        c1 = self.conv1.net if isinstance(self.conv1, MVF) else self.conv1
        convs = [c1, self.conv2, self.conv3] + ([self.downsample[0]] if self.downsample is not None else [])
        if not ops.infer_eligible(x, *bns) or any(type(c) is not nn.Conv2d or c.bias is not None or c.groups != 1 or
                                                  c.in_channels % 64 or c.out_channels % 64 for c in convs):
            return None
        if x.shape[1] != c1.in_channels or (self.conv2.stride[0] == 2 and (x.shape[2] % 2 or x.shape[3] % 2)):
            return None
        if isinstance(self.conv1, MVF) and self.conv1.num_shift_channel:
            mvf = self.conv1
            if mvf.num_shift_channel % 64:
                return None
            from .mvf import mvf_slab_forward
            cfg, wt, wh, ww, gamma, beta, rm, rv = mvf._kernel_args()
            slab, xk, _, _, _ = mvf_slab_forward(x, cfg, wt, wh, ww, gamma, beta, rm, rv, out="slab")
            out = ops.conv1x1_bnact(xk, c1.weight, self.norm1, True, slab=slab, k0=cfg.Cs)
        else:
            out = ops.conv1x1_bnact(x, c1.weight, self.norm1, True)
        out = ops.conv_window_bnact(out, self.conv2.weight, self.norm2, self.conv2.stride[0], True)
        identity = x
        if self.downsample is not None:
            ds = self.downsample[0]
            if ds.stride[0] == 1:
                identity = ops.conv1x1_bnact(x, ds.weight, self.downsample[1], False)
            else:
                identity = ops.conv_window_bnact(x, ds.weight, self.downsample[1], ds.stride[0], False)
        return ops.conv1x1_bnact(out, self.conv3.weight, self.norm3, True, residual=identity)

    def forward(self, x):
        """resnet.py:208-244.  `self.conv1` is the MVF wrapper in the stages `mvf_freq` selects."""
        if not torch.is_grad_enabled() and not self.training:
            out = self._infer(x)
            if out is not None:
                return out
        fuse = x.is_cuda and x.dtype == torch.bfloat16
        identity = x
        fused = _conv1x1_id(self.conv1, x) if fuse else None
        if fused is not None:
            out, sums, identity = fused          # the identity / down-sample path reads the alias: one fused gradient
            x = identity
        else:
            out, sums = _conv1x1(self.conv1, x, fuse)
        out = _bn_act(self.norm1, out, True, sums=sums, relu_module=self.relu)
        out, sums = _conv3x3(self.conv2, out, fuse)
        out = _bn_act(self.norm2, out, True, sums=sums, relu_module=self.relu)
        out, sums = _conv1x1(self.conv3, out, fuse)
        if self.downsample is not None:
            ds, dsums = _conv1x1(self.downsample[0], x, fuse)
            identity = _bn_act(self.downsample[1], ds, False, sums=dsums)
        return _bn_act(self.norm3, out, True, residual=identity, sums=sums, relu_module=self.relu)


def make_res_layer(block, inplanes, planes, blocks, stride=1, dilation=1, style='pytorch', norm_cfg=None,
                   with_cp=False, avg_down=False, avd=False, avd_first=False):
    """One ResNet stage; the first block gets a 1x1-stride-s conv + norm on the identity path when the
    shape changes (resnet.py:279-326)."""
    if avg_down:
        raise NotImplementedError('avg_down is not used by the MVFNet configs')
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = nn.Sequential(
            nn.Conv2d(inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
            get_norm_type(norm_cfg)(planes * block.expansion))
    kw = dict(style=style, norm_cfg=norm_cfg, with_cp=with_cp, avd=avd, avd_first=avd_first)
    layers = [block(inplanes, planes, stride, dilation, downsample, **kw)]
    layers += [block(planes * block.expansion, planes, 1, dilation, **kw) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


@BACKBONES.register_module
class ResNet(nn.Module):
    """ResNet backbone (resnet.py:330-527); depth in {50, 101, 152} (bottleneck variants)."""

    arch_settings = {50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)), 152: (Bottleneck, (3, 8, 36, 3))}

    def __init__(self, depth, pretrained=None, in_channels=3, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style='pytorch', frozen_stages=-1,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, norm_frozen=False,
                 partial_norm=False, with_cp=False, avg_down=False, avd=False, avd_first=False,
                 deep_stem=False, stem_width=64):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError('invalid depth {} for resnet'.format(depth))
        if deep_stem:
            raise NotImplementedError('deep_stem is not used by the MVFNet configs')
        assert 1 <= num_stages <= 4 and len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages
        self.depth, self.in_channels, self.pretrained, self.num_stages = depth, in_channels, pretrained, num_stages
        self.strides, self.dilations, self.out_indices, self.style = strides, dilations, out_indices, style
        self.frozen_stages, self.norm_cfg, self.norm_eval = frozen_stages, norm_cfg, norm_eval
        self.norm_frozen, self.partial_norm, self.with_cp = norm_frozen, partial_norm, with_cp
        self.block, stage_blocks = self.arch_settings[depth]
        self.stage_blocks = stage_blocks[:num_stages]
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, 64, postfix=1)
        self.add_module(self.norm1_name, norm1)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.res_layers = []
        for i, num_blocks in enumerate(self.stage_blocks):
            planes = 64 * 2 ** i
            layer = make_res_layer(self.block, self.inplanes, planes, num_blocks, stride=strides[i],
                                   dilation=dilations[i], style=style, norm_cfg=norm_cfg, with_cp=with_cp,
                                   avg_down=avg_down, avd=avd, avd_first=avd_first)
            self.inplanes = planes * self.block.expansion
            name = 'layer{}'.format(i + 1)
            self.add_module(name, layer)
            self.res_layers.append(name)
        self.feat_dim = self.block.expansion * 64 * 2 ** (len(self.stage_blocks) - 1)

    norm1 = property(lambda self: getattr(self, self.norm1_name))

    def init_weights(self):
        """kaiming fan_out convs + unit BN when `pretrained is None`; else a non-strict checkpoint load
        (resnet.py:464-477).  Runs BEFORE the MVF splice, so checkpoints carry plain `conv1.weight`."""
        if isinstance(self.pretrained, str):
            from .checkpoint import load_checkpoint
            load_checkpoint(self, self.pretrained, map_location='cpu', strict=False)
        elif self.pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    kaiming_init(m)
                elif isinstance(m, nn.BatchNorm2d):
                    constant_init(m, 1)
        else:
            raise TypeError('pretrained must be a str or None')

    def _stem(self, x):
        """conv1 -> norm1 -> relu -> maxpool (resnet.py:481-484); own kernels in the bf16 configuration."""
        from . import ops
        if ops.stem_eligible(x, self.conv1) and ops.bn_infer_ok(self.norm1):
            out = ops.stem_bnact(x, self.conv1.weight, self.norm1)
            return ops.maxpool3x3s2(out) if ops.maxpool_eligible(out, self.maxpool) else self.maxpool(out)
        if ops.stem_eligible(x, self.conv1):
            out, sums = ops.stem_conv(x, self.conv1.weight, True)
        else:
            out, sums = self.conv1(x), None
        if ops.bn_relu_maxpool_eligible(out, self.norm1, self.maxpool):   # one pass, the activation is never stored
            return ops.bn_relu_maxpool(out, self.norm1, sums=sums)
        out = _bn_act(self.norm1, out, True, sums=sums, relu_module=self.relu)
        if ops.maxpool_eligible(out, self.maxpool):
            return ops.maxpool3x3s2(out)
        return self.maxpool(out)

    def forward(self, x):
        x = self._stem(x)
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        return outs[0] if len(outs) == 1 else tuple(outs)

    def train(self, mode=True):
        """norm_eval / partial_norm / frozen_stages handling of resnet.py:496-527."""
        super().train(mode)
        if self.norm_eval:
            for m in self.modules():
                if isinstance(m, _BatchNorm):
                    m.eval()
                    if self.norm_frozen:
                        for p in m.parameters():
                            p.requires_grad = False
        if self.partial_norm:
            for i in range(1, self.frozen_stages + 1):
                for m in getattr(self, 'layer{}'.format(i)).modules():
                    if isinstance(m, _BatchNorm):
                        m.eval()
                        m.weight.requires_grad = False
                        m.bias.requires_grad = False
        if mode and self.frozen_stages >= 0:
            for p in list(self.conv1.parameters()) + list(self.norm1.parameters()):
                p.requires_grad = False
            self.norm1.eval()
            for i in range(1, self.frozen_stages + 1):
                mod = getattr(self, 'layer{}'.format(i))
                mod.eval()
                for p in mod.parameters():
                    p.requires_grad = False
        return self
