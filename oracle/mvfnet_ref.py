"""CPU oracle (torch, fp32/fp64) for the whole MVFNet-R50/R101 Recognizer2D step.

TEST INFRASTRUCTURE ONLY -- never imported by the product package `mvfnet_b200`; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.

The reference's arithmetic on this path lives in stock PyTorch ops, so this port is a
functional restatement over a flat `state_dict` with the reference's exact key names, calling the same
ATen ops in the same order as the reference modules do -- including MVF's split / cat / transpose /
contiguous data movement (MVF.py:109-137) -- so that timing it on host cores reproduces the
reference's own CPU path (bench.py `cpu_baseline.kind = "port"`).  It is pinned against the real
reference by `oracle/make_golden.py` -> `tests/golden/model_r50_*.npz`.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

ARCH = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}  # backbones/resnet.py:357-363


def param_shapes(depth=50, n_segment=8, alpha=0.125, mvf_freq=(0, 0, 1, 1), mode="THW", share=False,
                 num_classes=400, buffers=True):
    """Ordered {key: shape} of the reference Recognizer2D state_dict.

    Key order follows module registration order in the reference (backbones/resnet.py:424-455,
    Bottleneck.__init__ :157-186, MVF.__init__ MVF.py:57-87, heads/tsn_clshead.py:66-69).
    """
    shapes = OrderedDict()

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".bias"] = (c,)
        if buffers:
            shapes[prefix + ".running_mean"] = (c,)
            shapes[prefix + ".running_var"] = (c,)
            shapes[prefix + ".num_batches_tracked"] = ()

    shapes["backbone.conv1.weight"] = (64, 3, 7, 7)
    bn("backbone.bn1", 64)
    inplanes = 64
    for si, nblocks in enumerate(ARCH[depth]):
        planes = 64 * 2 ** si
        for bi in range(nblocks):
            pre = "backbone.layer%d.%d" % (si + 1, bi)
            if mvf_freq[si]:
                cs = int(inplanes * alpha)
                shapes[pre + ".conv1.net.weight"] = (planes, inplanes, 1, 1)
                if cs:
                    shapes[pre + ".conv1.shift_conv.weight"] = (cs, 1, 3, 1, 1)
                    bn(pre + ".conv1.bn", cs)
                    if not share:
                        if mode in ("THW", "TH"):
                            shapes[pre + ".conv1.h_conv.weight"] = (cs, 1, 1, 3, 1)
                        if mode == "THW":
                            shapes[pre + ".conv1.w_conv.weight"] = (cs, 1, 1, 1, 3)
            else:
                shapes[pre + ".conv1.weight"] = (planes, inplanes, 1, 1)
            shapes[pre + ".conv2.weight"] = (planes, planes, 3, 3)
            bn(pre + ".bn1", planes)
            bn(pre + ".bn2", planes)
            shapes[pre + ".conv3.weight"] = (planes * 4, planes, 1, 1)
            bn(pre + ".bn3", planes * 4)
            if bi == 0:
                shapes[pre + ".downsample.0.weight"] = (planes * 4, inplanes, 1, 1)
                bn(pre + ".downsample.1", planes * 4)
            inplanes = planes * 4
    shapes["cls_head.new_fc.weight"] = (num_classes, 2048)
    shapes["cls_head.new_fc.bias"] = (num_classes,)
    return shapes


def synth_state_dict(seed=0, dtype=torch.float32, conditioned=False, **kw):
    """Deterministic non-trivial weights for every key of `param_shapes` (generator-seeded, so the
    reference model in make_golden.py and the models under test load identical values).

    `conditioned=True` scales the last BatchNorm of every residual branch (bn3.weight) by 0.2: with gamma ~ U(0.5, 1.5)
    on all 16 branches the random network is so ill-conditioned that float32 arithmetic alone moves its parameter
    gradients by 2 % (measured against float64) and bf16 weight rounding by 100 %; damped, float32 stays below 0.5 %
    and a bf16 comparison measures the implementation instead of chaos."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in param_shapes(**kw).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = (0.1 * torch.randn(shp, generator=g)).to(dtype)
        elif k.endswith("running_var"):
            sd[k] = (0.5 + torch.rand(shp, generator=g)).to(dtype)
        elif len(shp) == 1 and k.endswith(".weight"):              # BN gamma
            sd[k] = (0.5 + torch.rand(shp, generator=g)).to(dtype)
            if conditioned and k.endswith("bn3.weight"):
                sd[k] = sd[k] * 0.2
        elif len(shp) == 1:                                          # BN beta / fc bias
            sd[k] = (0.1 * torch.randn(shp, generator=g)).to(dtype)
        elif len(shp) == 5:                                          # MVF taps, MVF.py:95-97
            n = 3 * shp[0]
            sd[k] = (torch.randn(shp, generator=g) * (math.sqrt(2.0 / n) * 8)).to(dtype)
        elif len(shp) == 2:                                          # new_fc
            sd[k] = (0.01 * torch.randn(shp, generator=g)).to(dtype)
        else:                                                        # conv: kaiming fan_out
            fan_out = shp[0] * shp[2] * shp[3]
            sd[k] = (torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_out)).to(dtype)
    return sd


class RefModel:
    """Functional MVFNet Recognizer2D over a state_dict (params become autograd leaves)."""

    def __init__(self, sd, depth=50, n_segment=8, alpha=0.125, mvf_freq=(0, 0, 1, 1), mode="THW",
                 share=False, use_hs=True, dropout_ratio=0.5, eps=1e-5, momentum=0.1):
        self.depth, self.T, self.alpha, self.mvf_freq = depth, n_segment, alpha, mvf_freq
        self.mode, self.share, self.use_hs = mode, share, use_hs
        self.dropout_ratio, self.eps, self.momentum = dropout_ratio, eps, momentum
        self.p = OrderedDict()
        for k, v in sd.items():
            v = v.clone()
            if v.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
                v.requires_grad_(True)
            self.p[k] = v
        self.training = True

    def parameters(self):
        return [v for v in self.p.values() if v.requires_grad]

    def named_parameters(self):
        return [(k, v) for k, v in self.p.items() if v.requires_grad]

    # nn.BatchNorm2d/3d (common/norm.py:66; MVF.py:69)
    def _bn(self, x, pre):
        p = self.p
        if self.training:
            p[pre + ".num_batches_tracked"] += 1
        return F.batch_norm(x, p[pre + ".running_mean"], p[pre + ".running_var"], p[pre + ".weight"],
                            p[pre + ".bias"], self.training, self.momentum, self.eps)

    # modules/MVF.py:104-138
    def _mvf(self, x, pre):
        p = self.p
        nt, c, h, w = x.size()
        cs = int(c * self.alpha)
        if cs != 0:
            n = nt // self.T
            x = x.view(n, self.T, c, h, w).transpose(1, 2)
            x0, x1 = x.split([cs, c - cs], dim=1)
            wt = p[pre + ".shift_conv.weight"]
            conv_t = F.conv3d(x0, wt, None, 1, (1, 0, 0), 1, cs)
            if self.mode in ("THW", "TH"):
                if self.share:
                    tmp_h = F.conv3d(x0.transpose(2, 3), wt, None, 1, (1, 0, 0), 1, cs).transpose(2, 3)
                else:
                    tmp_h = F.conv3d(x0, p[pre + ".h_conv.weight"], None, 1, (0, 1, 0), 1, cs)
            if self.mode == "THW":
                if self.share:
                    tmp_w = F.conv3d(x0.permute(0, 1, 4, 2, 3), wt, None, 1, (1, 0, 0), 1, cs).permute(0, 1, 3, 4, 2)
                else:
                    tmp_w = F.conv3d(x0, p[pre + ".w_conv.weight"], None, 1, (0, 0, 1), 1, cs)
                x0 = conv_t + tmp_h + tmp_w
            elif self.mode == "TH":
                x0 = conv_t + tmp_h
            else:
                x0 = conv_t
            if self.use_hs:
                x0 = self._bn(x0, pre + ".bn")
                x0 = x0 * (F.relu6(x0 + 3) / 6)               # common/se_module.py:11,22
            x = torch.cat([x0, x1], dim=1)
            x = x.transpose(1, 2).contiguous().view(nt, c, h, w)
        return F.conv2d(x, p[pre + ".net.weight"])

    # backbones/resnet.py:208-244
    def _bottleneck(self, x, pre, stride, has_mvf, has_ds):
        p = self.p
        identity = x
        out = self._mvf(x, pre + ".conv1") if has_mvf else F.conv2d(x, p[pre + ".conv1.weight"])
        out = F.relu(self._bn(out, pre + ".bn1"))
        out = F.conv2d(out, p[pre + ".conv2.weight"], None, stride, 1)
        out = F.relu(self._bn(out, pre + ".bn2"))
        out = F.conv2d(out, p[pre + ".conv3.weight"])
        out = self._bn(out, pre + ".bn3")
        if has_ds:
            identity = F.conv2d(x, p[pre + ".downsample.0.weight"], None, stride)
            identity = self._bn(identity, pre + ".downsample.1")
        return F.relu(out + identity)

    # backbones/resnet.py:479-494
    def backbone(self, x):
        p = self.p
        x = F.conv2d(x, p["backbone.conv1.weight"], None, 2, 3)
        x = F.relu(self._bn(x, "backbone.bn1"))
        x = F.max_pool2d(x, 3, 2, 1)
        for si, nblocks in enumerate(ARCH[self.depth]):
            for bi in range(nblocks):
                stride = 2 if (bi == 0 and si > 0) else 1
                x = self._bottleneck(x, "backbone.layer%d.%d" % (si + 1, bi), stride,
                                     bool(self.mvf_freq[si]), bi == 0)
        return x

    # heads/tsn_clshead.py:71-98 + segmental_consensuses/simple_consensus.py:54-58
    def head(self, x, num_seg):
        p = self.p
        x = F.adaptive_avg_pool2d(x, 1)
        if self.dropout_ratio:
            x = F.dropout(x, self.dropout_ratio, self.training)
        x = x.view(x.size(0), -1)
        s = F.linear(x, p["cls_head.new_fc.weight"], p["cls_head.new_fc.bias"])
        return s.reshape((-1, num_seg) + s.shape[1:]).mean(dim=1)

    # recognizers/recognizer2d.py:132-149 ; heads/base.py:40-45
    def forward_train(self, img_group, label):
        b = img_group.shape[0]
        imgs = img_group.reshape((-1, 3) + img_group.shape[3:])
        feat = self.backbone(imgs)
        score = self.head(feat, imgs.shape[0] // b)
        return F.cross_entropy(score, label.reshape(-1)), score

    # recognizers/recognizer2d.py:151-179 (non-fcn) ; recognizers/base.py:47-74
    def forward_test(self, img_group, average_clips="prob"):
        imgs = img_group.reshape((-1, 3) + img_group.shape[3:])
        feat = self.backbone(imgs)
        score = self.head(feat, self.T)
        if average_clips == "prob":
            score = F.softmax(score, dim=1).mean(dim=0, keepdim=True)
        elif average_clips == "score":
            score = score.mean(dim=0, keepdim=True)
        return score
