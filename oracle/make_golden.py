"""Generate tests/golden/*.npz by running the UNMODIFIED reference (whwu95/MVFNet @ 0ddc7e2).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no /root/reference):

    python oracle/make_golden.py            # writes tests/golden/{mvf_cases,bottleneck_cases,model_r50,model_r50_224}.npz

The reference is imported from /root/reference with `oracle/mmcv_stub` standing in for mmcv 0.4.3
(import-time names and init helpers only; no arithmetic).  Inputs are generator-seeded torch tensors;
outputs/gradients come from the reference modules themselves:
  * codes/models/modules/MVF.py::MVF (wrapping nn.Identity so the output is x')      -> mvf_cases
  * codes/models/backbones/resnet.py::Bottleneck with MVF spliced into conv1          -> bottleneck_cases
  * codes/models/builder.py::build_recognizer(config model dict) full forward/backward -> model_r50
"""
import io
import os
import sys
import contextlib

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVFNET_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "mmcv_stub"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

with contextlib.redirect_stdout(io.StringIO()):
    from codes.models import build_recognizer                      # noqa: E402
    from codes.models.modules.MVF import MVF                        # noqa: E402
    from codes.models.backbones.resnet import Bottleneck           # noqa: E402

from oracle.mvfnet_ref import synth_state_dict                      # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (name, N, T, C, H, W, alpha, mode, share, use_hs, training)
MVF_CASES = [
    ("thw_train", 2, 4, 16, 5, 6, 0.25, "THW", False, True, True),
    ("thw_eval", 2, 4, 16, 5, 6, 0.25, "THW", False, True, False),
    ("thw_nohs", 1, 8, 16, 4, 4, 0.5, "THW", False, False, True),
    ("thw_share_train", 2, 3, 8, 4, 5, 0.5, "THW", True, True, True),
    ("t_train", 2, 4, 16, 3, 3, 0.25, "T", False, True, True),
    ("th_eval", 1, 4, 16, 4, 3, 0.25, "TH", False, True, False),
    ("th_share_eval", 2, 2, 8, 3, 4, 0.5, "TH", True, True, False),
    ("t1_train", 3, 1, 8, 4, 4, 0.5, "THW", False, True, True),      # T == 1: temporal taps see only zeros
    ("alpha0", 1, 4, 8, 3, 3, 0.0, "THW", False, True, True),         # num_shift_channel == 0 bypass
    ("r50_l4_shape", 1, 8, 2048 // 16, 7, 7, 0.125, "THW", False, True, True),
]


def gen(shape, g, scale=1.0, shift=0.0):
    return torch.randn(shape, generator=g, dtype=torch.float64) * scale + shift


def mvf_case(name, n, t, c, h, w, alpha, mode, share, use_hs, training, seed):
    g = torch.Generator().manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = MVF(nn.Identity(), t, c, alpha=alpha, use_hs=use_hs, share=share, mode=mode).double()
    cs = m.num_shift_channel
    rec = {"meta": np.array([n, t, c, h, w, cs, int(share), int(use_hs), int(training)]), "mode": np.array(mode)}
    x = gen((n * t, c, h, w), g).requires_grad_(True)
    rec["x"] = x.detach().numpy().copy()
    if cs:
        with torch.no_grad():
            for pn, p in m.named_parameters():
                if pn.endswith("conv.weight"):
                    p.copy_(gen(p.shape, g, 0.6))
                elif pn == "bn.weight":
                    p.copy_(gen(p.shape, g, 0.3, 1.0))
                elif pn == "bn.bias":
                    p.copy_(gen(p.shape, g, 1.5))          # spread over all three hardswish regimes
            m.bn.running_mean.copy_(gen((cs,), g, 0.5))
            m.bn.running_var.copy_(torch.rand((cs,), generator=g, dtype=torch.float64) + 0.5)
        for pn, p in m.named_parameters():
            rec["p." + pn] = p.detach().numpy().copy()
        rec["rm"] = m.bn.running_mean.numpy().copy()
        rec["rv"] = m.bn.running_var.numpy().copy()
    m.train(training)
    y = m(x)
    gy = gen(y.shape, g)
    rec["gy"] = gy.numpy().copy()
    y.backward(gy)
    rec["out"] = y.detach().numpy().copy()
    rec["dx"] = x.grad.numpy().copy()
    if cs:
        for pn, p in m.named_parameters():
            if p.grad is not None:                       # use_hs=False leaves bn.* unused
                rec["g." + pn] = p.grad.numpy().copy()
        rec["rm_after"] = m.bn.running_mean.numpy().copy()
        rec["rv_after"] = m.bn.running_var.numpy().copy()
    return {name + "/" + k: v for k, v in rec.items()}


# (name, F=N*T, T, inplanes, planes, H, W, stride, downsample, mvf(alpha or None), training)
BNECK_CASES = [
    ("mvf_s1_train", 4, 2, 32, 8, 6, 6, 1, False, 0.25, True),
    ("mvf_s2_ds_train", 4, 4, 16, 8, 6, 6, 2, True, 0.25, True),
    ("mvf_s2_ds_eval", 4, 2, 16, 8, 8, 8, 2, True, 0.25, False),
    ("plain_s1_ds_train", 3, 1, 8, 4, 5, 5, 1, True, None, True),
]


def bottleneck_case(name, f, t, inplanes, planes, h, w, stride, ds, alpha, training, seed):
    g = torch.Generator().manual_seed(seed)
    downsample = None
    if ds:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False),
                                   nn.BatchNorm2d(planes * 4))       # make_res_layer, resnet.py:299-304
    blk = Bottleneck(inplanes, planes, stride, 1, downsample)
    if alpha is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            blk.conv1 = MVF(blk.conv1, t, inplanes, alpha, True, False, "THW")   # MVF.py:38-39
    blk = blk.double()
    with torch.no_grad():
        for pn, p in blk.named_parameters():
            if p.dim() == 1 and pn.endswith("weight"):
                p.copy_(gen(p.shape, g, 0.3, 1.0))
            elif p.dim() == 1:
                p.copy_(gen(p.shape, g, 0.5))
            else:
                p.copy_(gen(p.shape, g, 0.4))
        for bn_, b in blk.named_buffers():
            if bn_.endswith("running_mean"):
                b.copy_(gen(b.shape, g, 0.3))
            elif bn_.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g, dtype=torch.float64) + 0.5)
    rec = {"meta": np.array([f, t, inplanes, planes, h, w, stride, int(ds), int(alpha is not None), int(training)]),
           "alpha": np.array(0.0 if alpha is None else alpha)}
    for k, v in blk.state_dict().items():
        rec["sd." + k] = v.numpy().copy()
    x = gen((f, inplanes, h, w), g).requires_grad_(True)
    rec["x"] = x.detach().numpy().copy()
    blk.train(training)
    y = blk(x)
    gy = gen(y.shape, g)
    rec["gy"] = gy.numpy().copy()
    y.backward(gy)
    rec["out"] = y.detach().numpy().copy()
    rec["dx"] = x.grad.numpy().copy()
    for pn, p in blk.named_parameters():
        rec["g." + pn] = p.grad.numpy().copy()
    for k, v in blk.state_dict().items():
        if "running" in k:
            rec["after." + k] = v.numpy().copy()
    return {name + "/" + k: v for k, v in rec.items()}


def model_cfg(depth, t, dropout):
    """configs/MVFNet/K400/mvf_kinetics400_2d_rgb_r50_dense.py:20-48 with pretrained=None."""
    return dict(
        type="Recognizer2D",
        backbone=dict(type="ResNet", pretrained=None, depth=depth, out_indices=(3,), norm_eval=False,
                      partial_norm=False, norm_cfg=dict(type="BN", requires_grad=True)),
        cls_head=dict(type="TSNClsHead", spatial_size=-1, spatial_type="avg", with_avg_pool=False,
                      temporal_feature_size=1, spatial_feature_size=1, dropout_ratio=dropout,
                      in_channels=2048, init_std=0.01, num_classes=400),
        module_cfg=dict(type="MVF", n_segment=t, alpha=0.125, mvf_freq=(0, 0, 1, 1), mode="THW"))


def model_case(depth=50, t=4, b=2, px=64, seed=0):
    """Whole Recognizer2D: eval probabilities, train loss, per-parameter gradient norms + a few
    full gradients, on synth_state_dict(seed) weights (fp32, as the reference runs)."""
    with contextlib.redirect_stdout(io.StringIO()):
        m = build_recognizer(model_cfg(depth, t, 0.0), None, dict(average_clips="prob"))
    sd = synth_state_dict(seed, depth=depth, n_segment=t)
    ref_sd = m.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "oracle key order differs from the reference"
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn((b, t, 3, px, px), generator=g)
    label = torch.randint(0, 400, (b, 1), generator=g)
    rec = {"meta": np.array([depth, t, b, px, seed]), "img": img.numpy(), "label": label.numpy(),
           "keys": np.array(list(sd.keys())),
           "shapes": np.array([",".join(map(str, v.shape)) for v in sd.values()])}
    m.eval()
    with torch.no_grad():
        rec["eval_prob"] = m(img, None, return_loss=False)
    m.train()
    loss = m(img, label)["loss_cls"]
    loss.backward()
    rec["train_loss"] = np.array(loss.item())
    names, norms = [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(p.grad.double().norm().item())
        if k in ("cls_head.new_fc.bias", "backbone.layer4.2.conv1.shift_conv.weight",
                 "backbone.layer3.0.conv1.h_conv.weight", "backbone.layer3.0.conv1.bn.weight",
                 "backbone.layer4.0.conv1.w_conv.weight", "backbone.bn1.weight"):
            rec["grad." + k] = p.grad.numpy().copy()
    rec["grad_names"] = np.array(names)
    rec["grad_norms"] = np.array(norms)
    rec["rm_after.layer4.2.conv1.bn"] = m.state_dict()["backbone.layer4.2.conv1.bn.running_mean"].numpy().copy()
    rec["n_params"] = np.array(sum(p.numel() for p in m.parameters()))
    return rec


FULL_GRADS_224 = ("backbone.conv1.weight", "backbone.layer1.0.conv1.weight", "backbone.layer2.0.conv2.weight",
                  "backbone.layer3.2.conv1.net.weight", "cls_head.new_fc.bias")


def model_case_224(depth=50, t=8, b=2, px=224, seed=3):
    """The bench configuration's geometry (T = 8, 224 px: 56/28/14/7 stage resolutions, MVF on 14x14 and 7x7 slabs) at
    B = 2 clips, train mode, in FLOAT64 (the reference model .double(): the truth that float32 / bf16 implementations
    are measured against) on synth_state_dict(conditioned=True) weights: loss, every parameter's gradient norm, the full
    gradient of every 1-D parameter (BatchNorm affine, bias), of every MVF tap tensor and of a few convolutions, and the
    same quantities from the reference in float32 (its own rounding noise: the floor of any float32 comparison).
    The input is NOT stored (9.6 MB): it is regenerated from the seed with the CPU generator, its checksum is."""
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn((b, t, 3, px, px), generator=g)
    label = torch.randint(0, 400, (b, 1), generator=g)
    rec = {"meta": np.array([depth, t, b, px, seed]), "label": label.numpy(),
           "img_checksum": np.array([img.double().sum().item(), img.double().abs().sum().item()])}

    def run(dtype):
        with contextlib.redirect_stdout(io.StringIO()):
            m = build_recognizer(model_cfg(depth, t, 0.0), None, dict(average_clips="prob"))
        m.load_state_dict(synth_state_dict(seed, depth=depth, n_segment=t, conditioned=True))
        m = m.to(dtype).train()
        loss = m(img.to(dtype), label)["loss_cls"]
        loss.backward()
        return loss.item(), {k: p.grad.double().numpy() for k, p in m.named_parameters()}

    loss64, g64 = run(torch.float64)
    loss32, g32 = run(torch.float32)
    rec["train_loss"], rec["train_loss_f32"] = np.array(loss64), np.array(loss32)
    names = list(g64)
    rec["grad_names"] = np.array(names)
    rec["grad_norms"] = np.array([np.linalg.norm(g64[k]) for k in names])
    rec["f32_rel_l2"] = np.array([np.linalg.norm(g32[k] - g64[k]) / max(np.linalg.norm(g64[k]), 1e-300) for k in names])
    for k in names:
        if g64[k].ndim == 1 or "shift_conv" in k or "h_conv" in k or "w_conv" in k or k in FULL_GRADS_224:
            rec["grad." + k] = g64[k].astype(np.float32)
    return rec


def tail_cases(seed=400):
    """The ends of the step, from the reference's own classes: Normalize + FormatShape (datasets/pipelines), TSNClsHead +
    BaseHead.loss (models/heads), and clip_grad_norm_ + torch.optim.SGD in DistOptimizerHook's order (core/dist_utils.py)."""
    with contextlib.redirect_stdout(io.StringIO()):
        from codes.datasets.pipelines.augmentations import Normalize
        from codes.datasets.pipelines.formating import FormatShape
        from codes.models.heads.tsn_clshead import TSNClsHead
    rng = np.random.RandomState(seed)
    rec = {}
    # -- input pipeline (config r50_dense.py:70-75)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    frames = rng.randint(0, 256, (8, 6, 8, 3)).astype(np.uint8)
    r = Normalize(mean=mean, std=std, div_255=False, to_rgb=True)(dict(img_group=[f.copy() for f in frames], modality="RGB"))
    r.update(num_clips=1, clip_len=4)
    rec["norm/frames"], rec["norm/mean"], rec["norm/std"] = frames, np.array(mean), np.array(std)
    rec["norm/out"] = FormatShape("NCHW")(r)["img_group"]
    # -- head + loss (fp64, dropout off: the reference's dropout mask is torch's RNG stream)
    g = torch.Generator().manual_seed(seed)
    nb, t, c, nc = 3, 4, 64, 10
    with contextlib.redirect_stdout(io.StringIO()):
        h = TSNClsHead(spatial_size=-1, spatial_type="avg", with_avg_pool=False, temporal_feature_size=1,
                       spatial_feature_size=1, dropout_ratio=0.0, in_channels=c, init_std=0.01, num_classes=nc).double()
    with torch.no_grad():
        h.new_fc.weight.copy_(gen(h.new_fc.weight.shape, g, 0.3))
        h.new_fc.bias.copy_(gen(h.new_fc.bias.shape, g, 0.5))
    x = gen((nb * t, c, 3, 3), g).requires_grad_(True)
    labels = torch.randint(0, nc, (nb, 1), generator=g)
    loss = h.loss(h(x, t), labels.squeeze())["loss_cls"]
    loss.backward()
    rec.update({"head/x": x.detach().numpy(), "head/w": h.new_fc.weight.detach().numpy(), "head/b": h.new_fc.bias.detach().numpy(),
                "head/labels": labels.numpy(), "head/T": np.array(t), "head/loss": np.array(loss.item()),
                "head/dx": x.grad.numpy(), "head/dw": h.new_fc.weight.grad.numpy(), "head/db": h.new_fc.bias.grad.numpy()})
    # -- optimizer tail: two steps of (sum over 2 ranks) / 2 -> clip 40 -> SGD nesterov (r50_dense.py:152-154)
    shapes = [(7, 5), (11,), (3, 2, 2, 2)]
    ps = [torch.nn.Parameter(gen(sh, g)) for sh in shapes]
    opt = torch.optim.SGD(ps, lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True)
    rec["sgd/p0"] = np.concatenate([p_.detach().numpy().ravel() for p_ in ps])
    for step in range(2):
        gs = [gen(sh, g, 30.0 if step == 0 else 0.5) for sh in shapes]          # step 0 clips, step 1 does not
        rec["sgd/g%d" % step] = np.concatenate([gi.numpy().ravel() for gi in gs])
        for p_, gi in zip(ps, gs):
            p_.grad = gi / 2.0                                                  # _allreduce_coalesced: / world_size
        total = torch.nn.utils.clip_grad_norm_(ps, max_norm=40, norm_type=2)
        opt.step()
        rec["sgd/norm%d" % step] = np.array(float(total))
        rec["sgd/p%d" % (step + 1)] = np.concatenate([p_.detach().numpy().ravel() for p_ in ps])
    rec["sgd/shapes"] = np.array([",".join(map(str, sh)) for sh in shapes])
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    recs = {}
    for i, c in enumerate(MVF_CASES):
        recs.update(mvf_case(*c, seed=100 + i))
    np.savez_compressed(os.path.join(OUT, "mvf_cases.npz"), **recs)
    recs = {}
    for i, c in enumerate(BNECK_CASES):
        recs.update(bottleneck_case(*c, seed=200 + i))
    np.savez_compressed(os.path.join(OUT, "bottleneck_cases.npz"), **recs)
    np.savez_compressed(os.path.join(OUT, "model_r50.npz"), **model_case(50, 4, 2, 64, 0))
    np.savez_compressed(os.path.join(OUT, "model_r50_224.npz"), **model_case_224())
    np.savez_compressed(os.path.join(OUT, "tail_cases.npz"), **tail_cases())
    # structural known-answers (config docstrings r50_dense.py:1-5 / r101_dense.py:1-5)
    counts = {}
    for depth in (50, 101):
        with contextlib.redirect_stdout(io.StringIO()):
            m = build_recognizer(model_cfg(depth, 8, 0.5), None, None)
        counts["params_r%d" % depth] = np.array(sum(p.numel() for p in m.parameters()))
        counts["keys_r%d" % depth] = np.array(list(m.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, "structure.npz"), **counts)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
