"""GPU parity of the Bottleneck (+MVF) block and the whole Recognizer2D against golden vectors produced by
the unmodified reference (oracle/make_golden.py), fp32 storage, tolerance 1e-3 relative."""
import numpy as np
import pytest
import torch

from conftest import load_cases, GOLDEN
from oracle.mvfnet_ref import synth_state_dict

pytestmark = pytest.mark.gpu

BNECK = load_cases("bottleneck_cases.npz")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize("name", sorted(BNECK))
@pytest.mark.parametrize("channels_last", [False, True])
def test_bottleneck_golden(name, channels_last):
    from mvfnet_b200 import Bottleneck, MVF
    import torch.nn as nn
    c = BNECK[name]
    f, t, inpl, planes, h, w, stride, ds, has_mvf, training = [int(v) for v in c["meta"]]
    downsample = None
    if ds:
        downsample = nn.Sequential(nn.Conv2d(inpl, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
    blk = Bottleneck(inpl, planes, stride, 1, downsample)
    if has_mvf:
        blk.conv1 = MVF(blk.conv1, t, inpl, float(c["alpha"]), True, False, "THW")
    sd = {k[3:]: torch.as_tensor(v) for k, v in c.items() if k.startswith("sd.")}
    blk.load_state_dict(sd)
    blk = blk.float().cuda().train(bool(training))
    x = torch.as_tensor(c["x"]).float().cuda()
    gy = torch.as_tensor(c["gy"]).float().cuda()
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = blk(x)
    y.backward(gy)
    assert rel_err(y.detach().cpu().numpy(), c["out"]) < 1e-3
    assert rel_err(x.grad.cpu().numpy(), c["dx"]) < 1e-3
    for k, p in blk.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), c["g." + k]) < 1e-3, k
    if training:
        for k, v in blk.state_dict().items():
            if "running" in k:
                assert rel_err(v.cpu().numpy(), c["after." + k]) < 1e-3, k


def model_cfg(depth, t, dropout):
    return dict(
        type="Recognizer2D",
        backbone=dict(type="ResNet", pretrained=None, depth=depth, out_indices=(3,), norm_eval=False,
                      partial_norm=False, norm_cfg=dict(type="BN", requires_grad=True)),
        cls_head=dict(type="TSNClsHead", spatial_size=-1, spatial_type="avg", with_avg_pool=False,
                      temporal_feature_size=1, spatial_feature_size=1, dropout_ratio=dropout,
                      in_channels=2048, init_std=0.01, num_classes=400),
        module_cfg=dict(type="MVF", n_segment=t, alpha=0.125, mvf_freq=(0, 0, 1, 1), mode="THW"))


def test_whole_model_golden_fp32():
    from mvfnet_b200 import build_recognizer
    z = np.load(GOLDEN + "/model_r50.npz")
    depth, t, b, px, seed = [int(v) for v in z["meta"]]
    m = build_recognizer(model_cfg(depth, t, 0.0), None, dict(average_clips="prob"))
    sd = synth_state_dict(seed, depth=depth, n_segment=t)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    img, label = torch.from_numpy(z["img"]).cuda(), torch.from_numpy(z["label"]).cuda()
    m.eval()
    with torch.no_grad():
        prob = m(img, None, return_loss=False)
    np.testing.assert_allclose(prob, z["eval_prob"], rtol=1e-3, atol=1e-6)
    m.train()
    loss = m(img, label)["loss_cls"]
    loss.backward()
    assert abs(loss.item() - float(z["train_loss"])) < 1e-3 * abs(float(z["train_loss"]))
    norms = dict(zip([str(k) for k in z["grad_names"]], z["grad_norms"]))
    for k, p in m.named_parameters():
        assert abs(p.grad.double().norm().item() - norms[k]) <= 5e-3 * norms[k] + 1e-7, k
    for k in z.files:
        if k.startswith("grad."):
            g = dict(m.named_parameters())[k[5:]].grad.cpu().numpy()
            # full gradients of early layers carry the rounding differences of ~50 train-mode BN layers evaluated
            # on 8 frames at 2x2..16x16 resolution (cuDNN vs the reference's oneDNN summation order)
            assert rel_err(g, z[k]) < 3e-2, k
    rm = m.state_dict()["backbone.layer4.2.conv1.bn.running_mean"].cpu().numpy()
    assert rel_err(rm, z["rm_after.layer4.2.conv1.bn"]) < 1e-3


def _golden_224():
    z = np.load(GOLDEN + "/model_r50_224.npz")
    depth, t, b, px, seed = [int(v) for v in z["meta"]]
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn((b, t, 3, px, px), generator=g)
    label = torch.randint(0, 400, (b, 1), generator=g)
    chk = np.array([img.double().sum().item(), img.double().abs().sum().item()])
    np.testing.assert_allclose(chk, z["img_checksum"], rtol=1e-12, err_msg="the seeded input is not the golden run's")
    assert np.array_equal(label.numpy(), z["label"])
    sd = synth_state_dict(seed, depth=depth, n_segment=t, conditioned=True)
    return z, depth, t, sd, img, label


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _record(name, rec):
    import json
    import os
    try:
        out = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(rec, f, indent=1, sort_keys=True)
    except OSError:
        pass


def test_whole_model_224_fp32_gradients_vs_reference():
    """R50 8x8 at 224 px (the bench geometry), B = 2, fp32 storage, against the FLOAT64 run of the unmodified reference
    (oracle/make_golden.py::model_case_224): loss within 1e-5; every stored gradient within 1e-2 relative L2 and every
    parameter's gradient norm within 1e-2.  The reference's own float32 run sits 2.8e-3 (mean) / 7.2e-3 (max) from its
    float64 run on this network (`f32_rel_l2` in the fixture): float32 cannot do better than that floor, whatever the
    implementation; measured values go to gpurun_out/fp32_model_grad_parity.json."""
    from mvfnet_b200 import build_recognizer
    z, depth, t, sd, img, label = _golden_224()
    m = build_recognizer(model_cfg(depth, t, 0.0), None, None)
    m.load_state_dict(sd)
    m = m.cuda().train()
    loss = m(img.cuda(), label.cuda())["loss_cls"]
    loss.backward()
    assert abs(loss.item() - float(z["train_loss"])) < 1e-5 * abs(float(z["train_loss"]))
    params = dict(m.named_parameters())
    norms = dict(zip([str(k) for k in z["grad_names"]], z["grad_norms"]))
    rec = {"norm_ratio": {k: p.grad.double().norm().item() / max(norms[k], 1e-30) for k, p in params.items()},
           "rel_l2": {k[5:]: _rel_l2(params[k[5:]].grad.cpu().numpy(), z[k]) for k in z.files if k.startswith("grad.")},
           "reference_f32_floor": {"mean": float(z["f32_rel_l2"].mean()), "max": float(z["f32_rel_l2"].max())}}
    _record("fp32_model_grad_parity.json", rec)
    bad = {k: v for k, v in rec["norm_ratio"].items() if abs(v - 1.0) > 1e-2}
    assert not bad, sorted(bad.items(), key=lambda kv: -abs(kv[1] - 1))[:5]
    worst = sorted(rec["rel_l2"].items(), key=lambda kv: -kv[1])[:5]
    assert worst[0][1] < 1e-2, worst


def _reference_autocast_grads(depth, t, sd, img, label):
    """Gradients of the UNMODIFIED reference model (baseline/_ref) under torch.autocast(bfloat16) on this GPU: the bf16
    semantics of the reference itself, and the yardstick for how far bf16 storage moves this network's gradients."""
    import contextlib
    import io
    import os
    import sys
    root = os.path.join(os.path.dirname(GOLDEN), "..", "baseline", "_ref")
    if not os.path.isdir(os.path.join(root, "MVFNet", "codes")):
        return None
    sys.path[:0] = [os.path.join(root, "mmcv_stub"), os.path.join(root, "MVFNet")]
    with contextlib.redirect_stdout(io.StringIO()):
        from codes.models import build_recognizer as ref_build
        ref = ref_build(model_cfg(depth, t, 0.0), None, None)
    ref.load_state_dict(sd)
    ref = ref.cuda().train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = ref(img.cuda(), label.cuda())["loss_cls"]
    loss.backward()
    return loss.item(), {k: p.grad.detach().float().cpu().numpy() for k, p in ref.named_parameters()}


def test_whole_model_224_bf16_fused_gradients_vs_reference():
    """The SAME fixture through the production configuration (bf16 autocast, channels_last: every convolution,
    BatchNorm, MVF module, stem, max-pool, head and loss on this library's kernels).  Whole-network gradients of a
    random-weight R50 are sensitive: rounding the weights alone to bf16 moves them by ~0.3 relative L2 (fp64 math,
    measured with the oracle port), so the meaningful statement is comparative -- per parameter,
        E_ours = relL2(our bf16 gradient, float64 reference)   vs   E_ref = relL2(reference under torch.autocast, float64 reference)
    on the same GPU: the median of E_ours / E_ref must be <= 1.15 and no parameter may exceed 2x (nor E_ours > 0.6),
    i.e. this library's bf16 path is as close to the float64 truth as the reference's own bf16 path is.  The loss must
    agree with the float64 loss within 1e-2.  Measured values go to gpurun_out/bf16_model_grad_parity.json."""
    from mvfnet_b200 import build_recognizer, _lib
    from mvfnet_b200.utils import to_channels_last
    z, depth, t, sd, img, label = _golden_224()
    m = build_recognizer(model_cfg(depth, t, 0.0), None, None)
    m.load_state_dict(sd)
    m = to_channels_last(m.cuda()).train()
    before = _lib.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m(img.cuda(), label.cuda())["loss_cls"]
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - before > 300, "the fused kernels were not used"
    params = dict(m.named_parameters())
    ref = _reference_autocast_grads(depth, t, sd, img, label)
    if ref is None:
        pytest.skip("baseline/_ref (the unmodified reference) is not installed next to the tests")
    ref_loss, ref_grads = ref
    keys = [k[5:] for k in z.files if k.startswith("grad.")]
    rec = {"loss": loss.item(), "loss_f64": float(z["train_loss"]), "loss_reference_autocast": ref_loss, "E_ours": {}, "E_ref": {}}
    for k in keys:
        rec["E_ours"][k] = _rel_l2(params[k].grad.float().cpu().numpy(), z["grad." + k])
        rec["E_ref"][k] = _rel_l2(ref_grads[k], z["grad." + k])
    ratios = {k: rec["E_ours"][k] / max(rec["E_ref"][k], 1e-12) for k in keys}
    rec["ratio_median"] = float(np.median(list(ratios.values())))
    rec["ratio_max"] = float(max(ratios.values()))
    _record("bf16_model_grad_parity.json", rec)
    assert abs(rec["loss"] - rec["loss_f64"]) < 1e-2 * abs(rec["loss_f64"])
    assert rec["ratio_median"] <= 1.15, rec["ratio_median"]
    worst = sorted(ratios.items(), key=lambda kv: -kv[1])[:5]
    assert worst[0][1] <= 2.0, worst
    assert max(rec["E_ours"].values()) < 0.6, sorted(rec["E_ours"].items(), key=lambda kv: -kv[1])[:5]


def test_whole_model_bf16_channels_last_runs():
    """The bench configuration (bf16 autocast, channels_last): loss close to the fp32 golden loss."""
    from mvfnet_b200 import build_recognizer
    z = np.load(GOLDEN + "/model_r50.npz")
    depth, t, b, px, seed = [int(v) for v in z["meta"]]
    m = build_recognizer(model_cfg(depth, t, 0.0), None, dict(average_clips="prob"))
    m.load_state_dict(synth_state_dict(seed, depth=depth, n_segment=t))
    from mvfnet_b200.utils import to_channels_last
    m = to_channels_last(m.cuda()).train()
    img, label = torch.from_numpy(z["img"]).cuda(), torch.from_numpy(z["label"]).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m(img, label)["loss_cls"]
    loss.backward()
    assert abs(loss.item() - float(z["train_loss"])) < 3e-2 * abs(float(z["train_loss"]))
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_whole_model_bf16_eval_and_fcn_testing():
    """Inference path (BASELINE configs[4] style): eval-mode BN through the fused kernels in bf16, plain and fcn_testing
    heads (tsn_clshead.py:99-117 folds to mean + linear), against the reference's fp32 eval probabilities."""
    from mvfnet_b200 import build_recognizer, _lib
    from mvfnet_b200.utils import to_channels_last
    z = np.load(GOLDEN + "/model_r50.npz")
    depth, t, b, px, seed = [int(v) for v in z["meta"]]
    m = build_recognizer(model_cfg(depth, t, 0.0), None, dict(average_clips="prob"))
    m.load_state_dict(synth_state_dict(seed, depth=depth, n_segment=t))
    m = to_channels_last(m.cuda()).eval()
    img = torch.from_numpy(z["img"]).cuda()
    before = _lib.launch_count()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        prob = m(img, None, return_loss=False)
        m.fcn_testing = True
        m.cls_head.fcn_testing = True
        prob_fcn = m(img, None, return_loss=False)
    assert _lib.launch_count() - before > 100, "the fused kernels were not used"
    ref = z["eval_prob"]
    assert prob.shape == ref.shape and abs(prob.sum() - 1.0) < 1e-3
    # The synthetic weights give logits of O(100): the fp32 softmax is nearly one-hot and a 1 % bf16 perturbation of the
    # logits moves probability mass between the top classes, so the distribution is compared through its top class and
    # the mass the reference's top-5 classes receive; the two bf16 heads (per-frame linear then mean vs fcn: mean then
    # 1x1x1 conv) are algebraically identical in eval mode but round differently (bf16 pooled features vs an fp32 mean);
    # on these peaked distributions that moves up to ~0.1 of probability mass.
    top5 = np.argsort(ref[0])[-5:]
    for got in (prob, prob_fcn):
        assert got.argmax() == ref.argmax()
        assert abs(got[0, top5].sum() - ref[0, top5].sum()) < 0.1
        assert np.abs(got - ref).sum() < 0.5
    assert np.abs(prob - prob_fcn).sum() < 0.15


def test_fcn_testing_256_vs_reference_on_gpu():
    """BASELINE configs[4] semantics at 256 x 256: the UNMODIFIED reference (baseline/_ref) with fcn_testing=True on model
    and head, eval-mode BatchNorm, average_clips='prob' (test_recognizer.py:72-77, recognizer2d.py:151-179,
    tsn_clshead.py:99-117 -- its Conv3d head is created with .cuda(), so this comparison can only run on the GPU box)
    against (a) our fp32 model, 1e-4, and (b) the production inference path: bf16, eval BatchNorm folded into the
    convolution epilogues, uint8 frames normalised on the GPU, replayed from a CUDA graph (mvfnet_b200/infer.py)."""
    import contextlib
    import io
    import os
    import sys
    from mvfnet_b200 import build_recognizer, _lib
    from mvfnet_b200.infer import GraphedInference
    from mvfnet_b200.utils import to_channels_last
    from mvfnet_b200.tail import IMG_NORM_MEAN, IMG_NORM_STD
    root = os.path.join(os.path.dirname(GOLDEN), "..", "baseline", "_ref")
    if not os.path.isdir(os.path.join(root, "MVFNet", "codes")):
        pytest.skip("baseline/_ref (the unmodified reference) is not installed next to the tests")
    sys.path[:0] = [os.path.join(root, "mmcv_stub"), os.path.join(root, "MVFNet")]
    t, clips, px = 8, 3, 256
    cfg = model_cfg(50, t, 0.5)
    cfg["fcn_testing"] = True
    cfg["cls_head"]["fcn_testing"] = True
    sd = synth_state_dict(11, depth=50, n_segment=t, conditioned=True)
    with contextlib.redirect_stdout(io.StringIO()):
        from codes.models import build_recognizer as ref_build
        ref = ref_build({k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}, None, dict(average_clips="prob"))
    ref.load_state_dict(sd)
    ref = ref.cuda().eval()
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (1, clips * t, px, px, 3), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(IMG_NORM_MEAN), torch.tensor(IMG_NORM_STD)
    img = ((u8.float().flip(-1) - mean) / std).permute(0, 1, 4, 2, 3).contiguous()     # the reference's float32 wire format
    with torch.no_grad():
        want = ref(img.cuda(), None, return_loss=False)                                   # (1, 400) numpy
    assert want.shape == (1, 400) and abs(want.sum() - 1) < 1e-4

    ours = build_recognizer({k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}, None, dict(average_clips="prob"))
    ours.load_state_dict(sd)
    ours = ours.cuda().eval()
    with torch.no_grad():
        got32 = ours(img.cuda(), None, return_loss=False)
    np.testing.assert_allclose(got32, want, rtol=1e-3, atol=1e-6)

    ours = to_channels_last(ours)
    before = _lib.launch_count()
    eng = GraphedInference(ours, u8.cuda(), uint8_input=True)
    assert _lib.launch_count() - before > 150, "the fused inference kernels were not used"
    for _ in range(2):
        got = eng(u8.cuda()).float().cpu().numpy()
    assert got.shape == (1, 400) and abs(got.sum() - 1) < 1e-3
    assert got.argmax() == want.argmax()
    assert np.abs(got - want).sum() < 0.1, np.abs(got - want).sum()                     # L1 distance of the distributions
    top5 = np.argsort(want[0])[-5:]
    assert abs(got[0, top5].sum() - want[0, top5].sum()) < 0.03
    eager = eng._forward(u8.cuda()).float().cpu().numpy()
    np.testing.assert_allclose(got, eager, rtol=0, atol=1e-6)                           # the graph replays what eager computes
