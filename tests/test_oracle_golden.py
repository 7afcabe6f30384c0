"""CPU: the oracle (oracle/mvf_oracle.py, oracle/mvfnet_ref.py) against golden vectors produced by
the unmodified reference (oracle/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import numpy as np
import pytest
import torch

from conftest import load_cases, GOLDEN
from oracle import mvf_oracle as O
from oracle.mvfnet_ref import RefModel, synth_state_dict, param_shapes

MVF = load_cases("mvf_cases.npz")
BNECK = load_cases("bottleneck_cases.npz")


def mvf_args(c):
    n, t, C, h, w, cs, share, use_hs, training = [int(v) for v in c["meta"]]
    mode = str(c["mode"])
    kw = dict(mode=mode, share=bool(share), use_hs=bool(use_hs), training=bool(training))
    tap = lambda k: c[k].reshape(cs, 3) if k in c else None
    w3 = dict(wt=tap("p.shift_conv.weight"), wh=tap("p.h_conv.weight"), ww=tap("p.w_conv.weight"))
    bn = dict(gamma=c.get("p.bn.weight"), beta=c.get("p.bn.bias"), running_mean=c.get("rm"), running_var=c.get("rv"))
    return t, cs, w3, bn, kw


@pytest.mark.parametrize("name", sorted(MVF))
def test_mvf_forward_backward_vs_reference(name):
    c = MVF[name]
    t, cs, w3, bn, kw = mvf_args(c)
    r = O.mvf_forward(c["x"], t, cs, **w3, **bn, **kw)
    np.testing.assert_allclose(r["out"], c["out"], rtol=1e-11, atol=1e-12)
    if cs:
        # pass-through channels are bit-exact (MVF.py:110,135)
        assert np.array_equal(r["out"][:, cs:], c["x"][:, cs:])
        if kw["training"] and kw["use_hs"]:
            np.testing.assert_allclose(r["new_running_mean"], c["rm_after"], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(r["new_running_var"], c["rv_after"], rtol=1e-11, atol=1e-13)
    bn_b = {k: v for k, v in bn.items()}
    kwb = {k: v for k, v in kw.items()}
    g = O.mvf_backward(c["gy"], c["x"], t, cs, **w3, **bn_b, **kwb)
    np.testing.assert_allclose(g["dx"], c["dx"], rtol=1e-9, atol=1e-11)
    if cs:
        np.testing.assert_allclose(g["dwt"].ravel(), c["g.shift_conv.weight"].ravel(), rtol=1e-9, atol=1e-10)
        if "g.h_conv.weight" in c:
            np.testing.assert_allclose(g["dwh"].ravel(), c["g.h_conv.weight"].ravel(), rtol=1e-9, atol=1e-10)
        if "g.w_conv.weight" in c:
            np.testing.assert_allclose(g["dww"].ravel(), c["g.w_conv.weight"].ravel(), rtol=1e-9, atol=1e-10)
        if "g.bn.weight" in c:
            np.testing.assert_allclose(g["dgamma"], c["g.bn.weight"], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(g["dbeta"], c["g.bn.bias"], rtol=1e-9, atol=1e-10)


def test_clip_boundary_no_temporal_leak():
    """MVF.py:109: the T stencil never crosses from clip n to clip n+1 in the N*T batch axis."""
    rng = np.random.default_rng(0)
    n, t, c, h, w, cs = 3, 4, 8, 3, 3, 4
    x = rng.standard_normal((n * t, c, h, w))
    wt, wh, ww = (rng.standard_normal((cs, 3)) for _ in range(3))
    base = O.mvf_forward(x, t, cs, wt, wh, ww, use_hs=False)["out"]
    x2 = x.copy()
    x2[t:2 * t] += 10.0                                     # perturb clip 1 only
    pert = O.mvf_forward(x2, t, cs, wt, wh, ww, use_hs=False)["out"]
    assert np.array_equal(base[:t], pert[:t]) and np.array_equal(base[2 * t:], pert[2 * t:])


def bneck_params(c):
    sd = {k[3:]: v for k, v in c.items() if k.startswith("sd.")}
    f, t, inpl, planes, h, w, stride, ds, has_mvf, training = [int(v) for v in c["meta"]]
    c1 = "conv1.net.weight" if has_mvf else "conv1.weight"
    p = dict(conv1=sd[c1], conv2=sd["conv2.weight"], conv3=sd["conv3.weight"])
    for i in (1, 2, 3):
        p.update({"bn%d_gamma" % i: sd["bn%d.weight" % i], "bn%d_beta" % i: sd["bn%d.bias" % i],
                  "bn%d_rm" % i: sd["bn%d.running_mean" % i], "bn%d_rv" % i: sd["bn%d.running_var" % i]})
    if ds:
        p.update(ds_w=sd["downsample.0.weight"], ds_gamma=sd["downsample.1.weight"], ds_beta=sd["downsample.1.bias"],
                 ds_rm=sd["downsample.1.running_mean"], ds_rv=sd["downsample.1.running_var"])
    mvf = None
    if has_mvf:
        cs = int(inpl * float(c["alpha"]))
        mvf = dict(num_shift=cs, wt=sd["conv1.shift_conv.weight"].reshape(cs, 3), wh=sd["conv1.h_conv.weight"].reshape(cs, 3),
                   ww=sd["conv1.w_conv.weight"].reshape(cs, 3), gamma=sd["conv1.bn.weight"], beta=sd["conv1.bn.bias"],
                   rm=sd["conv1.bn.running_mean"], rv=sd["conv1.bn.running_var"])
    return p, mvf, t, stride, bool(training), bool(has_mvf), bool(ds)


@pytest.mark.parametrize("name", sorted(BNECK))
def test_bottleneck_vs_reference(name):
    c = BNECK[name]
    p, mvf, t, stride, training, has_mvf, ds = bneck_params(c)
    out, cache = O.bottleneck_forward(c["x"], p, stride, training, t, mvf)
    np.testing.assert_allclose(out, c["out"], rtol=1e-9, atol=1e-10)
    dx, g = O.bottleneck_backward(c["gy"], c["x"], p, stride, training, cache, t, mvf)
    np.testing.assert_allclose(dx, c["dx"], rtol=1e-8, atol=1e-9)
    names = {"conv2": "conv2.weight", "conv3": "conv3.weight", "conv1": "conv1.net.weight" if has_mvf else "conv1.weight"}
    for i in (1, 2, 3):
        names["bn%d_gamma" % i] = "bn%d.weight" % i
        names["bn%d_beta" % i] = "bn%d.bias" % i
    if ds:
        names.update(ds_w="downsample.0.weight", ds_gamma="downsample.1.weight", ds_beta="downsample.1.bias")
    if has_mvf:
        names.update(mvf_dwt="conv1.shift_conv.weight", mvf_dwh="conv1.h_conv.weight", mvf_dww="conv1.w_conv.weight",
                     mvf_dgamma="conv1.bn.weight", mvf_dbeta="conv1.bn.bias")
    for k, ref_k in names.items():
        np.testing.assert_allclose(g[k].ravel(), c["g." + ref_k].ravel(), rtol=1e-8, atol=1e-8, err_msg=k)
    if training:
        np.testing.assert_allclose(cache["s3"]["new_running_var"], c["after.bn3.running_var"], rtol=1e-10)


def test_structure_known_answers():
    """Config docstrings (r50_dense.py:1-5, r101_dense.py:1-5): 24.34 M / 43.36 M params; key contract."""
    z = np.load(GOLDEN + "/structure.npz")
    assert int(z["params_r50"]) == 24342416 and int(z["params_r101"]) == 43358480
    for depth in (50, 101):
        shapes = param_shapes(depth=depth)
        assert list(shapes.keys()) == [str(k) for k in z["keys_r%d" % depth]]
        n = sum(int(np.prod(s)) for k, s in shapes.items() if "running" not in k and "num_batches" not in k)
        assert n == int(z["params_r%d" % depth])


def test_whole_model_port_vs_reference():
    """oracle/mvfnet_ref.RefModel (the CPU-baseline port) reproduces the reference Recognizer2D."""
    z = np.load(GOLDEN + "/model_r50.npz")
    depth, t, b, px, seed = [int(v) for v in z["meta"]]
    sd = synth_state_dict(seed, depth=depth, n_segment=t)
    assert [str(k) for k in z["keys"]] == list(sd.keys())
    m = RefModel(sd, depth=depth, n_segment=t, dropout_ratio=0.0)
    img, label = torch.from_numpy(z["img"]), torch.from_numpy(z["label"])
    m.training = False
    with torch.no_grad():
        prob = m.forward_test(img)
    np.testing.assert_allclose(prob.numpy(), z["eval_prob"], rtol=2e-4, atol=1e-7)
    m.training = True
    loss, _ = m.forward_train(img, label)
    loss.backward()
    assert abs(loss.item() - float(z["train_loss"])) < 2e-5
    norms = dict(zip([str(k) for k in z["grad_names"]], z["grad_norms"]))
    for k, p in m.named_parameters():
        assert abs(p.grad.double().norm().item() - norms[k]) <= 2e-3 * norms[k] + 1e-7, k
    for k in z.files:
        if k.startswith("grad."):
            np.testing.assert_allclose(m.p[k[5:]].grad.numpy(), z[k], rtol=5e-3, atol=1e-5 * np.abs(z[k]).max())
    np.testing.assert_allclose(m.p["backbone.layer4.2.conv1.bn.running_mean"].numpy(), z["rm_after.layer4.2.conv1.bn"], rtol=1e-3, atol=1e-6)


TAIL = load_cases("tail_cases.npz")


def test_tail_normalize_format_vs_reference_pipeline():
    """Normalize + FormatShape of the reference's data pipeline (augmentations.py:343-396, formating.py:134-185)."""
    c = TAIL["norm"]
    out = O.normalize_format(c["frames"], c["mean"], c["std"], to_rgb=True)
    assert out.shape == c["out"].shape and out.dtype == np.float32
    assert np.array_equal(out, c["out"])                                 # bit-exact restatement of cv2's in-place arithmetic


def test_tail_head_loss_vs_reference_head():
    """TSNClsHead.forward + BaseHead.loss and autograd's gradients (tsn_clshead.py:71-98, heads/base.py:40-45)."""
    c = TAIL["head"]
    r = O.head_loss(c["x"], c["w"], c["b"], c["labels"], int(c["T"]))
    assert abs(r["loss"] - float(c["loss"])) < 1e-12
    np.testing.assert_allclose(r["dx"], c["dx"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(r["dw"], c["dw"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(r["db"], c["db"], rtol=1e-10, atol=1e-14)


def test_tail_sgd_step_vs_torch_in_hook_order():
    """/ world -> clip_grad_norm_(40) -> SGD(momentum 0.9, wd 1e-4, nesterov) (dist_utils.py:59-67, r50_dense.py:152-154)."""
    c = TAIL["sgd"]
    p, m = [c["p0"].copy()], [np.zeros_like(c["p0"])]
    for step in range(2):
        p, m, total = O.sgd_step(p, m, [c["g%d" % step]], 0.015, 0.9, 1e-4, True, max_norm=40, world=2)
        assert abs(total - float(c["norm%d" % step])) < 1e-9 * total
        np.testing.assert_allclose(p[0], c["p%d" % (step + 1)], rtol=1e-12, atol=1e-14)
    assert float(c["norm0"]) > 40 > float(c["norm1"])        # the first step clips, the second does not
