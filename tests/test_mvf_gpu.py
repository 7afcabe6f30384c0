"""GPU parity of the MVF CUDA path (through the C ABI of libmvf_b200.so) against
 (a) the golden vectors produced by the unmodified reference (tests/golden/mvf_cases.npz),
 (b) the numpy oracle (oracle/mvf_oracle.py) on seeded inputs at the R50/R101 slab shapes,
 (c) size-independent properties at BASELINE.json's full sizes.
Tolerances (BASELINE.json north_star): 1e-3 relative for fp32 storage, 1e-2 for bf16 storage."""
import numpy as np
import pytest
import torch

from conftest import load_cases
from oracle import mvf_oracle as O

pytestmark = pytest.mark.gpu

MVF_GOLD = load_cases("mvf_cases.npz")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def rel_err_kinks(a, b, max_outliers=8):
    """rel_err for dL/dx through hard-swish: its derivative jumps by |u|/6 at u = +-3, so a handful of elements
    whose u lies within rounding distance of a kink legitimately take the other branch in fp32 / bf16 than in
    the fp64 oracle.  Ignore at most `max_outliers` such elements (out of >= 1e5)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    err = np.abs(a - b)
    k = min(max_outliers, max(err.size // 20000, 0))
    if k:
        err = np.partition(err, err.size - k - 1)[:err.size - k]
    return float(err.max() / max(np.abs(b).max(), 1e-12))


def bf16_taps(w):
    """bf16 storage semantics: kernels that take bf16 activations use bf16 stencil taps (each of the nine taps rounded
    to bf16, exactly what the reference's autocast path does to its Conv3d weights -- include/mvf_b200.h).  The
    oracle is evaluated on the same rounded taps so that only the kernels' own arithmetic remains in the comparison."""
    if w is None:
        return None
    return torch.as_tensor(np.asarray(w)).to(torch.bfloat16).double().numpy()


def build_mvf(cs_in, t, alpha, mode, share, use_hs, params, rm, rv, training, dev="cuda"):
    from mvfnet_b200 import MVF
    m = MVF(torch.nn.Identity(), t, cs_in, alpha=alpha, use_hs=use_hs, share=share, mode=mode)
    if m.num_shift_channel:
        with torch.no_grad():
            for k, v in params.items():
                dict(m.named_parameters())[k].copy_(torch.as_tensor(v))
            m.bn.running_mean.copy_(torch.as_tensor(rm))
            m.bn.running_var.copy_(torch.as_tensor(rv))
    m = m.to(dev).float()
    m.train(training)
    return m


def run_case(c, dtype, channels_last):
    n, t, C, h, w, cs, share, use_hs, training = [int(v) for v in c["meta"]]
    mode = str(c["mode"])
    alpha = (cs + 0.5) / C
    params = {k[2:]: v for k, v in c.items() if k.startswith("p.")}
    m = build_mvf(C, t, alpha, mode, bool(share), bool(use_hs), params, c.get("rm"), c.get("rv"), bool(training))
    assert m.num_shift_channel == cs
    x = torch.as_tensor(c["x"]).to("cuda", dtype)
    gy = torch.as_tensor(c["gy"]).to("cuda", dtype)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
        gy = gy.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = m(x)
    y.backward(gy)
    return m, x, y


@pytest.mark.parametrize("name", sorted(MVF_GOLD))
@pytest.mark.parametrize("variant", ["f32_nchw", "f32_nhwc", "bf16_nhwc", "bf16_nchw"])
def test_mvf_golden(name, variant):
    c = MVF_GOLD[name]
    dtype = torch.float32 if variant.startswith("f32") else torch.bfloat16
    tol = 1e-3 if dtype == torch.float32 else 1e-2
    m, x, y = run_case(c, dtype, variant.endswith("nhwc"))
    cs = int(c["meta"][5])
    xin = torch.as_tensor(c["x"]).to(dtype).double().numpy()
    # the golden output was computed from the fp64 input; with bf16 storage compare against the oracle
    # evaluated on the ROUNDED input so that only the kernel's own arithmetic + output rounding remains
    if dtype == torch.float32:
        ref_out, ref_dx = c["out"], c["dx"]
        ref = None
    else:
        t = int(c["meta"][1])
        kw = dict(mode=str(c["mode"]), share=bool(c["meta"][6]), use_hs=bool(c["meta"][7]), training=bool(c["meta"][8]))
        tap = lambda k: bf16_taps(c[k].reshape(cs, 3)) if k in c else None
        w3 = dict(wt=tap("p.shift_conv.weight"), wh=tap("p.h_conv.weight"), ww=tap("p.w_conv.weight"))
        bn = dict(gamma=c.get("p.bn.weight"), beta=c.get("p.bn.bias"), running_mean=c.get("rm"), running_var=c.get("rv"))
        gy = torch.as_tensor(c["gy"]).to(dtype).double().numpy()
        ref = O.mvf_backward(gy, xin, t, cs, **w3, **bn, **kw)
        ref_out = O.mvf_forward(xin, t, cs, **w3, **bn, **kw)["out"]
        ref_dx = ref["dx"]
    out = y.detach().double().cpu().numpy()
    assert rel_err(out, ref_out) < tol, "forward"
    if cs:
        # pass-through channels bit-exact (MVF.py:110,135)
        assert torch.equal(y.detach()[:, cs:], x.detach()[:, cs:])
    assert rel_err(x.grad.double().cpu().numpy(), ref_dx) < tol * 2, "dx"
    if cs == 0:
        return
    gtol = tol * (2 if dtype == torch.float32 else 4)
    grads = {k: p.grad.double().cpu().numpy().ravel() for k, p in m.named_parameters() if p.grad is not None}
    if ref is None:
        for k, g in grads.items():
            assert rel_err(g, c["g." + k].ravel()) < gtol, k
        assert set("g." + k for k in grads) == {k for k in c if k.startswith("g.")}
    else:
        names = {"shift_conv.weight": "dwt", "h_conv.weight": "dwh", "w_conv.weight": "dww", "bn.weight": "dgamma",
                 "bn.bias": "dbeta"}
        for k, g in grads.items():
            assert rel_err(g, ref[names[k]].ravel()) < gtol, k
    if bool(c["meta"][8]) and bool(c["meta"][7]):
        assert rel_err(m.bn.running_mean.cpu().numpy(), c["rm_after"]) < tol * 2
        assert rel_err(m.bn.running_var.cpu().numpy(), c["rv_after"]) < tol * 2
        assert int(m.bn.num_batches_tracked) == 1


# (C, H, Cs) of every MVF instance in R50/R101 at 224 and 256 px + ragged shapes
SHAPES = [(512, 28, 64), (1024, 14, 128), (2048, 7, 256), (512, 32, 64), (1024, 16, 128), (2048, 8, 256),
          (24, 5, 3), (40, 9, 5)]


@pytest.mark.parametrize("C,H,Cs", SHAPES)
@pytest.mark.parametrize("T", [4, 8, 16])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_mvf_vs_oracle_model_shapes(C, H, Cs, T, training, dtype):
    from mvfnet_b200 import MVF
    N = 2
    W = H if H > 5 else H + 2
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    tol = 1e-3 if dtype == "f32" else 1e-2
    g = torch.Generator().manual_seed(1234 + C + H + T)
    m = MVF(torch.nn.Identity(), T, C, alpha=(Cs + 0.5) / C, use_hs=True, share=False, mode="THW")
    assert m.num_shift_channel == Cs
    with torch.no_grad():
        for p in (m.shift_conv.weight, m.h_conv.weight, m.w_conv.weight):
            p.copy_(torch.randn(p.shape, generator=g) * 0.6)
        m.bn.weight.copy_(1 + 0.1 * torch.randn(Cs, generator=g))
        m.bn.bias.copy_(0.5 * torch.randn(Cs, generator=g))
        m.bn.running_mean.copy_(torch.randn(Cs, generator=g))
        m.bn.running_var.copy_(0.5 + 1.5 * torch.rand(Cs, generator=g))
    rm0, rv0 = m.bn.running_mean.double().numpy().copy(), m.bn.running_var.double().numpy().copy()
    m = m.cuda().train(training)
    x = torch.randn((N * T, C, H, W), generator=g).to(tdt)
    gy = torch.randn((N * T, C, H, W), generator=g).to(tdt)
    xd = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = m(xd)
    y.backward(gy.cuda().contiguous(memory_format=torch.channels_last))
    exact = lambda p: p.detach().double().cpu().numpy().reshape(Cs, 3)
    tap = (lambda p: bf16_taps(exact(p))) if dtype == "bf16" else exact
    kw = dict(wt=tap(m.shift_conv.weight), wh=tap(m.h_conv.weight), ww=tap(m.w_conv.weight),
              gamma=m.bn.weight.detach().double().cpu().numpy(), beta=m.bn.bias.detach().double().cpu().numpy(),
              running_mean=rm0, running_var=rv0, mode="THW", use_hs=True, training=training)
    xs, gs = x.double().numpy(), gy.double().numpy()
    rf = O.mvf_forward(xs, T, Cs, **kw)
    kw.pop("training")
    rb = O.mvf_backward(gs, xs, T, Cs, training=training, **kw)
    assert rel_err(y.detach().double().cpu().numpy(), rf["out"]) < tol
    if dtype == "bf16":        # and against the reference semantics proper (fp32 taps): still within the bf16 tolerance
        kx = dict(kw, wt=exact(m.shift_conv.weight), wh=exact(m.h_conv.weight), ww=exact(m.w_conv.weight))
        assert rel_err(y.detach().double().cpu().numpy(), O.mvf_forward(xs, T, Cs, training=training, **kx)["out"]) < tol
    assert torch.equal(y.detach()[:, Cs:], xd.detach()[:, Cs:])
    assert rel_err_kinks(xd.grad.double().cpu().numpy(), rb["dx"]) < 2 * tol
    gt = 2 * tol if dtype == "f32" else 4 * tol
    assert rel_err(m.shift_conv.weight.grad.double().cpu().numpy().reshape(Cs, 3), rb["dwt"]) < gt
    assert rel_err(m.h_conv.weight.grad.double().cpu().numpy().reshape(Cs, 3), rb["dwh"]) < gt
    assert rel_err(m.w_conv.weight.grad.double().cpu().numpy().reshape(Cs, 3), rb["dww"]) < gt
    assert rel_err(m.bn.weight.grad.double().cpu().numpy(), rb["dgamma"]) < gt
    assert rel_err(m.bn.bias.grad.double().cpu().numpy(), rb["dbeta"]) < gt
    if training:
        assert rel_err(m.bn.running_mean.double().cpu().numpy(), rf["new_running_mean"]) < 2 * tol
        assert rel_err(m.bn.running_var.double().cpu().numpy(), rf["new_running_var"]) < 2 * tol


@pytest.mark.parametrize("mode,share,use_hs", [("T", False, True), ("TH", False, True), ("THW", True, True),
                                                ("TH", True, False), ("THW", False, False)])
@pytest.mark.parametrize("dtype,training", [("f32", True), ("bf16", True), ("bf16", False)])
def test_mvf_variants_vs_oracle(mode, share, use_hs, dtype, training):
    from mvfnet_b200 import MVF
    N, T, C, H, W, Cs = 3, 8, 64, 14, 14, 16
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    tol = 1e-3 if dtype == "f32" else 1e-2
    g = torch.Generator().manual_seed(7)
    m = MVF(torch.nn.Identity(), T, C, alpha=0.25, use_hs=use_hs, share=share, mode=mode)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.5 + (1.0 if p.dim() == 1 else 0.0))
    rm0 = torch.randn(Cs, generator=g)
    rv0 = 0.5 + torch.rand(Cs, generator=g)
    with torch.no_grad():
        m.bn.running_mean.copy_(rm0)
        m.bn.running_var.copy_(rv0)
    m = m.cuda().train(training)
    x = torch.randn((N * T, C, H, W), generator=g).to(tdt)
    gy = torch.randn((N * T, C, H, W), generator=g).to(tdt)
    xd = x.cuda()
    gd = gy.cuda()
    if dtype == "bf16":
        xd = xd.contiguous(memory_format=torch.channels_last)
        gd = gd.contiguous(memory_format=torch.channels_last)
    xd.requires_grad_(True)
    y = m(xd)
    y.backward(gd)
    rnd = bf16_taps if dtype == "bf16" else (lambda w: w)
    tap = lambda name: rnd(getattr(m, name).weight.detach().double().cpu().numpy().reshape(Cs, 3)) if hasattr(m, name) else None
    kw = dict(wt=tap("shift_conv"), wh=tap("h_conv"), ww=tap("w_conv"), gamma=m.bn.weight.detach().double().cpu().numpy(),
              beta=m.bn.bias.detach().double().cpu().numpy(), running_mean=rm0.double().numpy(), running_var=rv0.double().numpy(),
              mode=mode, share=share, use_hs=use_hs)
    rf = O.mvf_forward(x.double().numpy(), T, Cs, training=training, **kw)
    rb = O.mvf_backward(gy.double().numpy(), x.double().numpy(), T, Cs, training=training, **kw)
    assert rel_err(y.detach().double().cpu().numpy(), rf["out"]) < tol
    assert rel_err_kinks(xd.grad.double().cpu().numpy(), rb["dx"]) < 2 * tol
    gt = 2 * tol if dtype == "f32" else 4 * tol
    assert rel_err(m.shift_conv.weight.grad.double().cpu().numpy().reshape(Cs, 3), rb["dwt"]) < gt
    if rb["dwh"] is not None:
        assert rel_err(m.h_conv.weight.grad.double().cpu().numpy().reshape(Cs, 3), rb["dwh"]) < gt
    if rb["dww"] is not None:
        assert rel_err(m.w_conv.weight.grad.double().cpu().numpy().reshape(Cs, 3), rb["dww"]) < gt
    if use_hs:
        assert rel_err(m.bn.weight.grad.double().cpu().numpy(), rb["dgamma"]) < gt
        assert rel_err(m.bn.bias.grad.double().cpu().numpy(), rb["dbeta"]) < gt
    else:
        assert m.bn.weight.grad is None


def test_mvf_full_size_properties():
    """BASELINE config 2 size (B=12 clips of T=8 at the layer3 shape): properties that need no oracle."""
    from mvfnet_b200 import MVF
    B, T, C, H, Cs = 12, 8, 1024, 14, 128
    torch.manual_seed(0)
    m = MVF(torch.nn.Identity(), T, C, alpha=0.125, use_hs=False).cuda()
    x = torch.randn(B * T, C, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = m(x)
        # 1. untouched channels bit-exact
        assert torch.equal(y[:, Cs:], x[:, Cs:])
        # 2. clips are independent: perturbing clip 3 changes only clip 3
        x2 = x.clone()
        x2[3 * T:4 * T] += 1.0
        y2 = m(x2)
        same = torch.ones(B * T, dtype=torch.bool)
        same[3 * T:4 * T] = False
        assert torch.equal(y[same], y2[same]) and not torch.equal(y[~same], y2[~same])
        # 3. identity taps (centre 1/3 per view, no BN) reproduce the input exactly
        for conv in (m.shift_conv, m.h_conv, m.w_conv):
            conv.weight.zero_()
            conv.weight.view(Cs, 3)[:, 1] = 1.0
        y3 = m(x)
        assert rel_err(y3[:, :Cs].float().cpu().numpy(), 3 * x[:, :Cs].float().cpu().numpy()) < 1e-2
        # 4. linearity in x without BN/hardswish (fp32 to keep rounding out of the way)
        m.float()
        for conv in (m.shift_conv, m.h_conv, m.w_conv):
            conv.weight.normal_(0, 0.5)
        a = torch.randn(2 * T, C, H, H, device="cuda")
        b = torch.randn(2 * T, C, H, H, device="cuda")
        lhs = m(2 * a + 3 * b)[:, :Cs]
        rhs = 2 * m(a)[:, :Cs] + 3 * m(b)[:, :Cs]
        assert rel_err(lhs.cpu().numpy(), rhs.cpu().numpy()) < 1e-5


def test_mvf_errors():
    from mvfnet_b200 import MVF, _lib
    m = MVF(torch.nn.Identity(), 4, 16, alpha=0.25).cuda()
    with pytest.raises(ValueError):
        m(torch.zeros(6, 16, 4, 4, device="cuda"))                   # 6 frames is not a multiple of T=4
    with pytest.raises(TypeError):
        m(torch.zeros(8, 16, 4, 4, device="cuda", dtype=torch.float16))
    with pytest.raises(RuntimeError):
        m.cpu()(torch.zeros(8, 16, 4, 4))                            # no CPU path
    assert _lib.launch_count() > 0
