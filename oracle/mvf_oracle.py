"""CPU oracle (numpy) for the MVFNet hot path: MVF module + ResNet bottleneck arithmetic.

TEST INFRASTRUCTURE ONLY -- never imported by the product package `mvfnet_b200`; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.

This is a closed-form restatement (plain numpy, any float dtype; tests use float64 and float32)
of what the reference computes with stock torch modules.  Each function cites the reference
file:line it follows (paths relative to whwu95/MVFNet @ 0ddc7e2).

Parity pin: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build container by
`oracle/make_golden.py` (imports the unmodified reference through `oracle/mmcv_stub`) and committed
under `tests/golden/`; `tests/test_oracle_golden.py` checks every function here against them.
"""
from __future__ import annotations

import numpy as np

MODES = ("THW", "T", "TH")


# --------------------------------------------------------------------------------------------
# codes/models/common/se_module.py:5-24  HardSigmoid / HardSwish
# --------------------------------------------------------------------------------------------
def hardsigmoid(u):
    """relu6(u + 3) / 6            (se_module.py:11-13)"""
    return np.clip(u + 3.0, 0.0, 6.0) / 6.0


def hardswish(u):
    """u * hardsigmoid(u)          (se_module.py:22-24)"""
    return u * hardsigmoid(u)


def hardswish_grad(u):
    """d/du [u * relu6(u+3)/6] = relu6(u+3)/6 + u * 1[0 < u+3 < 6] / 6  (autograd of se_module.py)"""
    inside = ((u + 3.0) > 0.0) & ((u + 3.0) < 6.0)
    return hardsigmoid(u) + u * inside.astype(u.dtype) / 6.0


# --------------------------------------------------------------------------------------------
# codes/models/modules/MVF.py:104-138   MVF.forward  (everything except the wrapped self.net)
# --------------------------------------------------------------------------------------------
def _shift(a, axis, k):
    """Return b with b[..., i, ...] = a[..., i + k, ...] and zeros outside (conv zero padding)."""
    b = np.zeros_like(a)
    n = a.shape[axis]
    src = [slice(None)] * a.ndim
    dst = [slice(None)] * a.ndim
    if k > 0:
        src[axis] = slice(k, n)
        dst[axis] = slice(0, n - k)
    elif k < 0:
        src[axis] = slice(0, n + k)
        dst[axis] = slice(-k, n)
    b[tuple(dst)] = a[tuple(src)]
    return b


def _view_weights(wt, wh, ww, mode, share):
    """share=True re-uses shift_conv for every view (MVF.py:113-116,125-126)."""
    if share:
        wh = wt
        ww = wt
    if mode == "T":
        wh = ww = None
    elif mode == "TH":
        ww = None
    return wt, wh, ww


def mvf_stencil(xs, wt, wh=None, ww=None):
    """Depthwise 3-tap cross-correlations along T, H, W and their sum (MVF.py:65-81,112-129).

    xs: (N, T, Cs, H, W) slab.  wt/wh/ww: (Cs, 3) taps = shift_conv.weight[c,0,:,0,0],
    h_conv.weight[c,0,0,:,0], w_conv.weight[c,0,0,0,:].  Zero padding 1 on the convolved axis;
    the T axis never crosses a clip.  Add order (t + h) + w as in MVF.py:120.
    """
    def view(w, axis):
        w = w.reshape(1, 1, -1, 1, 1, 3)
        return (w[..., 0] * _shift(xs, axis, -1) + w[..., 1] * xs) + w[..., 2] * _shift(xs, axis, +1)

    z = view(wt, 1)
    if wh is not None:
        z = z + view(wh, 3)
    if ww is not None:
        z = z + view(ww, 4)
    return z


def mvf_stencil_transpose(dz, wt, wh=None, ww=None):
    """Adjoint of `mvf_stencil` w.r.t. xs (autograd of the three Conv3d, MVF.py:118-120)."""
    def view(w, axis):
        w = w.reshape(1, 1, -1, 1, 1, 3)
        return w[..., 0] * _shift(dz, axis, +1) + w[..., 1] * dz + w[..., 2] * _shift(dz, axis, -1)

    dx = view(wt, 1)
    if wh is not None:
        dx = dx + view(wh, 3)
    if ww is not None:
        dx = dx + view(ww, 4)
    return dx


def _tap_grads(dz, xs, axis):
    """dw[c,k] = sum_{n,t,h,w} dz * xs shifted by (k-1) along `axis`."""
    red = (0, 1, 3, 4)
    return np.stack([(dz * _shift(xs, axis, k - 1)).sum(axis=red) for k in range(3)], axis=1)


def mvf_forward(x, n_segment, num_shift, wt, wh=None, ww=None, gamma=None, beta=None,
                running_mean=None, running_var=None, mode="THW", share=False, use_hs=True,
                training=False, eps=1e-5, momentum=0.1):
    """MVF.forward up to (not including) `self.net` (MVF.py:104-137).

    x: (N*T, C, H, W).  Returns dict(out=(N*T,C,H,W), z, mean, var (biased), new_running_mean,
    new_running_var).  Channels >= num_shift pass through bit-exactly (MVF.py:110,135).
    """
    assert mode in MODES
    nt, c, h, w = x.shape
    assert nt % n_segment == 0
    n = nt // n_segment
    out = x.copy()
    res = dict(out=out, z=None, mean=None, var=None, new_running_mean=running_mean,
               new_running_var=running_var)
    if num_shift == 0:                                     # MVF.py:108
        return res
    x5 = x.reshape(n, n_segment, c, h, w)                  # MVF.py:109 (native addressing)
    xs = x5[:, :, :num_shift]                              # MVF.py:110
    wt_, wh_, ww_ = _view_weights(wt, wh, ww, mode, share)
    z = mvf_stencil(xs, wt_, wh_, ww_)
    res["z"] = z
    y = z
    if use_hs:                                             # MVF.py:131-134
        m_count = z.size // num_shift
        if training:                                       # nn.BatchNorm3d train semantics
            mean = z.mean(axis=(0, 1, 3, 4))
            var = z.var(axis=(0, 1, 3, 4))                 # biased, used to normalise
            if running_mean is not None:
                unbiased = var * m_count / max(m_count - 1, 1)
                res["new_running_mean"] = (1 - momentum) * running_mean + momentum * mean
                res["new_running_var"] = (1 - momentum) * running_var + momentum * unbiased
        else:
            mean, var = running_mean, running_var
        res["mean"], res["var"] = mean, var
        bc = (1, 1, -1, 1, 1)
        u = (z - mean.reshape(bc)) / np.sqrt(var.reshape(bc) + eps) * gamma.reshape(bc) + beta.reshape(bc)
        y = hardswish(u)
    out.reshape(n, n_segment, c, h, w)[:, :, :num_shift] = y   # MVF.py:135-137
    return res


def mvf_backward(g, x, n_segment, num_shift, wt, wh=None, ww=None, gamma=None, beta=None,
                 running_mean=None, running_var=None, mode="THW", share=False, use_hs=True,
                 training=False, eps=1e-5):
    """Gradient of `mvf_forward` (autograd of MVF.py:104-137).

    g = dL/d out, (N*T, C, H, W).  Returns dict(dx, dwt, dwh, dww, dgamma, dbeta); with share=True
    the three views' tap gradients are summed into dwt (one Parameter, MVF.py:113-116).
    """
    nt, c, h, w = x.shape
    n = nt // n_segment
    dx = g.copy()                                          # pass-through channels
    res = dict(dx=dx, dwt=None, dwh=None, dww=None, dgamma=None, dbeta=None)
    if num_shift == 0:
        return res
    xs = x.reshape(n, n_segment, c, h, w)[:, :, :num_shift]
    gs = g.reshape(n, n_segment, c, h, w)[:, :, :num_shift]
    wt_, wh_, ww_ = _view_weights(wt, wh, ww, mode, share)
    z = mvf_stencil(xs, wt_, wh_, ww_)
    dz = gs
    if use_hs:
        bc = (1, 1, -1, 1, 1)
        red = (0, 1, 3, 4)
        if training:
            mean, var = z.mean(axis=red), z.var(axis=red)
        else:
            mean, var = running_mean, running_var
        rstd = 1.0 / np.sqrt(var + eps)
        zhat = (z - mean.reshape(bc)) * rstd.reshape(bc)
        u = zhat * gamma.reshape(bc) + beta.reshape(bc)
        du = gs * hardswish_grad(u)
        res["dgamma"] = (du * zhat).sum(axis=red)
        res["dbeta"] = du.sum(axis=red)
        if training:
            m_count = z.size // num_shift
            dz = (gamma * rstd).reshape(bc) * (
                du - res["dbeta"].reshape(bc) / m_count - zhat * res["dgamma"].reshape(bc) / m_count)
        else:
            dz = du * (gamma * rstd).reshape(bc)
    dxs = mvf_stencil_transpose(dz, wt_, wh_, ww_)
    dx.reshape(n, n_segment, c, h, w)[:, :, :num_shift] = dxs
    dwt = _tap_grads(dz, xs, 1)
    dwh = _tap_grads(dz, xs, 3) if wh_ is not None else None
    dww = _tap_grads(dz, xs, 4) if ww_ is not None else None
    if share:
        for extra in (dwh, dww):
            if extra is not None:
                dwt = dwt + extra
        dwh = dww = None
    res.update(dwt=dwt, dwh=dwh, dww=dww)
    return res


# --------------------------------------------------------------------------------------------
# torch.nn.Conv2d / BatchNorm2d / ReLU as used by Bottleneck (backbones/resnet.py:157-186,208-244)
# --------------------------------------------------------------------------------------------
def conv2d(x, w, stride=1, pad=0):
    """Cross-correlation, NCHW x (F,Cin,H,W), w (Cout,Cin,kh,kw), no bias (resnet.py:157-180)."""
    f, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (wd + 2 * pad - kw) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    y = np.zeros((f, cout, ho, wo), dtype=np.result_type(x, w))
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, :, r:r + stride * ho:stride, s:s + stride * wo:stride]
            y += np.einsum("fchw,oc->fohw", patch, w[:, :, r, s], optimize=True)
    return y


def conv2d_backward(dy, x, w, stride=1, pad=0):
    """Returns (dx, dw) of `conv2d`."""
    f, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    ho, wo = dy.shape[2:]
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dxp = np.zeros_like(xp, dtype=np.result_type(dy, w))
    dw = np.zeros_like(w, dtype=np.result_type(dy, x))
    for r in range(kh):
        for s in range(kw):
            sl = (slice(None), slice(None), slice(r, r + stride * ho, stride), slice(s, s + stride * wo, stride))
            dw[:, :, r, s] = np.einsum("fohw,fchw->oc", dy, xp[sl], optimize=True)
            dxp[sl] += np.einsum("fohw,oc->fchw", dy, w[:, :, r, s], optimize=True)
    dx = dxp[:, :, pad:pad + h, pad:pad + wd]
    return dx, dw


def batchnorm2d(x, gamma, beta, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d (common/norm.py:4-10,66): batch stats (biased var) in train, running in eval."""
    red = (0, 2, 3)
    bc = (1, -1, 1, 1)
    if training:
        mean, var = x.mean(axis=red), x.var(axis=red)
        m = x.size // x.shape[1]
        new_rm = (1 - momentum) * running_mean + momentum * mean
        new_rv = (1 - momentum) * running_var + momentum * var * m / max(m - 1, 1)
    else:
        mean, var, new_rm, new_rv = running_mean, running_var, running_mean, running_var
    rstd = 1.0 / np.sqrt(var + eps)
    y = (x - mean.reshape(bc)) * rstd.reshape(bc) * gamma.reshape(bc) + beta.reshape(bc)
    return y, dict(mean=mean, rstd=rstd, new_running_mean=new_rm, new_running_var=new_rv)


def batchnorm2d_backward(dy, x, gamma, mean, rstd, training):
    red = (0, 2, 3)
    bc = (1, -1, 1, 1)
    xhat = (x - mean.reshape(bc)) * rstd.reshape(bc)
    dgamma = (dy * xhat).sum(axis=red)
    dbeta = dy.sum(axis=red)
    if training:
        m = x.size // x.shape[1]
        dx = (gamma * rstd).reshape(bc) * (dy - dbeta.reshape(bc) / m - xhat * dgamma.reshape(bc) / m)
    else:
        dx = dy * (gamma * rstd).reshape(bc)
    return dx, dgamma, dbeta


def relu(x):
    return np.maximum(x, 0)


def bottleneck_forward(x, p, stride, training, n_segment=None, mvf=None):
    """Bottleneck.forward (backbones/resnet.py:208-244), style='pytorch' (stride on conv2, :151-153).

    p: dict with conv1/conv2/conv3 weights (Cout,Cin,k,k), bn{1,2,3}_{gamma,beta,rm,rv}, optional
    ds_w + ds_{gamma,beta,rm,rv} (make_res_layer downsample, resnet.py:279-304).
    mvf: optional dict(num_shift, wt, wh, ww, gamma, beta, rm, rv, mode, share, use_hs) -> conv1 is
    the MVF wrapper (MVF.py:38-39,138).  Returns (out, cache) with every intermediate.
    """
    c = {}
    x1 = x
    if mvf is not None:
        r = mvf_forward(x, n_segment, mvf["num_shift"], mvf["wt"], mvf.get("wh"), mvf.get("ww"),
                        mvf.get("gamma"), mvf.get("beta"), mvf.get("rm"), mvf.get("rv"),
                        mode=mvf.get("mode", "THW"), share=mvf.get("share", False),
                        use_hs=mvf.get("use_hs", True), training=training)
        x1 = r["out"]
        c["mvf"] = r
    c["x1"] = x1
    c["y1"] = conv2d(x1, p["conv1"])
    c["a1"], c["s1"] = batchnorm2d(c["y1"], p["bn1_gamma"], p["bn1_beta"], p["bn1_rm"], p["bn1_rv"], training)
    c["r1"] = relu(c["a1"])
    c["y2"] = conv2d(c["r1"], p["conv2"], stride=stride, pad=1)
    c["a2"], c["s2"] = batchnorm2d(c["y2"], p["bn2_gamma"], p["bn2_beta"], p["bn2_rm"], p["bn2_rv"], training)
    c["r2"] = relu(c["a2"])
    c["y3"] = conv2d(c["r2"], p["conv3"])
    c["a3"], c["s3"] = batchnorm2d(c["y3"], p["bn3_gamma"], p["bn3_beta"], p["bn3_rm"], p["bn3_rv"], training)
    identity = x
    if "ds_w" in p:
        c["yd"] = conv2d(x, p["ds_w"], stride=stride)
        identity, c["sd"] = batchnorm2d(c["yd"], p["ds_gamma"], p["ds_beta"], p["ds_rm"], p["ds_rv"], training)
    c["pre"] = c["a3"] + identity
    out = relu(c["pre"])
    return out, c


def bottleneck_backward(dout, x, p, stride, training, cache, n_segment=None, mvf=None):
    """Gradient of `bottleneck_forward`: returns (dx, grads dict keyed like p / 'mvf_*')."""
    c = cache
    g = {}
    dpre = dout * (c["pre"] > 0)
    d3, g["bn3_gamma"], g["bn3_beta"] = batchnorm2d_backward(dpre, c["y3"], p["bn3_gamma"], c["s3"]["mean"], c["s3"]["rstd"], training)
    dr2, g["conv3"] = conv2d_backward(d3, c["r2"], p["conv3"])
    da2 = dr2 * (c["a2"] > 0)
    d2, g["bn2_gamma"], g["bn2_beta"] = batchnorm2d_backward(da2, c["y2"], p["bn2_gamma"], c["s2"]["mean"], c["s2"]["rstd"], training)
    dr1, g["conv2"] = conv2d_backward(d2, c["r1"], p["conv2"], stride=stride, pad=1)
    da1 = dr1 * (c["a1"] > 0)
    d1, g["bn1_gamma"], g["bn1_beta"] = batchnorm2d_backward(da1, c["y1"], p["bn1_gamma"], c["s1"]["mean"], c["s1"]["rstd"], training)
    dx1, g["conv1"] = conv2d_backward(d1, c["x1"], p["conv1"])
    if mvf is not None:
        r = mvf_backward(dx1, x, n_segment, mvf["num_shift"], mvf["wt"], mvf.get("wh"), mvf.get("ww"),
                         mvf.get("gamma"), mvf.get("beta"), mvf.get("rm"), mvf.get("rv"),
                         mode=mvf.get("mode", "THW"), share=mvf.get("share", False),
                         use_hs=mvf.get("use_hs", True), training=training)
        dx = r["dx"]
        for k in ("dwt", "dwh", "dww", "dgamma", "dbeta"):
            g["mvf_" + k] = r[k]
    else:
        dx = dx1
    if "ds_w" in p:
        dd, g["ds_gamma"], g["ds_beta"] = batchnorm2d_backward(dpre, c["yd"], p["ds_gamma"], c["sd"]["mean"], c["sd"]["rstd"], training)
        dxd, g["ds_w"] = conv2d_backward(dd, x, p["ds_w"], stride=stride)
        dx = dx + dxd
    else:
        dx = dx + dpre
    return dx, g


# --------------------------------------------------------------------------------------------
# The ends of the step; pinned by tests/golden/tail_cases.npz
# --------------------------------------------------------------------------------------------
def normalize_format(frames_u8, mean, std, to_rgb=True):
    """Normalize + FormatShape('NCHW') of the data pipeline (datasets/pipelines/augmentations.py:343-396,
    formating.py:134-185): frames (M, H, W, 3) uint8 -> (M, 3, H, W) float32, (float32(x) [BGR -> RGB] - mean) * (1 / std)
    with the channel statistics applied AFTER the optional swap, as cv2.subtract / cv2.multiply do in place."""
    x = frames_u8.astype(np.float32)
    if to_rgb:
        x = x[..., ::-1]
    # Normalize.__init__ stores mean / std as float32 (augmentations.py:357-358); imnormalize widens THOSE to float64
    mean = np.asarray(mean, dtype=np.float32).astype(np.float64).reshape(1, 1, 1, 3)
    stdinv = 1.0 / np.asarray(std, dtype=np.float32).astype(np.float64).reshape(1, 1, 1, 3)
    y = ((x.astype(np.float64) - mean).astype(np.float32).astype(np.float64) * stdinv).astype(np.float32)
    return np.ascontiguousarray(y.transpose(0, 3, 1, 2))


def head_loss(x, w, b, labels, num_seg, keep=None, p=0.0):
    """TSNClsHead.forward + BaseHead.loss and their gradients (heads/tsn_clshead.py:71-98, heads/base.py:40-45,
    segmental_consensuses/simple_consensus.py:41-61): x (F, C, h, w) -> adaptive average pool -> dropout (keep mask
    (F, C) given, scaled by 1/(1-p)) -> Linear(w (NC, C), b) -> mean over the num_seg frames of a clip -> mean
    cross-entropy over the clips.  Returns dict(loss, score, dx, dw, db)."""
    f, c, h, wd = x.shape
    nb = f // num_seg
    feat = x.mean(axis=(2, 3))
    scale = np.ones_like(feat) if keep is None else keep.astype(x.dtype) / (1.0 - p)
    fd = feat * scale
    logits = fd @ w.T + b
    score = logits.reshape(nb, num_seg, -1).mean(axis=1)
    m = score.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(score - m).sum(axis=1))
    lab = np.asarray(labels).reshape(-1)
    loss = float((lse - score[np.arange(nb), lab]).mean())
    soft = np.exp(score - lse[:, None])
    ds = soft.copy()
    ds[np.arange(nb), lab] -= 1.0
    ds /= nb
    dlogits = np.repeat(ds / num_seg, num_seg, axis=0)
    dw = dlogits.T @ fd
    db = dlogits.sum(axis=0)
    dfeat = (dlogits @ w) * scale
    dx = np.broadcast_to((dfeat / (h * wd))[:, :, None, None], x.shape).copy()
    return dict(loss=loss, score=score, dx=dx, dw=dw, db=db)


def sgd_step(p, m, g, lr, momentum, weight_decay, nesterov, max_norm=None, world=1):
    """DistOptimizerHook.after_train_iter after the all-reduce (core/dist_utils.py:29-32, 59-67): g is the SUM over
    ranks; / world, clip_grad_norm_(max_norm, 2) over ALL parameters (lists of arrays), then torch.optim.SGD with a
    zero-initialised momentum buffer.  Returns (new p list, new m list, total norm)."""
    g = [gi / world for gi in g]
    total = float(np.sqrt(sum(float((gi.astype(np.float64) ** 2).sum()) for gi in g)))
    if max_norm is not None:
        coef = min(1.0, max_norm / (total + 1e-6))
        g = [gi * coef for gi in g]
    newp, newm = [], []
    for pi, mi, gi in zip(p, m, g):
        gi = gi + weight_decay * pi
        mi = momentum * mi + gi
        gi = gi + momentum * mi if nesterov else mi
        newp.append(pi - lr * gi)
        newm.append(mi)
    return newp, newm, total
