"""Host-side wrappers (autograd Functions over the C ABI) for the bottleneck's 1x1 convolutions.

`conv1x1(x, weight)` replaces `nn.Conv2d(kernel_size=1, bias=False)` of Bottleneck.conv1 / conv3 /
downsample (backbones/resnet.py:157-162,179-180,299-303) for bf16 channels_last activations;
`mvf_conv1x1(x, mvf)` is MVF.forward as a whole (MVF.py:104-138): the fused MVF kernel writes only the compact
slab and the GEMM reads its A operand from two tensors, so the reference's cat / contiguous copies never exist.
Forward and input-gradient run on libmvf_b200's tcgen05 GEMM (`conv1x1_gemm`), the weight-gradient on its MN-major
tcgen05 GEMM (`conv1x1_wgrad`).  3x3 convolutions, BatchNorm (+ residual + ReLU), the stem and the max-pool follow.

bf16 operand forms of a weight (plain, transposed for the input-gradient, KRSC / rotated for the 3x3 kernels) are
cached per parameter VERSION (`_wform`): one cast per optimizer step instead of one per use.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch

from . import _lib
from ._lib import ptr
from . import mvf as _mvf


class GemmDesc(C.Structure):
    """mvfb_gemm_desc (include/mvf_b200.h)."""
    _fields_ = [("M", C.c_longlong), ("N", C.c_int), ("K", C.c_int), ("K0", C.c_int),
                ("lda0", C.c_longlong), ("lda1", C.c_longlong), ("ldb", C.c_longlong), ("ldd", C.c_longlong)]


class ConvDesc(C.Structure):
    """mvfb_conv_desc (include/mvf_b200.h)."""
    _fields_ = [(n, C.c_int) for n in ("F", "H", "W", "Cin", "Cout", "stride", "ksize")]


class BnDesc(C.Structure):
    """mvfb_bn_desc (include/mvf_b200.h)."""
    _fields_ = [("M", C.c_longlong), ("C", C.c_int), ("relu", C.c_int), ("training", C.c_int),
                ("eps", C.c_float), ("momentum", C.c_float)]


_declared = False
_VP, _LL = C.c_void_p, C.c_longlong


def _L():
    global _declared
    L = _lib.lib()
    if not _declared:
        L.conv1x1_gemm.restype = C.c_int
        L.conv1x1_gemm.argtypes = [C.POINTER(GemmDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP]
        L.conv1x1_wgrad.restype = C.c_int
        L.conv1x1_wgrad.argtypes = [C.POINTER(GemmDesc), _VP, _VP, _VP, _VP, _VP]
        L.conv3x3_gemm.restype = C.c_int
        L.conv3x3_gemm.argtypes = [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, _VP]
        L.conv3x3_wgrad.restype = C.c_int
        L.conv3x3_wgrad.argtypes = [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP]
        L.bn_stats.restype = C.c_int
        L.bn_stats.argtypes = [C.POINTER(BnDesc), _VP, _LL, _VP, _VP]
        L.bn_apply.restype = C.c_int
        L.bn_apply.argtypes = [C.POINTER(BnDesc), _VP, _LL, _VP, _LL, _VP, _LL, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]
        L.bn_bwd.restype = C.c_int
        L.bn_bwd.argtypes = [C.POINTER(BnDesc), _VP, _LL, _VP, _LL, _VP, _LL, _VP, _VP, _VP, _VP, _LL, _VP, _LL, _VP,
                             _VP, _VP, _VP, _VP]
        L.conv1x1_gemm_bnact.restype = C.c_int
        L.conv1x1_gemm_bnact.argtypes = [C.POINTER(GemmDesc), _VP, _VP, _VP, _VP, _VP, _VP, _LL, C.c_int, _VP, _VP]
        L.conv3x3_gemm_bnact.restype = C.c_int
        L.conv3x3_gemm_bnact.argtypes = [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, C.c_int, _VP, _VP]
        L.conv1x1_gemm_add_cols.restype = C.c_int
        L.conv1x1_gemm_add_cols.argtypes = [C.POINTER(GemmDesc), _VP, _VP, _VP, _VP, _LL, C.c_int, _VP, _VP]
        L.conv1x1s2_dgrad.restype = C.c_int
        L.conv1x1s2_dgrad.argtypes = [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP]
        L.conv3x3s2_dgrad.restype = C.c_int
        L.conv3x3s2_dgrad.argtypes = [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP]
        L.conv1x1_gemm_add.restype = C.c_int
        L.conv1x1_gemm_add.argtypes = [C.POINTER(GemmDesc), _VP, _VP, _VP, _VP, _LL, _VP, _VP]
        L.stem_im2col.restype = C.c_int
        L.stem_im2col.argtypes = [_VP, _VP, _LL, C.c_int, C.c_int, _VP]
        L.maxpool3x3s2_fwd.restype = C.c_int
        L.maxpool3x3s2_fwd.argtypes = [_VP, _VP, _VP, _LL, C.c_int, C.c_int, C.c_int, _VP]
        L.maxpool3x3s2_bwd.restype = C.c_int
        L.maxpool3x3s2_bwd.argtypes = [_VP, _VP, _VP, _LL, C.c_int, C.c_int, C.c_int, _VP]
        L.bn_relu_maxpool_fwd.restype = C.c_int
        L.bn_relu_maxpool_fwd.argtypes = [C.POINTER(BnDesc), _VP, _LL, C.c_int, C.c_int] + [_VP] * 10
        L.bn_relu_maxpool_bwd.restype = C.c_int
        L.bn_relu_maxpool_bwd.argtypes = [C.POINTER(BnDesc), _VP, _VP, _VP, _LL, C.c_int, C.c_int] + [_VP] * 8
        _declared = True
    return L


def _stream():
    # the raw handle of the current stream of the current device: two C calls (torch.cuda.current_stream() builds a Stream
    # object through several Python layers, ~16 us -- 270 launches per step made that a quarter of the host time at B = 12)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


_T = _mvf._Timed          # bench.py's per-family launch timing (no-op unless mvf.timing_begin() was called)

# (id(param), form) -> (weak reference to the parameter, its version, weight epoch, tensor).  `Tensor._version` is
# bumped by every in-place torch update, `_WEIGHT_EPOCH` by FlatSGD.step() (whose kernel updates the flat parameter
# buffer behind autograd's back), so a stale form is never used; the weak reference guards against a dead parameter's id
# being re-used by another tensor.  Forms:
#   "rows"   (Cout, Cin*kh*kw) bf16                     1x1 forward operand B
#   "rowsT"  (Cin, Cout) bf16                           1x1 input-gradient operand B (= W^T)
#   "nchw"   (Cout, Cin, kh, kw) bf16 in the parameter's memory format
#   "krsc"   (Cout, 3, 3, Cin) bf16 contiguous          3x3 forward operand B (a no-op view for channels_last weights)
#   "rot"    (Cin, 3, 3, Cout) bf16, taps rotated 180   3x3 stride-1 input-gradient operand B
#   "s2dgrad" four (Cin, taps*Cout) parity blocks       3x3 stride-2 input-gradient operand (conv3x3s2_dgrad)
#   "fcpad"  (ceil64(NC), K) bf16, zero rows appended   classification head: classes padded to the GEMM's N granularity
#   "fcpadT" (K, ceil64(NC)) bf16                       ... its input-gradient operand
_WFORMS = {}
_WEIGHT_EPOCH = 0
# id(param) -> (weak reference, bf16 view of the parameter kept current by FlatSGD's update kernel)
_BF16_SOURCES = {}


# id(param) -> (weak reference, fp32 view of the parameter's slice of FlatSGD's flat gradient buffer, with the parameter's
# own strides): the weight-gradient kernels write there directly and the view is what autograd receives, so nothing is
# copied when the step packs the gradients (the slice already holds them).
_GRAD_SINKS = {}
_SINK_STEP = 0            # bumped by FlatGrads.zero_(): a sink is written at most once per backward (a weight used twice
                          # in one graph gets an ordinary temporary for its second gradient, which autograd accumulates)


def register_grad_sinks(params, views):
    for p, v in zip(params, views):
        _GRAD_SINKS[id(p)] = [weakref.ref(p), v, -1]


def new_backward_pass():
    """Start of a training step (FlatGrads.zero_()): gradient sinks may be written again, the statistics arena is zeroed
    by ONE fill and handed out from its start, BatchNorm step counters deferred by the previous step are applied."""
    global _SINK_STEP
    _SINK_STEP += 1
    flush_counters()
    for a in _STAT_ARENAS.values():
        a[0].zero_()
        a[1], a[2] = 0, _SINK_STEP


# Per-column (sum, sum of squares) accumulators of the GEMM epilogues: 53 tiny zero-fills per step when allocated one by
# one.  Inside a step protocol (new_backward_pass() at every zero_grad) they are slices of one arena zeroed once.
_STAT_ARENAS = {}         # device -> [fp32 arena, next free element, step it was zeroed for]
_STAT_ARENA_ELEMS = 1 << 18


def _stat_sums(n, device):
    """A zeroed (2, n) fp32 tensor."""
    a = _STAT_ARENAS.get(device)
    if a is None:
        if _SINK_STEP == 0:                                      # nobody drives a step protocol: plain allocation
            return torch.zeros((2, n), dtype=torch.float32, device=device)
        a = _STAT_ARENAS[device] = [torch.zeros(_STAT_ARENA_ELEMS, dtype=torch.float32, device=device), 0, _SINK_STEP]
    need = (2 * n + 31) // 32 * 32                               # keep every slice 128-byte aligned
    if a[2] != _SINK_STEP or a[1] + need > a[0].numel():
        return torch.zeros((2, n), dtype=torch.float32, device=device)
    t = a[0][a[1]:a[1] + 2 * n].view(2, n)
    a[1] += need
    return t


# nn.BatchNorm2d.num_batches_tracked += 1 is one launch per BatchNorm and forward (62 per step).  With a step protocol
# (FlatSGD) the increments are collected and applied by one multi-tensor add in FlatSGD.step() / the next zero_grad.
_DEFER_COUNTERS = False
_PENDING_COUNTERS = []


def defer_counters(on=True):
    global _DEFER_COUNTERS
    flush_counters()
    _DEFER_COUNTERS = bool(on)


def count_batch(bn):
    """Deferred only for a BatchNorm whose weight lives in a FlatSGD's flat buffer: that optimizer's step() flushes."""
    hit = _GRAD_SINKS.get(id(bn.weight)) if _DEFER_COUNTERS and bn.weight is not None else None
    if hit is not None and hit[0]() is bn.weight:
        _PENDING_COUNTERS.append(bn.num_batches_tracked)
    else:
        bn.num_batches_tracked += 1


def flush_counters():
    if _PENDING_COUNTERS:
        with torch.no_grad():
            torch._foreach_add_(_PENDING_COUNTERS, 1)
        _PENDING_COUNTERS.clear()


def _grad_sink(weight, rows_cols=None):
    """The flat-buffer slice to write `weight`'s gradient into, or None.  `rows_cols`: require that the slice, read in
    memory order, is the (rows, cols) row-major matrix the kernel produces."""
    hit = _GRAD_SINKS.get(id(weight))
    if hit is None or hit[0]() is not weight or hit[1].device != weight.device:
        return None
    v = hit[1]
    if rows_cols is not None:
        if weight.dim() == 4 and weight.shape[2] * weight.shape[3] > 1 and not weight.is_contiguous(memory_format=torch.channels_last):
            return None                                        # the kernels emit (Cout, kh, kw, Cin): channels_last only
        if v.numel() != rows_cols[0] * rows_cols[1]:
            return None
    if hit[2] == _SINK_STEP:
        return None
    hit[2] = _SINK_STEP
    return v


def bump_weight_epoch():
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def register_bf16_sources(params, views):
    """FlatSGD: `views[i]` is a bf16 tensor of params[i]'s shape and strides that its step kernel rewrites together with
    the fp32 parameter -- the "nchw" / "rows" forms become views of it (no cast kernel per layer per step)."""
    for p, v in zip(params, views):
        _BF16_SOURCES[id(p)] = (weakref.ref(p), v)
    bump_weight_epoch()


# (id(param), form) -> (weak reference, tensor): derived bf16 forms FlatSGD keeps current itself (one batched transpose
# launch per step, csrc/tail.cu::transpose_tiles): "rowsT" of the 1x1 weights, "rot" of the channels_last 3x3 weights
_BF16_FORMS = {}


def register_bf16_form(param, form, tensor):
    _BF16_FORMS[(id(param), form)] = (weakref.ref(param), tensor)


def _bf16_source(weight):
    hit = _BF16_SOURCES.get(id(weight))
    if hit is not None and hit[0]() is weight and hit[1].device == weight.device:
        return hit[1]
    return None


def _wform(weight, form):
    key = (id(weight), form)
    ver = weight._version
    hit = _WFORMS.get(key)
    if (hit is not None and hit[0]() is weight and hit[1] == ver and hit[2] == _WEIGHT_EPOCH
            and hit[3].device == weight.device):
        return hit[3]
    kept = _BF16_FORMS.get(key)
    if kept is not None and kept[0]() is weight and kept[1].device == weight.device:
        return kept[1]
    if form == "nchw":
        t = _bf16_source(weight)
        if t is None:
            t = weight.detach().to(torch.bfloat16)
    elif form == "rows":
        w = _wform(weight, "nchw")
        t = w.reshape(w.shape[0], -1)
        if t.stride(1) != 1:                                     # a channels_last k x k weight viewed as rows
            t = t.contiguous()
    elif form == "rowsT":
        t = _wform(weight, "rows").t().contiguous()
    elif form == "krsc":
        t = _wform(weight, "nchw").permute(0, 2, 3, 1).contiguous()
    elif form == "rot":
        t = _wform(weight, "nchw").flip(2, 3).permute(1, 2, 3, 0).contiguous()
    elif form == "s2dgrad":
        t = _s2_dgrad_operand(_wform(weight, "nchw"))
    elif form == "fcpad":
        w = _wform(weight, "rows")
        npad = (w.shape[0] + 63) // 64 * 64
        t = torch.zeros((npad, w.shape[1]), dtype=torch.bfloat16, device=w.device)
        t[:w.shape[0]] = w
    elif form == "fcpadT":
        t = _wform(weight, "fcpad").t().contiguous()
    elif form == "stem":
        t = _stem_weight_matrix(weight)
    else:
        raise KeyError(form)
    if len(_WFORMS) > 4096:                                    # models that came and went (test suites)
        for k in [k for k, v in _WFORMS.items() if v[0]() is None]:
            del _WFORMS[k]
        for k in [k for k, v in _BF16_SOURCES.items() if v[0]() is None]:
            del _BF16_SOURCES[k]
        for k in [k for k, v in _BF16_FORMS.items() if v[0]() is None]:
            del _BF16_FORMS[k]
    _WFORMS[key] = (weakref.ref(weight), ver, _WEIGHT_EPOCH, t)
    return t


def enabled() -> bool:
    """MVFB_CONV1X1=0 routes the 1x1 convolutions back through torch / cuDNN (A/B measurements only)."""
    return os.environ.get("MVFB_CONV1X1", "1") != "0"


def _rows(t):
    """(F, C, H, W) channels_last tensor -> its (F*H*W, C) row-major view (no copy)."""
    f, c, h, w = t.shape
    return t.permute(0, 2, 3, 1).reshape(f * h * w, c)


def eligible(x, cin, cout):
    return (enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and cin % 64 == 0 and cout % 64 == 0
            and x.is_contiguous(memory_format=torch.channels_last) and x.shape[1] == cin)


def gemm_tn(a1, b, a0=None, k0=0, stats=False, out=None, add=None, add_col0=0):
    """out[m, n] = sum_k A[m, k] b[n, k] (+ add[m, n] for n >= add_col0) with A = [a0[:, :k0] | a1[:, k0:]]; 2-D bf16
    tensors whose last dim is contiguous.  Returns (out, colsum, colsq) -- the sums are None unless `stats`."""
    L = _L()
    m, k = a1.shape
    n = b.shape[0]
    assert b.shape[1] == k and a1.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        out = torch.empty((m, n), dtype=torch.bfloat16, device=a1.device)
    d = GemmDesc()
    d.M, d.N, d.K, d.K0 = m, n, k, k0
    d.lda1, d.ldb, d.ldd = a1.stride(0), b.stride(0), out.stride(0)
    d.lda0 = a0.stride(0) if a0 is not None else 0
    colsum = colsq = None
    if stats:
        sums = _stat_sums(n, a1.device)
        colsum, colsq = sums[0], sums[1]
        colsum.sums_buffer = sums                                # the (2, N) tensor itself (its ._base may be a whole arena)
    nbytes = 2 * (m * k + m * n + n * k)
    if add is not None:
        assert not stats and add.shape == (m, n) and add.stride(1) == 1 and add.dtype == torch.bfloat16
        with _T("gemm1x1", nbytes=nbytes + 2 * m * n, flops=2 * m * n * k):
            if add_col0:
                rc = L.conv1x1_gemm_add_cols(C.byref(d), ptr(a0), ptr(a1), ptr(b), ptr(add), add.stride(0), add_col0,
                                             ptr(out), _stream())
            else:
                rc = L.conv1x1_gemm_add(C.byref(d), ptr(a0), ptr(a1), ptr(b), ptr(add), add.stride(0), ptr(out), _stream())
        _lib.check(rc, "conv1x1_gemm_add")
        return out, None, None
    with _T("gemm1x1", nbytes=nbytes, flops=2 * m * n * k):
        rc = L.conv1x1_gemm(C.byref(d), ptr(a0), ptr(a1), ptr(b), ptr(out), ptr(colsum), ptr(colsq), _stream())
    _lib.check(rc, "conv1x1_gemm")
    return out, colsum, colsq


# ------------------------------------------------------------------------------------------------ inference (eval BN folded)
_BNFOLD = {}


def bn_fold(bn):
    """(scale, shift) fp32 of an eval-mode BatchNorm2d: y = x * scale + shift with scale = gamma / sqrt(running_var + eps),
    shift = beta - running_mean * scale; cached per version of the four tensors."""
    ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, _WEIGHT_EPOCH)
    hit = _BNFOLD.get(id(bn))
    if hit is not None and hit[0]() is bn and hit[1] == ver and hit[2].device == bn.weight.device:
        return hit[2], hit[3]
    with torch.no_grad():
        scale = (bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)).contiguous()
        shift = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
    _BNFOLD[id(bn)] = (weakref.ref(bn), ver, scale, shift)
    return scale, shift


def bn_infer_ok(*bns):
    """Eval-mode BatchNorm2d with running statistics whose channel count the GEMM epilogue takes."""
    return (enabled() and bn_enabled() and not torch.is_grad_enabled()
            and all(type(b) is torch.nn.BatchNorm2d and not b.training and b.track_running_stats and b.affine
                    and b.num_features % 64 == 0 for b in bns))


def infer_eligible(x, *bns):
    """Inference fast path: no autograd, bf16 channels_last activations, every BatchNorm in eval mode with running
    statistics -- the convolution's epilogue applies it (conv1x1_gemm_bnact / conv3x3_gemm_bnact)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
            and x.is_contiguous(memory_format=torch.channels_last) and bn_infer_ok(*bns))


def gemm_bnact(a1, b, scale, shift, relu, res=None, a0=None, k0=0):
    """out = [relu]((A b^T) * scale + shift [+ res]) -- A = [a0[:, :k0] | a1[:, k0:]]; bf16 2-D operands."""
    L = _L()
    m, k = a1.shape
    n = b.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=a1.device)
    d = GemmDesc()
    d.M, d.N, d.K, d.K0 = m, n, k, k0
    d.lda1, d.ldb, d.ldd = a1.stride(0), b.stride(0), n
    d.lda0 = a0.stride(0) if a0 is not None else 0
    with _T("gemm1x1", nbytes=2 * (m * k + m * n + n * k + (m * n if res is not None else 0)), flops=2 * m * n * k):
        rc = L.conv1x1_gemm_bnact(C.byref(d), ptr(a0), ptr(a1), ptr(b), ptr(scale), ptr(shift), ptr(res),
                                  res.stride(0) if res is not None else 0, int(relu), ptr(out), _stream())
    _lib.check(rc, "conv1x1_gemm_bnact")
    return out


def conv1x1_bnact(x, weight, bn, relu, residual=None, slab=None, k0=0):
    """Inference: 1x1 stride-1 convolution + eval BatchNorm (+ residual) (+ ReLU) in one kernel.  `slab`: the compact MVF
    slab that replaces the first k0 input channels (MVF.forward in eval mode)."""
    f, cin, h, w = x.shape
    scale, shift = bn_fold(bn)
    res = _rows(residual) if residual is not None else None
    out = gemm_bnact(_rows(x), _wform(weight, "rows"), scale, shift, relu, res=res,
                     a0=_rows(slab) if slab is not None else None, k0=k0)
    return _nhwc_from_rows(out, f, h, w)


def conv_window_bnact(x, weight, bn, stride, relu, residual=None):
    """Inference: 3x3 / pad 1 (4-D weight) or strided 1x1 (down-sampling) convolution + eval BatchNorm (+ ReLU)."""
    L = _L()
    f, cin, h, w = x.shape
    three = weight.shape[-1] == 3
    wk = _wform(weight, "krsc" if three else "rows")
    cout = wk.shape[0]
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    scale, shift = bn_fold(bn)
    d = ConvDesc()
    d.F, d.H, d.W, d.Cin, d.Cout, d.stride, d.ksize = f, h, w, cin, cout, stride, (3 if three else 1)
    out = torch.empty((f, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    taps, mo = (9 if three else 1), f * ho * wo
    res = residual.permute(0, 2, 3, 1) if residual is not None else None
    with _T("conv3x3" if three else "gemm1x1", nbytes=2 * (f * h * w * cin // (1 if three else stride * stride) + mo * cout
                                                           + taps * cin * cout), flops=2 * mo * cout * taps * cin):
        rc = L.conv3x3_gemm_bnact(C.byref(d), ptr(x), ptr(wk), ptr(scale), ptr(shift), ptr(res), int(relu), ptr(out), _stream())
    _lib.check(rc, "conv3x3_gemm_bnact")
    return out.permute(0, 3, 1, 2)


def stem_bnact(x, weight, bn):
    """Inference stem: im2col + GEMM with eval BatchNorm + ReLU in the epilogue (backbones/resnet.py:481-483)."""
    L = _L()
    f, _, h, w = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    a = torch.empty((f * ho * wo, STEM_KP), dtype=torch.bfloat16, device=x.device)
    with _T("stem_im2col", nbytes=2 * (xb.numel() + a.numel())):
        _lib.check(L.stem_im2col(ptr(xb), ptr(a), f, h, w, _stream()), "stem_im2col")
    scale, shift = bn_fold(bn)
    out = gemm_bnact(a, _wform(weight, "stem"), scale, shift, True)
    return _nhwc_from_rows(out, f, ho, wo)


def add_fusion_enabled() -> bool:
    """The two gradients of a Bottleneck's input (through conv1 and through the identity path, backbones/resnet.py:211-213,
    238) are summed in the epilogue of conv1's input-gradient GEMM (conv1x1_gemm_add: the addend tile is staged through
    shared memory, coalesced, while the tile's MMAs run) instead of by an autograd add kernel that streams three full
    tensors.  MVFB_ADDFUSE=0 switches it off (A/B measurements only)."""
    return os.environ.get("MVFB_ADDFUSE", "1") != "0"


def wgrad_enabled() -> bool:
    """MVFB_WGRAD=0 computes 1x1 weight gradients with torch.matmul (A/B measurements only)."""
    return os.environ.get("MVFB_WGRAD", "1") != "0"


def gemm_wgrad(g2, x1, x0=None, k0=0, out=None):
    """dW[n, k] = sum_m g2[m, n] X[m, k] with X = [x0[:, :k0] | x1[:, k0:]] -> fp32 (N, K).  bf16 row-major inputs.
    `out`: fp32 tensor of N*K elements whose memory receives the row-major result (a flat-gradient-buffer slice)."""
    if not wgrad_enabled():
        gt = g2.t()
        if x0 is None:
            return torch.matmul(gt, x1).float()
        return torch.cat([torch.matmul(gt, x0), torch.matmul(gt, x1[:, k0:])], dim=1).float()
    L = _L()
    m, n = g2.shape
    k = x1.shape[1]
    dw = out if out is not None else torch.empty((n, k), dtype=torch.float32, device=g2.device)
    d = GemmDesc()
    d.M, d.N, d.K, d.K0 = m, n, k, k0
    d.lda1, d.ldb, d.ldd = x1.stride(0), g2.stride(0), k
    d.lda0 = x0.stride(0) if x0 is not None else 0
    with _T("wgrad1x1", nbytes=2 * (m * n + m * k) + 4 * n * k, flops=2 * m * n * k):
        rc = L.conv1x1_wgrad(C.byref(d), ptr(g2), ptr(x0), ptr(x1), ptr(dw), _stream())
    _lib.check(rc, "conv1x1_wgrad")
    return dw


def _nhwc_from_rows(rows, f, h, w):
    return rows.view(f, h, w, rows.shape[1]).permute(0, 3, 1, 2)


class _Conv1x1(torch.autograd.Function):
    """1x1 stride-1 convolution on the tcgen05 GEMM.  `passthrough`: also returns an alias of x -- the Bottleneck
    uses THAT as its identity path, so both gradients of the block input arrive in this node and the input-gradient
    GEMM adds dL/d(identity) in its epilogue (conv1x1_gemm_add) instead of autograd running a full-tensor add."""

    @staticmethod
    def forward(ctx, x, weight, stats, passthrough):
        f, cin, h, w = x.shape
        wb = _wform(weight, "rows")
        out, colsum, _ = gemm_tn(_rows(x), wb, stats=stats)
        ctx.save_for_backward(x, wb)
        ctx.weight = weight
        ctx.stats, ctx.passthrough = stats, passthrough
        outs = [_nhwc_from_rows(out, f, h, w)]
        if stats:
            sums = colsum.sums_buffer     # the (2, N) buffer
            ctx.mark_non_differentiable(sums)
            ctx.set_materialize_grads(False)   # no zero "gradient" tensor for the statistics output (a fill launch per layer)
            outs.append(sums)
        if passthrough:
            outs.append(x.view_as(x))
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, g, *rest):
        x, wb = ctx.saved_tensors
        f, cin, h, w = x.shape
        g_id = rest[-1] if ctx.passthrough else None
        if g is None:                                            # the convolution's output was not used: only the identity path
            return g_id, None, None, None
        g = g.contiguous(memory_format=torch.channels_last)
        g2 = _rows(g)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            add = None
            if g_id is not None:
                add = _rows(g_id.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
            dx2, _, _ = gemm_tn(g2, _wform(ctx.weight, "rowsT"), add=add)   # dX = dY W (+ dL/d identity)  ==  TN GEMM against W^T
            dx = _nhwc_from_rows(dx2, f, h, w)
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(ctx.weight, (wb.shape[0], cin))
            dw = gemm_wgrad(g2, _rows(x), out=sink)
            dw = sink if sink is not None else dw.view(wb.shape[0], cin, 1, 1)
        return dx, dw, None, None


def conv1x1(x, weight, stats=False, passthrough=False):
    """1x1 stride-1 bias-free convolution of a bf16 channels_last tensor on the tcgen05 GEMM.  With `stats` also
    returns the (2, Cout) per-channel (sum, sum of squares) of the output, accumulated in the GEMM epilogue; with
    `passthrough` the last element of the result is an alias of x to be used as the residual identity."""
    return _Conv1x1.apply(x, weight, stats, passthrough)


class _MVFConv1x1(torch.autograd.Function):
    """MVF.forward (MVF.py:104-138) in two launches: fused MVF kernel -> compact slab; K-split GEMM.  `passthrough`: also
    returns an alias of x that the Bottleneck uses as its identity path, so that the identity's gradient arrives HERE and
    is summed inside the kernels (conv1x1_gemm_add_cols for the untouched channels, mvf_bwd_add for the slab) instead of
    by an autograd add over the full tensor."""

    @staticmethod
    def forward(ctx, x, weight, wt, wh, ww, gamma, beta, running_mean, running_var, cfg, stats, passthrough):
        f, c, h, w = x.shape
        slab, xk, layout, save_mean, save_rstd = _mvf.mvf_slab_forward(
            x, cfg, wt, wh, ww, gamma, beta, running_mean, running_var, out="slab")
        wb = _wform(weight, "rows")
        out, colsum, _ = gemm_tn(_rows(xk), wb, a0=_rows(slab), k0=cfg.Cs, stats=stats)
        ctx.cfg, ctx.layout, ctx.weight, ctx.passthrough = cfg, layout, weight, passthrough
        ctx.save_for_backward(xk, slab, wb, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd)
        outs = [_nhwc_from_rows(out, f, h, w)]
        if stats:
            sums = colsum.sums_buffer
            ctx.mark_non_differentiable(sums)
            ctx.set_materialize_grads(False)   # no zero "gradient" tensor for the statistics output (a fill launch per layer)
            outs.append(sums)
        if passthrough:
            outs.append(x.view_as(x))
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, g, *rest):
        L = _lib.lib()
        cfg, layout = ctx.cfg, ctx.layout
        xk, slab, wb, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd = ctx.saved_tensors
        f, c, h, w = xk.shape
        cs = cfg.Cs
        if g is None:                                            # the block's main path was not used: only the identity path
            return (rest[-1] if ctx.passthrough else None,) + (None,) * 11
        g = g.contiguous(memory_format=torch.channels_last)
        g2 = _rows(g)
        d = _mvf._make_desc(xk, layout, cfg)
        g_id = rest[-1] if ctx.passthrough else None
        fuse_add = None
        if g_id is not None:
            g_id = g_id.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            if _lib.plan(d, True) == "sweep":
                fuse_add = _rows(g_id)
        # dL/dx' (all C channels) = dY W (+ the identity path's gradient on the untouched channels); the slab columns are
        # then rewritten IN PLACE by mvf_bwd, which reads a frame of g completely before it writes that frame's dx
        # (kernel contract, include/mvf_b200.h) and adds the identity path's gradient of the slab channels itself
        dxp, _, _ = gemm_tn(g2, _wform(ctx.weight, "rowsT"), add=fuse_add, add_col0=cs if fuse_add is not None else 0)
        dw = None
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(ctx.weight, (wb.shape[0], c))
            dw = gemm_wgrad(g2, _rows(xk), x0=_rows(slab), k0=cs, out=sink)
            dw = sink if sink is not None else dw.view(wb.shape[0], c, 1, 1)
        dx = _nhwc_from_rows(dxp, f, h, w)
        dev = xk.device
        taps = torch.empty((3, cs, 3), dtype=torch.float32, device=dev)       # back to back: one memset in the kernel tier
        dwt = taps[0]
        dwh = taps[1] if (wh is not None and wh.data_ptr() != wt.data_ptr()) else None
        dww = taps[2] if (ww is not None and ww.data_ptr() != wt.data_ptr()) else None
        dgamma = torch.empty(cs, dtype=torch.float32, device=dev) if cfg.use_hs else None
        dbeta = torch.empty_like(dgamma) if cfg.use_hs else None
        nbytes = L.mvf_bwd_workspace_bytes(C.byref(d))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        args = (C.byref(d), ptr(dxp), c, ptr(xk), ptr(dxp), c, ptr(wt), ptr(wh), ptr(ww), ptr(gamma), ptr(beta),
                ptr(running_mean), ptr(running_var), ptr(save_mean), ptr(save_rstd), ptr(dwt), ptr(dwh), ptr(dww),
                ptr(dgamma), ptr(dbeta), ptr(ws), ws.numel())
        with _mvf._Timed("mvf_bwd", f * cs * h * w, xk.element_size()):
            if fuse_add is not None:
                rc = L.mvf_bwd_add(*args, ptr(fuse_add), _mvf._stream())
            else:
                rc = L.mvf_bwd(*args, _mvf._stream())
        _lib.check(rc, "mvf_bwd")
        if g_id is not None and fuse_add is None:
            dx = dx + g_id                                       # a tier without the fused addend: autograd's add, here
        return dx, dw, dwt, dwh, dww, dgamma, dbeta, None, None, None, None, None


def mvf_conv1x1(x, weight, wt, wh, ww, gamma, beta, running_mean, running_var, cfg, stats=False, passthrough=False):
    return _MVFConv1x1.apply(x, weight, wt, wh, ww, gamma, beta, running_mean, running_var, cfg, stats, passthrough)


# ------------------------------------------------------------------------------------------------ 3x3 convolution
def s2_dgrad_enabled() -> bool:
    """MVFB_S2DGRAD=0 sends the stride-2 3x3 input gradient back to torch / cuDNN (A/B measurements only)."""
    return os.environ.get("MVFB_S2DGRAD", "1") != "0"


def conv3x3_enabled() -> bool:
    """MVFB_CONV3X3=0 routes the 3x3 convolutions back through torch / cuDNN (A/B measurements only)."""
    return os.environ.get("MVFB_CONV3X3", "1") != "0"


def conv3x3_eligible(x, conv):
    return (conv3x3_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
            and type(conv) is torch.nn.Conv2d and conv.kernel_size == (3, 3) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None and conv.stride in ((1, 1), (2, 2))
            and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0 and x.shape[1] == conv.in_channels
            and x.is_contiguous(memory_format=torch.channels_last))


def conv3x3_raw(x, w_krsc, stride, stats=False):
    """x: (F, Cin, H, W) bf16 channels_last; w_krsc: (Cout, 3, 3, Cin) bf16 contiguous -> ((F, Cout, Ho, Wo), sums)."""
    L = _L()
    f, cin, h, w = x.shape
    cout = w_krsc.shape[0]
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    d = ConvDesc()
    d.F, d.H, d.W, d.Cin, d.Cout, d.stride, d.ksize = f, h, w, cin, cout, stride, (3 if w_krsc.dim() == 4 else 1)
    out = torch.empty((f, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    sums = _stat_sums(cout, x.device) if stats else None
    taps = 9 if w_krsc.dim() == 4 else 1
    mo = f * ho * wo
    with _T("conv3x3" if taps == 9 else "gemm1x1", nbytes=2 * (f * h * w * cin // (1 if taps == 9 else stride * stride)
                                                             + mo * cout + taps * cin * cout),
            flops=2 * mo * cout * taps * cin):
        rc = L.conv3x3_gemm(C.byref(d), ptr(x), ptr(w_krsc), ptr(out), ptr(sums[0]) if stats else None,
                            ptr(sums[1]) if stats else None, _stream())
    _lib.check(rc, "conv3x3_gemm")
    return out.permute(0, 3, 1, 2), sums


def _s2_dgrad_operand(w):
    """(Cout, Cin, 3, 3) bf16 -> the four parity operands of conv3x3s2_dgrad back to back (include/mvf_b200.h): parity
    (ph, pw) reads filter rows r(ph) = [1] | [2, 0] and columns s(pw) likewise, (Cin, taps*Cout) each, K order (tap, n)."""
    rows = {0: (1,), 1: (2, 0)}
    parts = []
    for ph in (0, 1):
        for pw in (0, 1):
            taps = torch.stack([w[:, :, r, s] for r in rows[ph] for s in rows[pw]], 0)     # (taps, Cout, Cin)
            parts.append(taps.permute(2, 0, 1).reshape(-1))                                # (Cin, taps, Cout)
    return torch.cat(parts).contiguous()


def conv3x3s2_dgrad_raw(g, weight, h, w):
    """dL/dx (F, Cin, h, w) of the stride-2 3x3 convolution from g (F, Cout, h/2, w/2): four parity GEMMs."""
    L = _L()
    f, cout = g.shape[0], g.shape[1]
    cin = weight.shape[1]
    d = ConvDesc()
    d.F, d.H, d.W, d.Cin, d.Cout, d.stride, d.ksize = f, h, w, cin, cout, 2, 3
    wq = _wform(weight, "s2dgrad")
    dx = torch.empty((f, h, w, cin), dtype=torch.bfloat16, device=g.device)
    with _T("conv3x3", nbytes=2 * (g.numel() + dx.numel() + 9 * cin * cout), flops=2 * f * (h // 2) * (w // 2) * 9 * cin * cout):
        rc = L.conv3x3s2_dgrad(C.byref(d), ptr(g), ptr(wq), ptr(dx), _stream())
    _lib.check(rc, "conv3x3s2_dgrad")
    return dx.permute(0, 3, 1, 2)


def conv3x3_wgrad_raw(g, x, cout, stride, ksize, out=None):
    """dW of the 3x3 (ksize 3 -> (Cout, 3, 3, Cin) fp32) or strided 1x1 (ksize 1 -> (Cout, Cin)) convolution; `out`: fp32
    tensor of as many elements whose memory receives it."""
    f, cin, h, w = x.shape
    d = ConvDesc()
    d.F, d.H, d.W, d.Cin, d.Cout, d.stride, d.ksize = f, h, w, cin, cout, stride, ksize
    shape = (cout, 3, 3, cin) if ksize == 3 else (cout, cin)
    dwk = out if out is not None else torch.empty(shape, dtype=torch.float32, device=x.device)
    mo = g.shape[0] * g.shape[2] * g.shape[3]
    taps = ksize * ksize
    with _T("wgrad3x3" if ksize == 3 else "wgrad1x1", nbytes=2 * (mo * cout + f * h * w * cin) + 4 * taps * cin * cout,
            flops=2 * mo * cout * taps * cin):
        rc = _L().conv3x3_wgrad(C.byref(d), ptr(g), ptr(x), ptr(dwk), _stream())
    _lib.check(rc, "conv3x3_wgrad")
    return dwk


class _Conv3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, stats):
        wb = _wform(weight, "nchw")
        y, sums = conv3x3_raw(x, _wform(weight, "krsc"), stride, stats)
        ctx.stride, ctx.weight = stride, weight
        ctx.save_for_backward(x, wb)
        if not stats:
            return y
        ctx.mark_non_differentiable(sums)
        ctx.set_materialize_grads(False)   # no zero "gradient" tensor for the statistics output (a fill launch per layer)
        return y, sums

    @staticmethod
    def backward(ctx, g, *unused):
        if g is None:
            return None, None, None, None
        x, wb = ctx.saved_tensors
        st = ctx.stride
        g = g.contiguous(memory_format=torch.channels_last)
        dx = dw = None
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_dx and st == 1:
            # stride-1 input gradient = the same convolution with spatially rotated, channel-transposed weights
            dx, _ = conv3x3_raw(g, _wform(ctx.weight, "rot"), 1)          # (Cin, 3, 3, Cout)
            need_dx = False
        elif need_dx and st == 2 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and s2_dgrad_enabled():
            dx = conv3x3s2_dgrad_raw(g, ctx.weight, x.shape[2], x.shape[3])
            need_dx = False
        # own 3x3 wgrad for Cin >= 256 (layer3/4: 1.4x cuDNN's time in isolation); the 9-tap re-read of dY makes it
        # 1.5-3.7x slower on the 56x56 / 28x28 layers (tools/wgrad_probe.py), which stay on the library this round
        own = need_dw and wgrad_enabled() and (x.shape[1] >= 256 or os.environ.get("MVFB_WGRAD3X3") == "all")
        sink = _grad_sink(ctx.weight, (wb.shape[0], 9 * x.shape[1])) if need_dw else None
        if own:
            dw = conv3x3_wgrad_raw(g, x, wb.shape[0], st, 3, out=sink)
            dw = sink if sink is not None else dw.permute(0, 3, 1, 2)      # (Cout, Cin, 3, 3) view of KRSC
            need_dw = False
        if need_dx or need_dw:
            wcl = wb.contiguous(memory_format=torch.channels_last)
            r = torch.ops.aten.convolution_backward(g, x, wcl, None, [st, st], [1, 1], [1, 1], False, [0, 0], 1,
                                                    [need_dx, need_dw, False])
            if need_dx:
                dx = r[0]
            if need_dw:
                if sink is not None:
                    sink.copy_(r[1])
                    dw = sink
                else:
                    dw = r[1].float()
        return dx, dw, None, None


class _Conv1x1Strided(torch.autograd.Function):
    """The stride-2 1x1 down-sampling convolution (make_res_layer, resnet.py:299-303): forward and weight gradient
    gather the strided pixels with TMA im2col (1x1 window); the input gradient is the plain GEMM on the compact
    gradient, scattered to the even pixels of a zeroed tensor."""

    @staticmethod
    def forward(ctx, x, weight, stride, stats):
        wb = _wform(weight, "rows")
        y, sums = conv3x3_raw(x, wb, stride, stats)
        ctx.stride, ctx.weight = stride, weight
        ctx.save_for_backward(x, wb)
        if not stats:
            return y
        ctx.mark_non_differentiable(sums)
        ctx.set_materialize_grads(False)   # no zero "gradient" tensor for the statistics output (a fill launch per layer)
        return y, sums

    @staticmethod
    def backward(ctx, g, *unused):
        if g is None:
            return None, None, None, None
        x, wb = ctx.saved_tensors
        st = ctx.stride
        f, cin, h, w = x.shape
        g = g.contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            cout = wb.shape[0]
            if st == 2 and h % 2 == 0 and w % 2 == 0 and cin % 64 == 0 and cout % 64 == 0 and s2_dgrad_enabled():
                # the GEMM's epilogue scatters its rows to the even pixels and zeroes the rest of each 2 x 2 cell
                d = ConvDesc()
                d.F, d.H, d.W, d.Cin, d.Cout, d.stride, d.ksize = f, h, w, cin, cout, 2, 1
                dx = torch.empty((f, h, w, cin), dtype=torch.bfloat16, device=x.device)
                with _T("gemm1x1", nbytes=2 * (g.numel() + dx.numel() + cin * cout), flops=2 * g.numel() * cin):
                    rc = _L().conv1x1s2_dgrad(C.byref(d), ptr(g), ptr(_wform(ctx.weight, "rowsT")), ptr(dx), _stream())
                _lib.check(rc, "conv1x1s2_dgrad")
            else:
                dxc, _, _ = gemm_tn(_rows(g), _wform(ctx.weight, "rowsT"))           # (F*Ho*Wo, Cin)
                dx = torch.zeros((f, h, w, cin), dtype=torch.bfloat16, device=x.device)
                dx[:, ::st, ::st, :] = dxc.view(f, g.shape[2], g.shape[3], cin)
            dx = dx.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(ctx.weight, (wb.shape[0], cin))
            dw = conv3x3_wgrad_raw(g, x, wb.shape[0], st, 1, out=sink)
            dw = sink if sink is not None else dw.view(wb.shape[0], cin, 1, 1)
        return dx, dw, None, None


def conv1x1_strided(x, weight, stride, stats=False):
    return _Conv1x1Strided.apply(x, weight, stride, stats)


def conv3x3(x, weight, stride=1, stats=False):
    """3x3 / pad 1 bias-free convolution of a bf16 channels_last tensor: forward and stride-1 input-gradient on the
    tcgen05 implicit GEMM (TMA im2col, or one halo band per tile on the 56x56 / 28x28 layers); stride-2 input gradient =
    four parity GEMMs (conv3x3s2_dgrad); weight gradient on tcgen05 for Cin >= 256, a library call below that."""
    return _Conv3x3.apply(x, weight, stride, stats)


# ------------------------------------------------------------------------------------------------ BatchNorm
# ---------------------------------------------------------------------------------------------- stem
STEM_KP = 192             # patch-matrix row: column kh*24 + kw*3 + c (21 values + 3 zeros per kernel row, 8 groups)


def _stem_weight_matrix(weight):
    """(64, 3, 7, 7) -> the (64, 192) bf16 B operand in stem_im2col's column order."""
    wm = torch.zeros((weight.shape[0], 8, 24), dtype=torch.bfloat16, device=weight.device)
    wm[:, :7, :21] = weight.detach().permute(0, 2, 3, 1).reshape(weight.shape[0], 7, 21)
    return wm.view(weight.shape[0], STEM_KP)


def stem_enabled() -> bool:
    """MVFB_STEM=0 routes the stem convolution and max-pool back through torch / cuDNN (A/B measurements only)."""
    return os.environ.get("MVFB_STEM", "1") != "0"


def stem_eligible(x, conv) -> bool:
    """ResNet.conv1 (3 -> 64, 7x7, stride 2, pad 3, no bias) in the bf16 configuration; the input needs no gradient."""
    bf16 = x.dtype == torch.bfloat16 or (torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16)
    return (stem_enabled() and x.is_cuda and bf16 and x.dim() == 4 and x.shape[1] == 3 and not x.requires_grad
            and type(conv) is torch.nn.Conv2d and conv.in_channels == 3 and conv.out_channels == 64
            and conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3) and conv.bias is None
            and conv.groups == 1 and conv.dilation == (1, 1))


class _StemConv(torch.autograd.Function):
    """conv1 of the ResNet stem as im2col (stem_im2col) + the tcgen05 GEMMs (backbones/resnet.py:424, 481)."""

    @staticmethod
    def forward(ctx, x, weight, stats):
        L = _L()
        f, _, h, w = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)   # (F, H, W, 3) in memory
        a = torch.empty((f * ho * wo, STEM_KP), dtype=torch.bfloat16, device=x.device)
        with _T("stem_im2col", nbytes=2 * (xb.numel() + a.numel())):
            _lib.check(L.stem_im2col(ptr(xb), ptr(a), f, h, w, _stream()), "stem_im2col")
        out, colsum, _ = gemm_tn(a, _wform(weight, "stem"), stats=stats)
        ctx.save_for_backward(a)
        ctx.wdtype = weight.dtype
        y = _nhwc_from_rows(out, f, ho, wo)
        if not stats:
            return y
        sums = colsum.sums_buffer
        ctx.mark_non_differentiable(sums)
        ctx.set_materialize_grads(False)   # no zero "gradient" tensor for the statistics output (a fill launch per layer)
        return y, sums

    @staticmethod
    def backward(ctx, g, *unused):
        (a,) = ctx.saved_tensors
        dw = None
        if g is not None and ctx.needs_input_grad[1]:
            g2 = _rows(g.contiguous(memory_format=torch.channels_last))
            dw = gemm_wgrad(g2, a).view(64, 8, 24)[:, :7, :21].reshape(64, 7, 7, 3).permute(0, 3, 1, 2).to(ctx.wdtype)
        return None, dw, None


def stem_conv(x, weight, stats=False):
    return _StemConv.apply(x, weight, stats)


def maxpool_eligible(x, pool) -> bool:
    ks = pool.kernel_size if isinstance(pool.kernel_size, tuple) else (pool.kernel_size,) * 2
    st = pool.stride if isinstance(pool.stride, tuple) else (pool.stride,) * 2
    pd = pool.padding if isinstance(pool.padding, tuple) else (pool.padding,) * 2
    dl = pool.dilation if isinstance(pool.dilation, tuple) else (pool.dilation,) * 2
    return (stem_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last) and ks == (3, 3) and st == (2, 2) and pd == (1, 1)
            and dl == (1, 1) and not pool.ceil_mode and not pool.return_indices)


class _MaxPool3x3s2(torch.autograd.Function):
    """nn.MaxPool2d(3, 2, 1) on bf16 channels_last tensors; one byte of arg-max per output element."""

    @staticmethod
    def forward(ctx, x):
        L = _L()
        f, c, h, w = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        y = torch.empty((f, ho, wo, c), dtype=torch.bfloat16, device=x.device)
        idx = torch.empty((f, ho, wo, c), dtype=torch.uint8, device=x.device)
        with _T("maxpool", nbytes=2 * x.numel() + 3 * y.numel()):
            _lib.check(L.maxpool3x3s2_fwd(ptr(x), ptr(y), ptr(idx), f, h, w, c, _stream()), "maxpool3x3s2_fwd")
        ctx.save_for_backward(idx)
        ctx.shape = (f, c, h, w)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        L = _L()
        (idx,) = ctx.saved_tensors
        f, c, h, w = ctx.shape
        g = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty((f, h, w, c), dtype=torch.bfloat16, device=g.device)
        with _T("maxpool", nbytes=3 * g.numel() + 2 * dx.numel()):
            _lib.check(L.maxpool3x3s2_bwd(ptr(g), ptr(idx), ptr(dx), f, h, w, c, _stream()), "maxpool3x3s2_bwd")
        return dx.permute(0, 3, 1, 2)


def maxpool3x3s2(x):
    return _MaxPool3x3s2.apply(x)


def bn_enabled() -> bool:
    """MVFB_BN=0 routes BatchNorm / ReLU / residual back through torch (A/B measurements only)."""
    return os.environ.get("MVFB_BN", "1") != "0"


def bn_eligible(x, bn):
    return (bn_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and type(bn) is torch.nn.BatchNorm2d
            and bn.affine and x.shape[1] % 8 == 0 and x.shape[1] <= 2048
            and x.is_contiguous(memory_format=torch.channels_last))


class _BNAct(torch.autograd.Function):
    """y = [relu](BatchNorm2d(x) [+ residual]) on bf16 channels_last tensors (bn_stats / bn_apply / bn_bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, sums, relu, training, eps, momentum):
        L = _L()
        f, c, h, w = x.shape
        xr = _rows(x)
        d = BnDesc()
        d.M, d.C, d.relu, d.training, d.eps, d.momentum = xr.shape[0], c, int(relu), int(training), eps, momentum
        dev = x.device
        y = torch.empty((f, h, w, c), dtype=torch.bfloat16, device=dev)
        save = torch.empty((2, c), dtype=torch.float32, device=dev)
        if training and sums is None:
            sums = torch.empty((2, c), dtype=torch.float32, device=dev)
            with _T("bn_fwd", nbytes=2 * xr.numel()):
                _lib.check(L.bn_stats(C.byref(d), ptr(xr), xr.stride(0), ptr(sums), _stream()), "bn_stats")
        rr = _rows(residual) if residual is not None else None
        g32 = gamma if gamma.dtype == torch.float32 else gamma.float()
        b32 = beta if beta.dtype == torch.float32 else beta.float()
        # ReLU bit mask for the backward (1/16 of y's bytes) -- only when a backward can happen
        mask = None
        if relu and any(ctx.needs_input_grad[:3]):
            mask = torch.empty(xr.shape[0] * (c // 8), dtype=torch.uint8, device=dev)
        ne = xr.numel()
        with _T("bn_fwd", nbytes=2 * ne * (3 if rr is not None else 2) + (ne // 8 if mask is not None else 0)):
            rc = L.bn_apply(C.byref(d), ptr(xr), xr.stride(0), ptr(rr), rr.stride(0) if rr is not None else 0, ptr(y), c,
                            ptr(sums), ptr(g32), ptr(b32), ptr(running_mean), ptr(running_var), ptr(save[0]),
                            ptr(save[1]), ptr(mask), _stream())
        _lib.check(rc, "bn_apply")
        yv = y.permute(0, 3, 1, 2)
        ctx.relu, ctx.training, ctx.has_res, ctx.eps = relu, training, residual is not None, eps
        ctx.save_for_backward(x, yv if (relu and mask is None) else None, g32, save, mask)
        return yv

    @staticmethod
    def backward(ctx, g):
        L = _L()
        x, y, gamma, save, mask = ctx.saved_tensors
        f, c, h, w = x.shape
        g = g.contiguous(memory_format=torch.channels_last)
        gr, xr = _rows(g), _rows(x)
        yr = _rows(y) if y is not None else None
        d = BnDesc()
        d.M, d.C, d.relu, d.training, d.eps, d.momentum = xr.shape[0], c, int(ctx.relu), int(ctx.training), ctx.eps, 0.0
        dev = x.device
        dx = torch.empty((f, h, w, c), dtype=torch.bfloat16, device=dev)
        dres = torch.empty((f, h, w, c), dtype=torch.bfloat16, device=dev) if ctx.has_res else None
        grads = torch.empty((2, c), dtype=torch.float32, device=dev)
        scratch = torch.empty((2, c), dtype=torch.float32, device=dev)
        ne = xr.numel()
        # reduce: reads g, x (, mask | y);  apply: reads g, x (, mask | y), writes dx (, d residual)
        rd = 2 * ne * 2 + (ne // 8 if mask is not None else (2 * ne if yr is not None else 0))
        with _T("bn_bwd", nbytes=2 * rd + 2 * ne * (2 if dres is not None else 1)):
            rc = L.bn_bwd(C.byref(d), ptr(gr), gr.stride(0), ptr(yr), yr.stride(0) if yr is not None else 0, ptr(xr),
                          xr.stride(0), ptr(gamma), ptr(save[0]), ptr(save[1]), ptr(dx), c, ptr(dres), c, ptr(grads[0]),
                          ptr(grads[1]), ptr(scratch), ptr(mask), _stream())
        _lib.check(rc, "bn_bwd")
        dres_v = dres.permute(0, 3, 1, 2) if dres is not None else None
        return dx.permute(0, 3, 1, 2), grads[0], grads[1], None, None, dres_v, None, None, None, None, None


def stem_fuse_enabled() -> bool:
    """MVFB_STEM_FUSE=0 keeps norm1 / ReLU / max-pool of the stem as separate launches (A/B measurements only)."""
    return os.environ.get("MVFB_STEM_FUSE", "1") != "0"


def bn_relu_maxpool_eligible(x, bn, pool) -> bool:
    """norm1 + ReLU + maxpool in one pass: bn_act's and maxpool3x3s2's conditions, even H and W, C / 8 dividing 256."""
    return (stem_fuse_enabled() and bn_eligible(x, bn) and maxpool_eligible(x, pool) and x.shape[2] % 2 == 0
            and x.shape[3] % 2 == 0 and 256 % (x.shape[1] // 8) == 0)


class _BNReluMaxPool(torch.autograd.Function):
    """maxpool3x3s2(relu(BatchNorm2d(x))) on a bf16 channels_last tensor without materialising the activation
    (bn_relu_maxpool_fwd / _bwd, include/mvf_b200.h; backbones/resnet.py:481-484)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, sums, training, eps, momentum):
        L = _L()
        f, c, h, w = x.shape
        d = BnDesc()
        d.M, d.C, d.relu, d.training, d.eps, d.momentum = f * h * w, c, 1, int(training), eps, momentum
        dev = x.device
        xr = _rows(x)
        if training and sums is None:
            sums = torch.empty((2, c), dtype=torch.float32, device=dev)
            with _T("bn_fwd", nbytes=2 * xr.numel()):
                _lib.check(L.bn_stats(C.byref(d), ptr(xr), xr.stride(0), ptr(sums), _stream()), "bn_stats")
        ho, wo = h // 2, w // 2
        y = torch.empty((f, ho, wo, c), dtype=torch.bfloat16, device=dev)
        idx = torch.empty((f, ho, wo, c), dtype=torch.uint8, device=dev)
        save = torch.empty((2, c), dtype=torch.float32, device=dev)
        g32 = gamma if gamma.dtype == torch.float32 else gamma.float()
        b32 = beta if beta.dtype == torch.float32 else beta.float()
        with _T("bn_fwd", nbytes=2 * xr.numel() + 3 * y.numel()):
            rc = L.bn_relu_maxpool_fwd(C.byref(d), ptr(xr), f, h, w, ptr(sums), ptr(g32), ptr(b32), ptr(running_mean),
                                       ptr(running_var), ptr(save[0]), ptr(save[1]), ptr(y), ptr(idx), _stream())
        _lib.check(rc, "bn_relu_maxpool_fwd")
        ctx.training, ctx.eps = training, eps
        ctx.save_for_backward(x, idx, g32, save)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        L = _L()
        x, idx, gamma, save = ctx.saved_tensors
        f, c, h, w = x.shape
        g = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        d = BnDesc()
        d.M, d.C, d.relu, d.training, d.eps, d.momentum = f * h * w, c, 1, int(ctx.training), ctx.eps, 0.0
        dev = x.device
        dx = torch.empty((f, h, w, c), dtype=torch.bfloat16, device=dev)
        grads = torch.empty((2, c), dtype=torch.float32, device=dev)
        scratch = torch.empty((2, c), dtype=torch.float32, device=dev)
        ne, npool = x.numel(), idx.numel()
        # both passes read the pooled gradient, the positions and the convolution output; the second writes dx
        with _T("bn_bwd", nbytes=2 * (3 * npool + 2 * ne) + 2 * ne):
            rc = L.bn_relu_maxpool_bwd(C.byref(d), ptr(g), ptr(idx), ptr(_rows(x)), f, h, w, ptr(gamma), ptr(save[0]),
                                       ptr(save[1]), ptr(dx), ptr(grads[0]), ptr(grads[1]), ptr(scratch), _stream())
        _lib.check(rc, "bn_relu_maxpool_bwd")
        return dx.permute(0, 3, 1, 2), grads[0], grads[1], None, None, None, None, None, None


def bn_relu_maxpool(x, bn, sums=None):
    """`pool(relu(bn(x)))` for nn.MaxPool2d(3, 2, 1); `bn` owns the parameters and running statistics (bn_act's contract)."""
    training = bool(bn.training or not bn.track_running_stats)
    rm = rv = None
    momentum = 0.0
    if bn.track_running_stats:
        rm, rv = bn.running_mean, bn.running_var
        if training:
            momentum = 1.0 / float(int(bn.num_batches_tracked) + 1) if bn.momentum is None else float(bn.momentum)
    if not training:
        sums = None
    y = _BNReluMaxPool.apply(x, bn.weight, bn.bias, rm, rv, sums, training, float(bn.eps), momentum)
    if training and bn.track_running_stats:
        count_batch(bn)
    return y


def bn_act(x, bn, relu=True, residual=None, sums=None):
    """`bn` is the nn.BatchNorm2d that owns the parameters / running statistics (same semantics as calling it,
    then adding `residual`, then ReLU)."""
    training = bool(bn.training or not bn.track_running_stats)
    rm = rv = None
    momentum = 0.0
    if bn.track_running_stats:
        rm, rv = bn.running_mean, bn.running_var
        if training:
            momentum = 1.0 / float(int(bn.num_batches_tracked) + 1) if bn.momentum is None else float(bn.momentum)
    if not training:
        sums = None
    y = _BNAct.apply(x, bn.weight, bn.bias, rm, rv, residual, sums, relu, training, float(bn.eps), momentum)
    if training and bn.track_running_stats:
        count_batch(bn)
    return y
