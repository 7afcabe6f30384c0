"""The ends of the training step around the bottleneck stack, on libmvf_b200's kernels:

* `preprocess_frames`  -- Normalize + FormatShape + ToTensor of the reference's data pipeline
  (codes/datasets/pipelines/augmentations.py:343-396, formating.py:134-185) on the GPU, from decoded uint8 frames:
  the host ships 1 byte per sample instead of the float32 (B, T, 3, H, W) wire format, and the result is already the
  bf16 NHWC tensor the stem reads (no cast / permute kernels).
* `head_loss`          -- TSNClsHead.forward + BaseHead.loss (codes/models/heads/tsn_clshead.py:71-98, heads/base.py:40-45):
  average pool -> dropout -> Linear (tcgen05 GEMM, classes padded to a multiple of 64) -> mean over the T segments ->
  cross-entropy, forward and backward, six launches instead of ~25 ATen ones.
* `FlatSGD`            -- what DistOptimizerHook.after_train_iter does after backward (codes/core/dist_utils.py:59-67 with
  r50_dense.py:152-154): ONE all-reduce of the flat gradient buffer, then `/ world`, clip-norm, weight decay, momentum
  and the Nesterov update in one pass over flat fp32 buffers, emitting the bf16 weights of the next forward.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ptr
from . import mvf as _mvf
from . import ops

_declared = False
_VP, _LL, _F, _I = C.c_void_p, C.c_longlong, C.c_float, C.c_int


def _L():
    global _declared
    L = ops._L()
    if not _declared:
        L.preprocess_u8.restype = _I
        L.preprocess_u8.argtypes = [_VP, _VP, _LL, C.POINTER(_F), C.POINTER(_F), _I, _VP]
        for fn in (L.head_pool_fwd, L.head_pool_bwd):
            fn.restype = _I
            fn.argtypes = [_VP, _VP, _LL, _I, _I, _F, C.c_ulonglong, _VP, _VP]
        L.head_ce_fwd.restype = _I
        L.head_ce_fwd.argtypes = [_VP, _LL, _VP, _VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]
        L.head_ce_bwd.restype = _I
        L.head_ce_bwd.argtypes = [_VP, _VP, _VP, _LL, _I, _I, _I, _VP]
        L.flat_sqnorm.restype = _I
        L.flat_sqnorm.argtypes = [_VP, _LL, _VP, _VP]
        L.sgd_nesterov_step.restype = _I
        L.sgd_nesterov_step.argtypes = [_VP, _VP, _VP, _VP, _LL, _VP, _VP, _F, _F, _F, _F, _F, _I, _VP]
        L.transpose_tiles.restype = _I
        L.transpose_tiles.argtypes = [_VP, _LL, _VP]
        _declared = True
    return L


_stream = ops._stream
_T = _mvf._Timed


# ------------------------------------------------------------------------------------------------ input pre-processing
IMG_NORM_MEAN = (123.675, 116.28, 103.53)         # configs/MVFNet/K400/*.py img_norm_cfg
IMG_NORM_STD = (58.395, 57.12, 57.375)


def preprocess_frames(frames_u8, mean=IMG_NORM_MEAN, std=IMG_NORM_STD, to_rgb=True):
    """(B, T, H, W, 3) uint8 CUDA frames (decoded images, BGR as cv2 / mmcv deliver them) -> the reference's
    `img_group` (B, T, 3, H, W), normalised, as a bf16 tensor whose memory is NHWC per frame: `Recognizer2D` reshapes it
    to (B*T, 3, H, W) channels_last without touching a byte."""
    if not (frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 5 and frames_u8.shape[-1] == 3):
        raise TypeError("preprocess_frames takes a (B, T, H, W, 3) uint8 CUDA tensor")
    x = frames_u8.contiguous()
    b, t, h, w, _ = x.shape
    y = torch.empty((b, t, h, w, 3), dtype=torch.bfloat16, device=x.device)
    m = (_F * 3)(*[float(v) for v in mean])
    s = (_F * 3)(*[float(v) for v in std])
    with _T("preprocess", nbytes=3 * x.numel()):
        rc = _L().preprocess_u8(ptr(x), ptr(y), b * t * h * w, m, s, int(bool(to_rgb)), _stream())
    _lib.check(rc, "preprocess_u8")
    return y.permute(0, 1, 4, 2, 3)


# ------------------------------------------------------------------------------------------------ head + loss
def _fc_padded(weight):
    """(NC, K) fp32 Linear weight -> (ceil64(NC), K) bf16 zero-padded GEMM operand (cached per parameter version)."""
    return ops._wform(weight, "fcpad")


def head_loss_eligible(x, head, labels):
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
            and x.shape[1] % 64 == 0 and head.spatial_size == -1 and head.spatial_type == 'avg' and not head.with_avg_pool
            and head.consensus_type == 'avg' and not head.extract_feat and not head.fcn_testing
            and head.new_fc.bias is not None and head.num_classes <= 512 and labels is not None and ops.enabled())


class _HeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, labels, num_seg, p, seed):
        L = _L()
        ctx.seed_dev = DEVICE_SEED
        f, c, h, w = x.shape
        hw, b = h * w, f // num_seg
        nc = weight.shape[0]
        dev = x.device
        feat = torch.empty((f, c), dtype=torch.bfloat16, device=dev)
        with _T("head", nbytes=2 * (x.numel() + feat.numel())):
            _lib.check(L.head_pool_fwd(ptr(x), ptr(feat), f, hw, c, p, seed, ptr(ctx.seed_dev), _stream()), "head_pool_fwd")
        wpad = _fc_padded(weight)
        logits, _, _ = ops.gemm_tn(feat, wpad)                                  # (F, ceil64(NC)) bf16
        ds = torch.empty((b, nc), dtype=torch.float32, device=dev)
        dbias = torch.empty(nc, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lab = labels.reshape(-1).to(torch.int64).contiguous()
        b32 = bias.detach().float()
        with _T("head", nbytes=2 * logits.numel()):
            rc = L.head_ce_fwd(ptr(logits), logits.stride(0), ptr(b32), ptr(lab), b, num_seg, nc, None, ptr(ds), ptr(dbias),
                               ptr(loss), _stream())
        _lib.check(rc, "head_ce_fwd")
        ctx.save_for_backward(feat, ds, dbias)
        ctx.weight, ctx.geo = weight, (f, c, h, w, num_seg, p, seed, logits.shape[1])
        ctx.mark_non_differentiable()
        return loss

    @staticmethod
    def backward(ctx, gout):
        L = _L()
        feat, ds, dbias = ctx.saved_tensors
        f, c, h, w, num_seg, p, seed, ldl = ctx.geo
        weight = ctx.weight
        nc, b = weight.shape[0], f // num_seg
        dev = feat.device
        gout = gout.detach().float().contiguous()
        dlogits = torch.empty((f, ldl), dtype=torch.bfloat16, device=dev)
        with _T("head", nbytes=2 * dlogits.numel()):
            _lib.check(L.head_ce_bwd(ptr(ds), ptr(gout), ptr(dlogits), ldl, b, num_seg, nc, _stream()), "head_ce_bwd")
        dw = ops.gemm_wgrad(dlogits, feat)[:nc]                                  # (NC, K) fp32
        dfeat, _, _ = ops.gemm_tn(dlogits, ops._wform(weight, "fcpadT"))         # (F, K) bf16 = dlogits W
        dx = torch.empty((f, h, w, c), dtype=torch.bfloat16, device=dev)
        with _T("head", nbytes=2 * (dx.numel() + dfeat.numel())):
            _lib.check(L.head_pool_bwd(ptr(dfeat), ptr(dx), f, h * w, c, p, seed, ptr(ctx.seed_dev), _stream()), "head_pool_bwd")
        return dx.permute(0, 3, 1, 2), dw.to(weight.dtype), (dbias * gout).to(weight.dtype), None, None, None, None


# Optional int64 device scalar mixed into the dropout seed by the kernels; `graph.GraphedTrainStep` installs one and
# advances it inside the captured step, so that replays draw fresh masks.
DEVICE_SEED = None


def _next_seed(device):
    """Dropout seed drawn from torch's generator state, so that torch.manual_seed() controls it."""
    return int(torch.randint(0, 2**31 - 1, (1,)).item())


def head_loss(x, head, labels, num_seg):
    """loss_cls of TSNClsHead on the (F, 2048, h, w) bf16 channels_last feature map: the fused replacement of
    `head.loss(head(x, num_seg), labels)`."""
    p = float(head.dropout_ratio) if (head.training and head.dropout is not None) else 0.0
    seed = _next_seed(x.device) if p > 0.0 else 0
    return _HeadLoss.apply(x, head.new_fc.weight, head.new_fc.bias, labels, int(num_seg), p, seed)


# ------------------------------------------------------------------------------------------------ optimizer tail
class FlatSGD:
    """torch.optim.SGD(lr, momentum, weight_decay, nesterov) + clip_grad_norm_(max_norm) + the data-parallel gradient
    average, on flat buffers.

    Construction moves every parameter into ONE flat fp32 buffer (`p.data` becomes a view with the parameter's own
    strides, so channels_last weights stay channels_last), allocates a flat momentum buffer and a flat bf16 copy of the
    parameters (registered with `ops` as the GEMM operands, so the forward launches no cast kernels).  `step()`:
    gradients -> flat buffer (one multi-tensor copy), ONE all-reduce when world > 1, sum of squares, fused update.
    Nothing in `step()` synchronises with the host.  The state_dict has torch.optim.SGD's format.

    One deliberate difference from torch.optim.SGD / the reference's `allreduce_grads` (core/dist_utils.py:38-49): a
    parameter that received NO gradient in a step is treated as having a zero gradient (weight decay and momentum still
    act on it, on one GPU and on many alike), where torch skips it.  Every parameter of the MVFNet recognizers receives a
    gradient in every step, so the trajectories coincide (tests/test_tail_gpu.py); a model with unused branches should
    keep those parameters out of this optimizer."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, nesterov=False, max_norm=None, dampening=0.0):
        from .dist import FlatGrads
        if dampening != 0.0:
            raise NotImplementedError("FlatSGD implements the configs' SGD: dampening must be 0")
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters to optimise")
        dev = self.params[0].device
        if any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise TypeError("FlatSGD takes fp32 parameters on one device")
        self.defaults = dict(lr=lr, momentum=momentum, dampening=0.0, weight_decay=weight_decay, nesterov=nesterov,
                             maximize=False, foreach=None, differentiable=False, fused=None)
        self.max_norm = max_norm
        self.grads = FlatGrads(self.params)
        (self.flat_g,) = self.grads.buffers.values()
        n = self.flat_g.numel()
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_p16 = torch.empty(n, dtype=torch.bfloat16, device=dev)
        self._sq = torch.zeros((), dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)       # total norm of the last step (device)
        self._views16, off = [], 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                dense = _dense_strides(p)
                v = self.flat_p[off:off + k].as_strided(p.shape, dense)
                v.copy_(p)
                p.data = v
                self._views16.append(self.flat_p16[off:off + k].as_strided(p.shape, dense))
                off += k
        self.grads.restride(self.params)
        self.flat_p16.copy_(self.flat_p)
        ops.register_bf16_sources(self.params, self._views16)
        self._build_transposed_forms()
        (views,) = self.grads.views.values()
        ops.register_grad_sinks(self.params, views)            # weight-gradient kernels write into the flat buffer
        ops.defer_counters(True)                               # num_batches_tracked: one multi-tensor add per step()
        self.steps = 0

    def _build_transposed_forms(self):
        """The W^T operands of the input-gradient GEMMs ((Cin, Cout) for 1x1 weights; (Cin, 3, 3, Cout) with the taps
        rotated by 180 degrees for channels_last 3x3 weights) live in a second flat bf16 buffer; a table of 32 x 32
        tiles, built once, lets ONE transpose_tiles launch per step rebuild all of them from the flat bf16 weights."""
        import numpy as np
        jobs, total = [], 0                                     # (param, view16, form, shape of the form)
        for p, v in zip(self.params, self._views16):
            if p.dim() != 4 or p.shape[0] % 64 or p.shape[1] % 64:
                continue
            if p.shape[2] == p.shape[3] == 1:
                jobs.append((p, v, "rowsT", (p.shape[1], p.shape[0])))
            elif p.shape[2] == p.shape[3] == 3 and p.is_contiguous(memory_format=torch.channels_last):
                jobs.append((p, v, "rot", (p.shape[1], 3, 3, p.shape[0])))
                jobs.append((p, v, "s2dgrad", (p.numel(),)))      # whichever stride the layer runs at picks its form
                total += p.numel()
            else:
                continue
            total += p.numel()
        self._tiles, self._ntiles = None, 0
        if not jobs:
            return
        self.flat_p16T = torch.zeros(total, dtype=torch.bfloat16, device=self.flat_p.device)
        rec = np.dtype([("src", "<u8"), ("dst", "<u8"), ("lds", "<i4"), ("ldd", "<i4"), ("rows", "<i4"), ("cols", "<i4"),
                        ("r0", "<i4"), ("c0", "<i4")])
        tiles, off = [], 0
        for p, v, form, shape in jobs:
            cout, cin = p.shape[0], p.shape[1]
            dst = self.flat_p16T[off:off + p.numel()].view(shape)
            off += p.numel()
            ops.register_bf16_form(p, form, dst)
            if form == "s2dgrad":
                # parity (ph, pw) of the stride-2 input gradient (ops._s2_dgrad_operand): (Cin, taps*Cout) blocks with the
                # filter rows [1] | [2, 0] and columns likewise, K order (tap, n)
                rows, base = {0: (1,), 1: (2, 0)}, 0
                for ph in (0, 1):
                    for pw in (0, 1):
                        sel = [(r, s) for r in rows[ph] for s in rows[pw]]
                        for j, (r, s) in enumerate(sel):
                            src = v.data_ptr() + 2 * (r * 3 + s) * cin
                            dptr = dst.data_ptr() + 2 * (base + j * cout)
                            for r0 in range(0, cout, 32):
                                for c0 in range(0, cin, 32):
                                    tiles.append((src, dptr, 9 * cin, len(sel) * cout, cout, cin, r0, c0))
                        base += len(sel) * cout * cin
                continue
            taps = 1 if form == "rowsT" else 9
            for tap in range(taps):                              # source tap (r, s) -> destination tap (2 - r, 2 - s)
                src = v.data_ptr() + 2 * tap * cin
                dptr = dst.data_ptr() + 2 * (taps - 1 - tap) * cout
                for r0 in range(0, cout, 32):
                    for c0 in range(0, cin, 32):
                        tiles.append((src, dptr, taps * cin, taps * cout, cout, cin, r0, c0))
        table = np.array(tiles, dtype=rec)
        self._tiles = torch.from_numpy(table.view(np.uint8).copy()).to(self.flat_p.device)
        self._ntiles = len(tiles)
        self._transpose()

    def _transpose(self):
        if self._ntiles:
            with _T("optimizer", nbytes=4 * self.flat_p16T.numel()):
                _lib.check(_L().transpose_tiles(ptr(self._tiles), self._ntiles, _stream()), "transpose_tiles")

    def zero_grad(self, set_to_none=True):
        self.grads.zero_()

    def step(self, world=None):
        """`world`: number of ranks the gradient is averaged over (default: the process group's size).  The all-reduce
        runs whenever a process group of more than one rank exists; passing `world` without one (tests) only scales."""
        from .dist import get_dist_info
        import torch.distributed as dist
        L = _L()
        ranks = get_dist_info()[1]
        if world is None:
            world = ranks
        self.grads.gather()
        if ranks > 1:
            dist.all_reduce(self.flat_g)                                       # the ONE collective of the step
        d = self.defaults
        n = self.flat_g.numel()
        clip = float(self.max_norm) if self.max_norm else 0.0
        if clip > 0.0:
            with _T("optimizer", nbytes=4 * n):
                _lib.check(L.flat_sqnorm(ptr(self.flat_g), n, ptr(self._sq), _stream()), "flat_sqnorm")
        with _T("optimizer", nbytes=22 * n):
            rc = L.sgd_nesterov_step(ptr(self.flat_p), ptr(self.flat_m), ptr(self.flat_g), ptr(self.flat_p16), n,
                                     ptr(self._sq) if clip > 0.0 else None, ptr(self.grad_norm), 1.0 / world, clip,
                                     float(d["lr"]), float(d["momentum"]), float(d["weight_decay"]), int(bool(d["nesterov"])),
                                     _stream())
        _lib.check(rc, "sgd_nesterov_step")
        self._transpose()                                                      # W^T / rotated operands of the next backward
        ops.bump_weight_epoch()                                                # derived bf16 forms (W^T, KRSC, ...) are stale
        ops.flush_counters()                                                   # the step's BatchNorm counters, one launch
        self.steps += 1

    # ---- torch.optim.SGD-compatible checkpoint format (utils/checkpoint.py:235-265 stores optimizer.state_dict())
    @property
    def param_groups(self):
        return [dict(self.defaults, params=self.params)]

    def state_dict(self):
        state, off = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            if self.steps:
                state[i] = {"momentum_buffer": self.flat_m[off:off + k].as_strided(p.shape, _dense_strides(p)).clone()}
            off += k
        return {"state": state, "param_groups": [dict(self.defaults, params=list(range(len(self.params))))]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        for k in ("lr", "momentum", "weight_decay", "nesterov"):
            self.defaults[k] = g[k]
        off = 0
        with torch.no_grad():
            for i, p in enumerate(self.params):
                k = p.numel()
                st = sd["state"].get(i) or sd["state"].get(str(i))
                if st is not None and st.get("momentum_buffer") is not None:
                    self.flat_m[off:off + k].as_strided(p.shape, _dense_strides(p)).copy_(st["momentum_buffer"])
                    self.steps = max(self.steps, 1)
                off += k


def _dense_strides(p):
    """The parameter's own strides when they describe a dense block (contiguous or channels_last), else contiguous."""
    if p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last)):
        return p.stride()
    return torch.empty(p.shape).stride()
