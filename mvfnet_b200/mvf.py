"""Multi-View Fusion module -- host-side mirror of codes/models/modules/MVF.py.

Same public surface as the reference (`make_multi_view_fusion`, `MVF(net, n_segment, in_channels, alpha,
use_hs, share, mode).forward(x)`, attribute names and state_dict keys, MVF.py:18-138), but `forward`
does not execute any torch op for the fusion itself: the view/transpose/split, the three depthwise
Conv3d, BatchNorm3d, HardSwish, cat and contiguous of MVF.py:109-137 are ONE call into
libmvf_b200.so (`mvf_fwd`, include/mvf_b200.h), and autograd's backward of all of them is `mvf_bwd`.
The `nn.Conv3d` / `nn.BatchNorm3d` sub-modules exist only to own the parameters under the
reference's names (`shift_conv.weight`, `h_conv.weight`, `w_conv.weight`, `bn.*`); they are never called.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import MvfDesc, ptr
from .common import HardSwish


# Optional launch timing used by bench.py: while a list is installed here every library call of the hot path is
# bracketed by CUDA events recorded on the launching stream, with its algorithmic bytes and flops (MVF
# forward 2*E*s, backward 3*E*s for a slab of E elements of s bytes; GEMMs 2*M*N*K flops and operand + result bytes;
# BatchNorm kernels the bytes they stream).
_TIMING = None


def timing_begin():
    global _TIMING
    _TIMING = []


def timing_end():
    """-> [(kind, algorithmic_bytes, start_event, end_event, flops)]; call after a device synchronize."""
    global _TIMING
    rec, _TIMING = _TIMING, None
    return rec or []


class _Timed:
    def __init__(self, kind, elems=0, esize=0, nbytes=None, flops=0):
        self.kind, self.flops = kind, flops
        self.bytes = nbytes if nbytes is not None else (2 if kind == "mvf_fwd" else 3) * elems * esize

    def __enter__(self):
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _TIMING is not None and exc[0] is None:
            self.e1.record()
            _TIMING.append((self.kind, self.bytes, self.e0, self.e1, self.flops))
        return False


def _layout_of(x):
    """(layout enum, tensor the kernel can address) for a (F, C, H, W) activation.  When C == 1 or
    H*W == 1 both layouts describe the same bytes, so the order of the tests does not matter."""
    if x.is_contiguous():
        return _lib.MVFB_NCHW, x
    if x.is_contiguous(memory_format=torch.channels_last):
        return _lib.MVFB_NHWC, x
    return _lib.MVFB_NCHW, x.contiguous()


def _dtype_of(x):
    if x.dtype == torch.float32:
        return _lib.MVFB_F32
    if x.dtype == torch.bfloat16:
        return _lib.MVFB_BF16
    raise TypeError("MVF kernels take float32 or bfloat16 activations, got %s" % x.dtype)


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


class _Cfg:
    __slots__ = ("T", "Cs", "mode", "share", "use_hs", "training", "eps", "momentum")


def _make_desc(x, layout, cfg):
    f, c, h, w = x.shape
    d = MvfDesc()
    d.N, d.T, d.C, d.Cs, d.H, d.W = f // cfg.T, cfg.T, c, cfg.Cs, h, w
    d.dtype, d.layout = _dtype_of(x), layout
    d.mode, d.use_hs, d.training = _lib.MODES[cfg.mode], int(cfg.use_hs), int(cfg.training)
    d.eps, d.momentum = cfg.eps, cfg.momentum
    return d


def _stream():
    # the raw handle of the current stream of the current device: two C calls (torch.cuda.current_stream() builds a Stream
    # object through several Python layers, ~16 us -- 270 launches per step made that a quarter of the host time at B = 12)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _out_stride(layout, c, h, w):
    return c * h * w if layout == _lib.MVFB_NCHW else c


def mvf_slab_forward(x, cfg, wt, wh, ww, gamma, beta, running_mean, running_var, out=None):
    """Raw forward: returns (y, save_mean, save_rstd).  `out=None`: y is a copy of x with the slab channels
    [0,Cs) replaced (what MVF.py:135-137 materialises); `out='slab'`: a compact (F,Cs,H,W) tensor in
    x's layout (consumed by the K-split 1x1 convolution)."""
    if not x.is_cuda:
        raise RuntimeError("mvfnet_b200.MVF runs only on CUDA tensors (no CPU path exists)")
    L = _lib.lib()
    layout, xk = _layout_of(x)
    f, c, h, w = xk.shape
    if f % cfg.T != 0:
        raise ValueError("batch of %d frames is not a multiple of n_segment=%d" % (f, cfg.T))
    d = _make_desc(xk, layout, cfg)
    if out == "slab":
        if layout == _lib.MVFB_NHWC:
            y = torch.empty((f, h, w, cfg.Cs), dtype=xk.dtype, device=xk.device).permute(0, 3, 1, 2)
        else:
            y = torch.empty((f, cfg.Cs, h, w), dtype=xk.dtype, device=xk.device)
        ystride = _out_stride(layout, cfg.Cs, h, w)
    else:
        y = xk.clone(memory_format=torch.preserve_format)
        ystride = _out_stride(layout, c, h, w)
    save_mean = save_rstd = None
    if cfg.use_hs:
        save_mean = torch.empty(cfg.Cs, dtype=torch.float32, device=x.device)
        save_rstd = torch.empty(cfg.Cs, dtype=torch.float32, device=x.device)
    nbytes = L.mvf_fwd_workspace_bytes(C.byref(d))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    with _Timed("mvf_fwd", f * cfg.Cs * h * w, xk.element_size()):
        rc = L.mvf_fwd(C.byref(d), ptr(xk), ptr(y), ystride, ptr(wt), ptr(wh), ptr(ww), ptr(gamma), ptr(beta),
                       ptr(running_mean), ptr(running_var), ptr(save_mean), ptr(save_rstd), ptr(ws), ws.numel(),
                       _stream())
    _lib.check(rc, "mvf_fwd")
    return y, xk, layout, save_mean, save_rstd


class _MVFFunction(torch.autograd.Function):
    """x -> x' (slab channels fused, the rest passed through bit-exactly).  MVF.py:109-137."""

    @staticmethod
    def forward(ctx, x, wt, wh, ww, gamma, beta, running_mean, running_var, cfg):
        y, xk, layout, save_mean, save_rstd = mvf_slab_forward(
            x, cfg, wt, wh, ww, gamma, beta, running_mean, running_var)
        ctx.cfg, ctx.layout = cfg, layout
        ctx.save_for_backward(xk, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd)
        return y

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        cfg, layout = ctx.cfg, ctx.layout
        xk, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd = ctx.saved_tensors
        f, c, h, w = xk.shape
        mf = torch.channels_last if layout == _lib.MVFB_NHWC else torch.contiguous_format
        # dL/dx' of the pass-through channels IS g: start from a private copy and let the kernel overwrite
        # the slab channels (the reference gets the same result from cat/split backward, MVF.py:110,135)
        gk = g.to(xk.dtype).contiguous(memory_format=mf)
        dx = gk.clone(memory_format=mf)
        d = _make_desc(xk, layout, cfg)
        stride = _out_stride(layout, c, h, w)
        dev = xk.device
        dwt = torch.empty((cfg.Cs, 3), dtype=torch.float32, device=dev)
        dwh = torch.empty_like(dwt) if (wh is not None and wh.data_ptr() != wt.data_ptr()) else None
        dww = torch.empty_like(dwt) if (ww is not None and ww.data_ptr() != wt.data_ptr()) else None
        dgamma = torch.empty(cfg.Cs, dtype=torch.float32, device=dev) if cfg.use_hs else None
        dbeta = torch.empty_like(dgamma) if cfg.use_hs else None
        nbytes = L.mvf_bwd_workspace_bytes(C.byref(d))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        with _Timed("mvf_bwd", f * cfg.Cs * h * w, xk.element_size()):
            rc = L.mvf_bwd(C.byref(d), ptr(gk), stride, ptr(xk), ptr(dx), stride, ptr(wt), ptr(wh), ptr(ww),
                           ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(save_mean),
                           ptr(save_rstd), ptr(dwt), ptr(dwh), ptr(dww), ptr(dgamma), ptr(dbeta), ptr(ws),
                           ws.numel(), _stream())
        _lib.check(rc, "mvf_bwd")
        return dx, dwt, dwh, dww, dgamma, dbeta, None, None, None


class MVF(nn.Module):
    """MVF Module (MVF.py:53-138): same constructor, attributes and state_dict keys as the reference."""

    def __init__(self, net, n_segment, in_channels, alpha=0.5, use_hs=True, share=False, mode='THW'):
        super().__init__()
        self.net = net
        self.n_segment = n_segment
        num_shift_channel = int(in_channels * alpha)                      # MVF.py:59
        self.num_shift_channel = num_shift_channel
        self.share = share
        if num_shift_channel != 0:
            cs = num_shift_channel
            self.split_sizes = [cs, in_channels - cs]
            # parameter holders under the reference's names / shapes (MVF.py:65-87); never called
            self.shift_conv = nn.Conv3d(cs, cs, [3, 1, 1], stride=1, padding=[1, 0, 0], groups=cs, bias=False)
            self.bn = nn.BatchNorm3d(cs)
            self.use_hs = use_hs
            # same attribute as the reference (MVF.py:71); the kernels apply it, the module is never called
            self.activation = HardSwish() if use_hs else nn.ReLU(inplace=True)
            self.mode = mode
            if mode not in ('THW', 'T', 'TH'):
                raise ValueError("mode must be one of 'THW', 'T', 'TH', got %r" % (mode,))
            if not share:
                if mode in ('THW', 'TH'):
                    self.h_conv = nn.Conv3d(cs, cs, [1, 3, 1], stride=1, padding=[0, 1, 0], groups=cs, bias=False)
                if mode == 'THW':
                    self.w_conv = nn.Conv3d(cs, cs, [1, 1, 3], stride=1, padding=[0, 0, 1], groups=cs, bias=False)
            self._initialize_weights()

    def _initialize_weights(self):
        """N(0, sqrt(2 / (3*Cs))) taps, BN gamma=1 beta=0 (MVF.py:91-102)."""
        for m in (getattr(self, n, None) for n in ('shift_conv', 'h_conv', 'w_conv')):
            if m is not None:
                n = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
        self.bn.weight.data.fill_(1)
        self.bn.bias.data.zero_()

    def _taps(self):
        cs = self.num_shift_channel
        wt = self.shift_conv.weight.view(cs, 3)
        wh = ww = None
        if self.mode in ('THW', 'TH'):
            wh = wt if self.share else self.h_conv.weight.view(cs, 3)
        if self.mode == 'THW':
            ww = wt if self.share else self.w_conv.weight.view(cs, 3)
        return wt, wh, ww

    def _cfg(self):
        bn = self.bn
        cfg = _Cfg()
        cfg.T, cfg.Cs, cfg.mode, cfg.share, cfg.use_hs = self.n_segment, self.num_shift_channel, self.mode, self.share, bool(self.use_hs)
        cfg.training = bool(bn.training or not bn.track_running_stats)
        cfg.eps = float(bn.eps)
        if bn.momentum is None:                                           # cumulative moving average
            cfg.momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)
        else:
            cfg.momentum = float(bn.momentum)
        return cfg

    def _kernel_args(self):
        """(cfg, taps, BN tensors) in the form the C ABI takes them (fp32, share expressed by aliasing)."""
        cfg = self._cfg()
        wt, wh, ww = self._taps()
        bn = self.bn
        gamma = beta = rm = rv = None
        if cfg.use_hs:
            gamma, beta = bn.weight, bn.bias
            if bn.track_running_stats:
                rm, rv = bn.running_mean, bn.running_var
        wt32, gamma32, beta32 = (t if t is None or t.dtype == torch.float32 else t.float() for t in (wt, gamma, beta))
        wh32 = wt32 if wh is wt else (wh if wh is None or wh.dtype == torch.float32 else wh.float())
        ww32 = wt32 if ww is wt else (ww if ww is None or ww.dtype == torch.float32 else ww.float())
        return cfg, wt32, wh32, ww32, gamma32, beta32, rm, rv

    def _count_batch(self, cfg):
        if cfg.use_hs and cfg.training and self.bn.track_running_stats:
            from . import ops
            ops.count_batch(self.bn)

    def fuse(self, x):
        """x' of MVF.py:137 (the tensor handed to self.net)."""
        if self.num_shift_channel == 0:                                   # MVF.py:108
            return x
        cfg, wt, wh, ww, gamma, beta, rm, rv = self._kernel_args()
        y = _MVFFunction.apply(x, wt, wh, ww, gamma, beta, rm, rv, cfg)
        self._count_batch(cfg)
        return y

    def forward(self, x, with_stats=False):
        """MVF.forward(x) of the reference.  `with_stats=True` (used by Bottleneck's fused path only) also returns
        the (2, Cout) per-channel sums of the output for the BatchNorm that follows, or None when the output was not
        produced by the GEMM path."""
        out = self._forward(x, with_stats)
        if with_stats:
            return out if isinstance(out, tuple) else (out, None)
        return out

    def forward_with_identity(self, x):
        """Bottleneck's fused configuration: (out, sums, identity) with `identity` an alias of x whose gradient is summed
        inside this module's backward kernels; None when the fused GEMM path does not apply."""
        from . import ops
        net = self.net
        if (self.num_shift_channel != 0 and self.num_shift_channel % 64 == 0 and isinstance(net, nn.Conv2d)
                and net.kernel_size == (1, 1) and net.stride == (1, 1) and net.bias is None and net.groups == 1
                and ops.eligible(x, net.in_channels, net.out_channels)):
            cfg, wt, wh, ww, gamma, beta, rm, rv = self._kernel_args()
            out = ops.mvf_conv1x1(x, net.weight, wt, wh, ww, gamma, beta, rm, rv, cfg, stats=True, passthrough=True)
            self._count_batch(cfg)
            return out
        return None

    def _forward(self, x, with_stats):
        net = self.net
        if (self.num_shift_channel != 0 and self.num_shift_channel % 64 == 0 and isinstance(net, nn.Conv2d)
                and net.kernel_size == (1, 1) and net.stride == (1, 1) and net.bias is None and net.groups == 1):
            from . import ops
            if ops.eligible(x, net.in_channels, net.out_channels):
                # MVF kernel -> compact slab; tcgen05 GEMM with a K-split A operand (slab | untouched channels of x)
                cfg, wt, wh, ww, gamma, beta, rm, rv = self._kernel_args()
                out = ops.mvf_conv1x1(x, net.weight, wt, wh, ww, gamma, beta, rm, rv, cfg, stats=with_stats)
                self._count_batch(cfg)
                return out
        return net(self.fuse(x))


def make_multi_view_fusion(net, n_segment, alpha, mvf_freq=(1, 1, 1, 1), use_hs=True, share=False, mode='THW'):
    """Wrap `conv1` of every block of the selected ResNet stages in an MVF module (MVF.py:18-49)."""
    n_segment_list = [n_segment] * 4
    assert n_segment_list[-1] > 0
    n_round = 1                                                           # MVF.py:26-30 (fixed 1)

    def make_block_mvf(stage, this_segment):
        blocks = list(stage.children())
        for i, b in enumerate(blocks):
            if i % n_round == 0:
                b.conv1 = MVF(b.conv1, this_segment, b.conv1.in_channels, alpha, use_hs, share, mode)
        return nn.Sequential(*blocks)

    for i, name in enumerate(('layer1', 'layer2', 'layer3', 'layer4')):
        if mvf_freq[i]:
            setattr(net, name, make_block_mvf(getattr(net, name), n_segment_list[i]))
