#!/usr/bin/env python
"""In-kernel timeline of the MVF forward kernel (mvf_b200_set_option(MVFB_OPT_SWEEP_DEBUG, 1) makes every CTA of mvf_sweep_kernel write
%globaltimer stamps: start, prologue done, first frame landed, [train: statistics sweep done, grid barrier passed],
last frame stored) next to the CUDA-event duration of the same launch and of an empty launch."""
import os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import MVF, _lib
from mvfnet_b200.mvf import _make_desc, ptr, _stream

lib = _lib.lib()
_lib.set_option(_lib.OPT_SWEEP_DEBUG, 1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
tiny = torch.zeros(32, device="cuda")
for it in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tiny.add_(1.0); e1.record(); torch.cuda.synchronize()
print("empty launch between events: %.1f us" % (e0.elapsed_time(e1) * 1e3))
shapes = [(8, 1024, 14, 128), (8, 2048, 7, 256), (8, 512, 28, 64)]
for (T, Cc, H, Cs) in shapes:
    for B in (64, 128):
        for training in (False, True):
            m = MVF(torch.nn.Identity(), T, Cc, alpha=0.125).cuda().train(training)
            x = torch.randn(B * T, Cc, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
            cfg = m._cfg()
            wt, wh, ww = (w.detach().float().contiguous() for w in m._taps())
            d = _make_desc(x, 1, cfg)
            y = torch.empty((B * T, H, H, Cs), dtype=torch.bfloat16, device="cuda")
            ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
            sm, sr = torch.empty(Cs, device="cuda"), torch.empty(Cs, device="cuda")
            for it in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.mvf_fwd(C.byref(d), ptr(x), ptr(y), Cs, ptr(wt), ptr(wh), ptr(ww), ptr(m.bn.weight), ptr(m.bn.bias),
                                 ptr(m.bn.running_mean), ptr(m.bn.running_var), ptr(sm), ptr(sr), ptr(ws), ws.numel(), _stream())
                e1.record(); torch.cuda.synchronize()
                assert rc == 0, lib.mvf_b200_last_error()
            st = ws[512 << 10:].view(torch.int64)[: 148 * 8].view(148, 8).cpu()
            st = st[st[:, 0] > 0]
            t0 = st[:, 0].min()
            rel = (st - t0).float() / 1e3
            msg = ("%dx%d Cs=%d B=%d %s: event %.1f us | CTAs %d start max %.1f | prologue done max %.1f | first frame max %.1f"
                   % (H, H, Cs, B, "train" if training else "eval", e0.elapsed_time(e1) * 1e3, len(st), rel[:, 0].max(),
                      rel[:, 1].max(), rel[:, 2].max()))
            if training:
                msg += (" | stats sweep done (warp 0) min/max %.1f/%.1f, all warps max %.1f | rows arrived min/max %.1f/%.1f | exchange done max %.1f"
                        % (rel[:, 3].min(), rel[:, 3].max(), rel[:, 6].max(), rel[:, 7].min(), rel[:, 7].max(), rel[:, 4].max()))
            msg += " | loop done min/max %.1f/%.1f" % (rel[:, 5].min(), rel[:, 5].max())
            print(msg, flush=True)
