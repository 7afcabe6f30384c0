#!/usr/bin/env python
"""In-kernel timeline of the MVF stream forward kernel (MVFB_STREAM_DEBUG=4 makes every CTA write %globaltimer
stamps: start, prologue done, first frame landed, main loop done) next to the CUDA-event duration."""
import os, sys, ctypes as C
os.environ["MVFB_STREAM_DEBUG"] = os.environ.get("MVFB_STREAM_DEBUG", "4")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import MVF, _lib
from mvfnet_b200.mvf import MvfDesc, _make_desc, _lib as L_, ptr, _stream

T, Cc, H, Cs = 8, 1024, 14, 128
for B in (64, 128):
    m = MVF(torch.nn.Identity(), T, Cc, alpha=0.125).cuda().eval()
    x = torch.randn(B * T, Cc, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    cfg = m._cfg()
    wt, wh, ww = (w.detach().float().contiguous() for w in m._taps())
    d = _make_desc(x, 1, cfg)
    lib = _lib.lib()
    y = torch.empty((B * T, H, H, Cs), dtype=torch.bfloat16, device="cuda")
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for it in range(4):
        flush.zero_(); flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.mvf_fwd(C.byref(d), ptr(x), ptr(y), Cs, ptr(wt), ptr(wh), ptr(ww), ptr(m.bn.weight), ptr(m.bn.bias),
                         ptr(m.bn.running_mean), ptr(m.bn.running_var), None, None, ptr(ws), ws.numel(), _stream())
        e1.record(); torch.cuda.synchronize()
        assert rc == 0, lib.mvf_b200_last_error()
    st = ws.view(torch.int64)[: 144 * 4].view(144, 4).cpu()
    t0 = st[:, 0].min()
    rel = (st - t0).float() / 1e3
    print("B=%d event %.1f us | CTA start min/max %.1f/%.1f | prologue done max %.1f | first frame max %.1f | loop done min/max %.1f/%.1f"
          % (B, e0.elapsed_time(e1) * 1e3, rel[:, 0].min(), rel[:, 0].max(), rel[:, 1].max(), rel[:, 2].max(), rel[:, 3].min(), rel[:, 3].max()))
