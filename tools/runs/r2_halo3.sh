#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
timeout 300 python - <<'PY' 2>&1 | tail -12
import torch
from mvfnet_b200 import ops, _lib
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for F, C, H in [(1280, 256, 14), (1280, 128, 28), (1280, 64, 56), (96, 64, 56), (96, 128, 28)]:
    x = torch.randn(F, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16().permute(0, 2, 3, 1).contiguous()
    _lib.set_option(_lib.OPT_SWEEP_DEBUG, 7)
    a = t(lambda: ops.conv3x3_raw(x, w, 1, stats=True))
    ya = ops.conv3x3_raw(x, w, 1, stats=True)[0].float()
    _lib.set_option(_lib.OPT_SWEEP_DEBUG, 0)
    _lib.set_option(_lib.OPT_CONV_HALO_OFF, 1)
    b = t(lambda: ops.conv3x3_raw(x, w, 1, stats=True))
    yb = ops.conv3x3_raw(x, w, 1, stats=True)[0].float()
    _lib.set_option(_lib.OPT_CONV_HALO_OFF, 0)
    fl = 2 * F * H * H * 9 * C * C
    print("F=%d C=%d %dx%d: halo %.0f us (%.0f TF/s)   im2col %.0f us (%.0f TF/s)  rel %.1e" % (F, C, H, H, a, fl / a / 1e6, b, fl / b / 1e6, ((ya - yb).norm() / yb.norm()).item()))
PY
