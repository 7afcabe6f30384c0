#!/bin/bash
# CTA-pair mode of the halo kernel: parity tests, then pair vs single timing at the layer2 shape
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== conv tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "conv3x3 or halo" 2>&1 | tail -15 | cut -c1-250
echo "== timing"; timeout 300 python - <<'PY' 2>&1 | tail -12
import torch
from mvfnet_b200 import ops, _lib
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for F, Cin, Cout, H in [(1280, 64, 64, 56), (1280, 128, 128, 28), (1280, 64, 128, 28), (640, 128, 256, 28)]:
    x = torch.randn(F, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)).bfloat16().permute(0, 2, 3, 1).contiguous()
    for stats in (False, True):
        a = t(lambda: ops.conv3x3_raw(x, w, 1, stats=stats))
        _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
        b = t(lambda: ops.conv3x3_raw(x, w, 1, stats=stats))
        _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
        fl = 2 * F * H * H * 9 * Cin * Cout
        print("Cin=%d Cout=%d %dx%d stats=%d: pair %.0f us (%.0f TF/s)   single %.0f us (%.0f TF/s)" % (Cin, Cout, H, H, stats, a, fl / a / 1e6, b, fl / b / 1e6))
PY
