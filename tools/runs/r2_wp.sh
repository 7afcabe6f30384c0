#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops + model tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
echo "== wgrad timing (pair vs single)"; timeout 300 python - <<'PY' 2>&1 | tail -12
import torch, ctypes as C
from mvfnet_b200 import ops, _lib
from mvfnet_b200._lib import ptr
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, cout, cin) in [(250880, 256, 1024), (250880, 1024, 256), (62720, 512, 2048), (62720, 2048, 512), (1003520, 512, 256)]:
    g = torch.randn(M, cout, device="cuda").bfloat16(); x = torch.randn(M, cin, device="cuda").bfloat16()
    a = t(lambda: ops.gemm_wgrad(g, x)); ya = ops.gemm_wgrad(g, x).clone()
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
    b = t(lambda: ops.gemm_wgrad(g, x)); yb = ops.gemm_wgrad(g, x).clone()
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
    print("wgrad1x1 M=%d %dx%d: pair %.0f us  single %.0f us  rel %.1e" % (M, cout, cin, a, b, ((ya - yb).norm() / yb.norm()).item()))
for (hw, c) in [(14, 256), (7, 512)]:
    F = 1280
    x = torch.randn(F, c, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    g = torch.randn(F, c, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    a = t(lambda: ops.conv3x3_wgrad_raw(g, x, c, 1, 3)); ya = ops.conv3x3_wgrad_raw(g, x, c, 1, 3).clone()
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
    b = t(lambda: ops.conv3x3_wgrad_raw(g, x, c, 1, 3)); yb = ops.conv3x3_wgrad_raw(g, x, c, 1, 3).clone()
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
    print("wgrad3x3 C=%d %dx%d: pair %.0f us  single %.0f us  rel %.1e" % (c, hw, hw, a, b, ((ya - yb).norm() / yb.norm()).item()))
PY
