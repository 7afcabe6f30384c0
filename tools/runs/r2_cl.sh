#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== cluster tests"; timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "pair or gemm_tn or conv3x3" 2>&1 | tail -8 | cut -c1-300
echo "== timing"; timeout 300 python - <<'PY' 2>&1 | tail -14
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
for M, N, K in [(250880, 1024, 256), (250880, 256, 1024), (62720, 2048, 512), (62720, 512, 2048), (1003520, 512, 256), (1003520, 256, 512)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    for stats in (False, True):
        x = t(lambda: ops.gemm_tn(a, b, stats=stats))
        _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
        y = t(lambda: ops.gemm_tn(a, b, stats=stats))
        _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
        print("M=%d N=%d K=%d stats=%d: pair %.0f us (%.0f TF/s)  single %.0f us (%.0f TF/s)" % (M, N, K, stats, x, 2*M*N*K/x/1e6, y, 2*M*N*K/y/1e6))
for F, C, H in [(1280, 256, 14), (1280, 512, 7)]:
    x = torch.randn(F, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16().permute(0, 2, 3, 1).contiguous()
    p = t(lambda: ops.conv3x3_raw(x, w, 1, stats=True))
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
    q = t(lambda: ops.conv3x3_raw(x, w, 1, stats=True))
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
    fl = 2 * F * H * H * 9 * C * C
    print("conv3x3 C=%d %dx%d: pair %.0f us (%.0f TF/s)  single %.0f us (%.0f TF/s)" % (C, H, H, p, fl/p/1e6, q, fl/q/1e6))
PY
