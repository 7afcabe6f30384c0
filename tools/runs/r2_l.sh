#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== s2 dgrad timing"; timeout 300 python - <<'PY' 2>&1 | tail -20
import torch, os
from mvfnet_b200 import ops
torch.manual_seed(0)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for F, C, H in [(1280, 128, 56), (1280, 256, 28), (1280, 512, 14), (96, 128, 56), (96, 512, 14)]:
    w = torch.randn(C, C, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_()
    g = torch.randn(F, C, H // 2, H // 2, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x = torch.randn(F, C, H, H, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wb = w.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ours = t(lambda: ops.conv3x3s2_dgrad_raw(g, w, H, H))
    lib = t(lambda: torch.ops.aten.convolution_backward(g, x, wb, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]))
    a = ops.conv3x3s2_dgrad_raw(g, w, H, H).float()
    b = torch.ops.aten.convolution_backward(g, x, wb, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0].float()
    print(F, C, H, "ours %.0f us  cudnn %.0f us  rel %.2e" % (ours, lib, ((a - b).norm() / b.norm()).item()))
PY
echo "== ops tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "== ncu sweep fwd (C=1024 14x14, 160 clips)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mvf_sweep_kernel -c 8 -f -o gpurun_out/l_sweep_fwd python tools/mvf_microbench.py --iters 1 --clips 160 --only-C 1024 --out gpurun_out/l_micro_ncu.jsonl > gpurun_out/l_ncu.log 2>&1; tail -3 gpurun_out/l_ncu.log
ls -la gpurun_out/l_sweep_fwd.ncu-rep
