#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops tests"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "maxpool or stem" 2>&1 | tail -5 | cut -c1-300
echo "== by shape"; timeout 500 python tools/step_by_shape.py --families maxpool,stem_im2col --out gpurun_out/r_by_shape.json 2>&1 | tail -5
timeout 300 python - <<'PY'
import torch
from mvfnet_b200 import ops
x = torch.randn(1280, 64, 112, 112, device="cuda").relu_().bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
y = ops.maxpool3x3s2(x)
g = torch.randn_like(y)
print("maxpool fwd %.0f us" % t(lambda: ops.maxpool3x3s2(x.detach())))
print("maxpool fwd+bwd %.0f us" % t(lambda: ops.maxpool3x3s2(x).backward(g)))
pool = torch.nn.MaxPool2d(3, 2, 1)
print("aten fwd %.0f us" % t(lambda: pool(x.detach())))
print("aten fwd+bwd %.0f us" % t(lambda: pool(x).backward(g)))
PY
