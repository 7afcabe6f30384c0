#!/bin/bash
# fused norm1 + ReLU + max-pool of the stem: parity, then fused vs separate launches at the step's size
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "maxpool or bn_act" 2>&1 | tail -25 | cut -c1-300
for v in ${VARIANTS:-default}; do
echo "== timing $v"; MVFB_VARIANT=$v timeout 300 python - <<'PY' 2>&1 | tail -4
import os, torch
from mvfnet_b200 import _lib
v = os.environ["MVFB_VARIANT"]
if v != "default": _lib.LIB_PATH = os.path.abspath("variants/%s/libmvf_b200.so" % v)
from mvfnet_b200 import ops
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
F, C, H = 1280, 64, 112
x = torch.randn(F, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
bn = torch.nn.BatchNorm2d(C).cuda()
sums = torch.stack([x.float().sum((0, 2, 3)), (x.float() ** 2).sum((0, 2, 3))]).contiguous()
gy = torch.randn(F, C, H // 2, H // 2, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
for name, fn in (("fused", lambda xx: ops.bn_relu_maxpool(xx, bn, sums=sums)),
                 ("separate", lambda xx: ops.maxpool3x3s2(ops.bn_act(xx, bn, relu=True, sums=sums)))):
    if name == "separate" and v != "default": continue
    xr = x.clone().requires_grad_(True)
    fwd = t(lambda: fn(xr))
    def both():
        y = fn(xr)
        y.backward(gy)
        xr.grad = None
    tot = t(both)
    print("%-9s fwd %.0f us   fwd+bwd %.0f us   (bwd %.0f us)" % (name, fwd, tot, tot - fwd))
PY
done
