#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
for v in default nomma nostore noload noloadnostore; do
echo "== $v"
VARIANT=$v timeout 300 python - <<'PY' 2>&1 | tail -4
import os, torch
v = os.environ["VARIANT"]
if v != "default":
    from mvfnet_b200 import _lib as L0
    L0.LIB_PATH = os.path.abspath("variants/%s/libmvf_b200.so" % v)
from mvfnet_b200 import ops, _lib
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for F, C, H in [(1280, 64, 56), (1280, 128, 28)]:
    x = torch.randn(F, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16().permute(0, 2, 3, 1).contiguous()
    a = t(lambda: ops.conv3x3_raw(x, w, 1, stats=False))
    print("C=%d %dx%d: halo %.0f us" % (C, H, H, a))
PY
done
