#!/bin/bash
# ablations of the halo kernel: which stage bounds a tile (variants built with tools/build_variant.py halo_X -DHALO_X)
set -u
cd "$(dirname "$0")/../.."
for v in ${VARIANTS:-default NO_MMA NO_STORE NO_CONVERT NO_CS}; do
MVFB_VARIANT=$v timeout 200 python - <<'PY' 2>&1 | tail -3
import os, torch
from mvfnet_b200 import _lib
v = os.environ["MVFB_VARIANT"]
if v != "default": _lib.LIB_PATH = os.path.abspath("variants/halo_%s/libmvf_b200.so" % v)
from mvfnet_b200 import ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out = []
for F, Cin, Cout, H in [(1280, 64, 64, 56), (1280, 128, 128, 28)]:
    x = torch.randn(F, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)).bfloat16().permute(0, 2, 3, 1).contiguous()
    out.append("C=%d: %.0f / %.0f us" % (Cin, t(lambda: ops.conv3x3_raw(x, w, 1, stats=False)), t(lambda: ops.conv3x3_raw(x, w, 1, stats=True))))
print("%-12s %s" % (v, "   ".join(out)))
PY
done
