#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
for v in default gnomma gnostore gnoload gnoloadstore gnoall; do
echo "== $v"
VARIANT=$v timeout 300 python - <<'PY' 2>&1 | tail -8
import os, torch
v = os.environ["VARIANT"]
if v != "default":
    from mvfnet_b200 import _lib as L0
    L0.LIB_PATH = os.path.abspath("variants/%s/libmvf_b200.so" % v)
from mvfnet_b200 import ops, _lib
_lib.set_option(_lib.OPT_GEMM_CLUSTER_OFF, 1)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, N, K in [(250880, 1024, 256), (250880, 256, 1024), (4014080, 256, 64), (4014080, 64, 256), (1003520, 512, 128)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    x = t(lambda: ops.gemm_tn(a, b, stats=False)); y = t(lambda: ops.gemm_tn(a, b, stats=True))
    print("M=%d N=%d K=%d: %.0f us, with stats %.0f us" % (M, N, K, x, y))
PY
done
