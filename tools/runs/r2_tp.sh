#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops + model tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_tail_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
echo "== timing"; timeout 300 python - <<'PY' 2>&1 | tail -8
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
for M, N, K in [(250880, 1024, 256), (250880, 256, 1024), (4014080, 256, 64), (4014080, 64, 256), (1003520, 512, 128), (1003520, 128, 512)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    x = t(lambda: ops.gemm_tn(a, b, stats=False)); y = t(lambda: ops.gemm_tn(a, b, stats=True))
    print("M=%d N=%d K=%d: %.0f us, with stats %.0f us" % (M, N, K, x, y))
PY
echo "== bench"; timeout 1200 python bench.py --no-gpu-bar --no-other-configs --no-graph --sweep "" > gpurun_out/tp_bench.json 2> gpurun_out/tp_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/tp_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "clocks", "sanity") if k in d})
for k, v in d["roofline_by_family"].items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
PY
tail -3 gpurun_out/tp_bench.err
