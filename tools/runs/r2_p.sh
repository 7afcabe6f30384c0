#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops tests"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
echo "== by shape"; timeout 500 python tools/step_by_shape.py --families gemm1x1,maxpool,stem_im2col,conv3x3 --out gpurun_out/p_by_shape.json 2>&1 | tail -40
echo "== bench"; timeout 1200 python bench.py --no-gpu-bar --no-other-configs --sweep "" --no-graph > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/p_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches") if k in d})
for k, v in d["roofline_by_family"].items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print(d["roofline"]["isolated"])
PY
tail -5 gpurun_out/p_bench.err
