#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops + model tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_tail_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
echo "== by shape"; timeout 500 python tools/step_by_shape.py --families gemm1x1,conv3x3,wgrad1x1,wgrad3x3 --out gpurun_out/el_by_shape.json 2>&1 | tail -60
echo "== bench"; timeout 1200 python bench.py --no-gpu-bar --no-other-configs > gpurun_out/el_bench.json 2> gpurun_out/el_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/el_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "sweep", "cuda_graph_step", "clocks") if k in d})
for k, v in d["roofline_by_family"].items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
PY
tail -3 gpurun_out/el_bench.err
