#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops + model + tail tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_tail_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
echo "== bench"; timeout 1200 python bench.py --no-gpu-bar --no-other-configs > gpurun_out/mat_bench.json 2> gpurun_out/mat_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/mat_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "sweep", "cuda_graph_step", "clocks", "sanity") if k in d})
PY
tail -3 gpurun_out/mat_bench.err
