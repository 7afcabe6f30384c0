#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/u_pytest.log; tail -12 gpurun_out/u_pytest.log | cut -c1-300
echo "== bench"; timeout 1200 python bench.py --no-gpu-bar --no-other-configs > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "sweep", "cuda_graph_step") if k in d})
print(d["roofline_by_family"]["other (ATen / NCCL / gaps, not bracketed)"])
PY
tail -5 gpurun_out/u_bench.err
