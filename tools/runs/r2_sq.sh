#!/bin/bash
# full GPU suite + short bench after the fused stem tail
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/sq_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/sq_pytest.log; tail -12 gpurun_out/sq_pytest.log | cut -c1-300
echo "== bench"; timeout 1200 python bench.py --no-gpu-bar --no-other-configs > gpurun_out/sq_bench.json 2> gpurun_out/sq_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/sq_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "sweep", "cuda_graph_step", "clocks", "sanity", "gpu_launches") if k in d})
print(d["config"].get("peak_hbm_gb"))
print({k: (round(v["ms_per_step"], 2), round(v.get("frac_hbm") or 0, 3), round(v.get("frac_tensor") or 0, 3)) for k, v in d["roofline_by_family"].items()})
PY
tail -3 gpurun_out/sq_bench.err
