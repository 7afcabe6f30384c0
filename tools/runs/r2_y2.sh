#!/bin/bash
# 2-GPU check of the final build, launched as the driver launches it (short: no GPU bar, no other configs)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== 2-GPU bench"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 12 --warmup 4 --no-gpu-bar --no-other-configs > gpurun_out/y2_bench.json 2> gpurun_out/y2_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/y2_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "e2e", "sanity", "clocks", "cuda_graph_step")})
PY
tail -4 gpurun_out/y2_bench.err
echo "== 2-GPU reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-400
