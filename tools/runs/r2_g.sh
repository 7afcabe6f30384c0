#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log; tail -22 gpurun_out/g_pytest.log | cut -c1-300
echo "== microbench bwd"; timeout 600 python tools/mvf_microbench.py --clips 160 --iters 10 --out gpurun_out/g_mvf_micro.jsonl > gpurun_out/g_mvf_micro.log 2>&1; cut -c1-230 gpurun_out/g_mvf_micro.log
echo "== bench B=160"; timeout 900 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 1 --no-gpu-bar --no-other-configs --sweep "12" > gpurun_out/g_bench_b160.json 2> gpurun_out/g_bench_b160.err; tail -c 700 gpurun_out/g_bench_b160.json; tail -5 gpurun_out/g_bench_b160.err
