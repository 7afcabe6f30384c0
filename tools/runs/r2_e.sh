#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest mvf + production + ops"; timeout 1500 python -m pytest tests/test_mvf_gpu.py tests/test_mvf_production_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=25 --durations=6 > gpurun_out/e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/e_pytest.log; tail -40 gpurun_out/e_pytest.log | cut -c1-300
echo "== microbench"; timeout 600 python tools/mvf_microbench.py --clips 64,160 --iters 10 --out gpurun_out/e_mvf_micro.jsonl > gpurun_out/e_mvf_micro.log 2>&1; grep mvf_bwd gpurun_out/e_mvf_micro.log | cut -c1-250
echo "== bench B=160"; timeout 900 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 1 --no-gpu-bar --no-other-configs --sweep "" > gpurun_out/e_bench_b160.json 2> gpurun_out/e_bench_b160.err; tail -c 700 gpurun_out/e_bench_b160.json; tail -5 gpurun_out/e_bench_b160.err
