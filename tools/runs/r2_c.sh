#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== debug T=16"; timeout 600 python tools/runs/debug_t16.py > gpurun_out/c_debug_t16.log 2>&1; tail -60 gpurun_out/c_debug_t16.log
echo "== pytest subset"; timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_tail_gpu.py tests/test_ops_gpu.py tests/test_mvf_production_gpu.py -m gpu -q --maxfail=40 > gpurun_out/c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest.log; tail -30 gpurun_out/c_pytest.log
echo "== gemm probe"; timeout 900 python tools/gemm_probe.py --out gpurun_out/c_gemm_probe.jsonl > gpurun_out/c_gemm_probe.log 2>&1; tail -70 gpurun_out/c_gemm_probe.log
echo "== bench B=160"; timeout 900 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 1 --no-gpu-bar --sweep "" > gpurun_out/c_bench_b160.json 2> gpurun_out/c_bench_b160.err; tail -c 600 gpurun_out/c_bench_b160.json; tail -5 gpurun_out/c_bench_b160.err
