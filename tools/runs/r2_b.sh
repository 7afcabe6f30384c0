#!/bin/bash
# round 2, GPU call B: full GPU test suite, headline bench (8-warp GEMM epilogue, aligned stem im2col, fused tail, u8 input)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --durations=12 > gpurun_out/b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest.log; tail -40 gpurun_out/b_pytest.log
echo "== bench B=160"; timeout 900 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 2 > gpurun_out/b_bench_b160.json 2> gpurun_out/b_bench_b160.err; tail -c 2500 gpurun_out/b_bench_b160.json; tail -5 gpurun_out/b_bench_b160.err
echo "== bench f32 input"; timeout 600 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 1 --no-gpu-bar --sweep "" --input f32 > gpurun_out/b_bench_f32.json 2> gpurun_out/b_bench_f32.err; tail -c 600 gpurun_out/b_bench_f32.json; tail -3 gpurun_out/b_bench_f32.err
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 1 --batch 160 --kernels-only > gpurun_out/b_ncu.log 2>&1; tail -2 gpurun_out/b_ncu.log
