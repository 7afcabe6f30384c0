#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/d_pytest.log; tail -25 gpurun_out/d_pytest.log
echo "== bench B=160 full"; timeout 1200 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 2 > gpurun_out/d_bench_b160.json 2> gpurun_out/d_bench_b160.err; tail -c 3500 gpurun_out/d_bench_b160.json; tail -5 gpurun_out/d_bench_b160.err
echo "== bench ADDFUSE"; MVFB_ADDFUSE=1 timeout 600 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 1 --no-gpu-bar --no-other-configs --sweep "" > gpurun_out/d_bench_addfuse.json 2> gpurun_out/d_bench_addfuse.err; tail -c 500 gpurun_out/d_bench_addfuse.json; tail -3 gpurun_out/d_bench_addfuse.err
