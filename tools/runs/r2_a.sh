#!/bin/bash
# round 2, GPU call A: full GPU test suite, headline bench with the GPU bar, batch comparison, sanitizer passes
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
free -g >> gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -x --deselect tests/test_mvf_production_gpu.py > gpurun_out/a_pytest_base.log 2>&1; echo "rc=$?" >> gpurun_out/a_pytest_base.log; tail -5 gpurun_out/a_pytest_base.log
echo "== pytest production"; timeout 1200 python -m pytest tests/test_mvf_production_gpu.py -m gpu -q --maxfail=40 --durations=15 > gpurun_out/a_pytest_prod.log 2>&1; echo "rc=$?" >> gpurun_out/a_pytest_prod.log; tail -25 gpurun_out/a_pytest_prod.log
echo "== bench B=160 with gpu bar"; timeout 900 python bench.py --steps 10 --warmup 4 --batch 160 --cpu-seconds 5 > gpurun_out/a_bench_b160.json 2> gpurun_out/a_bench_b160.err; tail -c 3000 gpurun_out/a_bench_b160.json; tail -5 gpurun_out/a_bench_b160.err
echo "== bench B=148"; timeout 600 python bench.py --steps 10 --warmup 4 --batch 148 --cpu-seconds 1 --no-gpu-bar --sweep "" > gpurun_out/a_bench_b148.json 2> gpurun_out/a_bench_b148.err; tail -c 1500 gpurun_out/a_bench_b148.json
echo "== sanitizer memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mvf_gpu.py -m gpu -q -x -k "golden or variants or errors" > gpurun_out/a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/a_memcheck.log; tail -8 gpurun_out/a_memcheck.log
echo "== sanitizer racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_mvf_gpu.py -m gpu -q -x -k "variants" > gpurun_out/a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/a_racecheck.log; tail -8 gpurun_out/a_racecheck.log
