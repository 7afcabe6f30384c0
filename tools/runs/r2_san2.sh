#!/bin/bash
# compute-sanitizer over the halo-band convolution and the GEMM kernels after the elected-issue change
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
K='halo or conv3x3_implicit or gemm_tn or gemm_add or wgrad or strided'
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "$K" > gpurun_out/san2_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/san2_memcheck.log; tail -5 gpurun_out/san2_memcheck.log | cut -c1-200
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "halo or (conv3x3_implicit and 40-64)" > gpurun_out/san2_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/san2_racecheck.log; tail -5 gpurun_out/san2_racecheck.log | cut -c1-200
