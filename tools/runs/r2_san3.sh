#!/bin/bash
# compute-sanitizer over the last kernels of the round: fused norm1 + ReLU + max-pool, CTA pairs in the halo convolution
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== memcheck"; timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "(bn_relu_maxpool and not 112) or conv_halo_cta_pair" > gpurun_out/san3_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/san3_memcheck.log; tail -4 gpurun_out/san3_memcheck.log | cut -c1-200
echo "== racecheck"; timeout 70 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "(bn_relu_maxpool and 8-8) or (conv_halo_cta_pair and 50-64)" > gpurun_out/san3_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/san3_racecheck.log; tail -4 gpurun_out/san3_racecheck.log | cut -c1-200
