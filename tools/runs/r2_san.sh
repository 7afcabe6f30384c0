#!/bin/bash
# compute-sanitizer over the kernels added / rewritten late in round 2 (scatter epilogues, TMA addend, pooling, stem)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
K='maxpool or stem or strided or add_epilogue or add_cols or bnact or (conv3x3_implicit and not 100-128)'
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "$K" > gpurun_out/san_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/san_memcheck.log; tail -6 gpurun_out/san_memcheck.log | cut -c1-200
echo "== racecheck (pooling, stem, scatter epilogues)"; timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "maxpool or stem_conv or strided" > gpurun_out/san_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/san_racecheck.log; tail -6 gpurun_out/san_racecheck.log | cut -c1-200
