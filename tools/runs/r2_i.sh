#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest.log; tail -22 gpurun_out/i_pytest.log | cut -c1-400
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; tail -c 4500 gpurun_out/i_bench.json; tail -5 gpurun_out/i_bench.err
