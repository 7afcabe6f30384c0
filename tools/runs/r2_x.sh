#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops + model tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
echo "== by shape"; timeout 500 python tools/step_by_shape.py --families conv3x3 --out gpurun_out/x_by_shape.json 2>&1 | tail -12
