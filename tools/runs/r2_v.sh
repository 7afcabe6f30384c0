#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ops tests"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "stem" 2>&1 | tail -5 | cut -c1-300
echo "== by shape"; timeout 500 python tools/step_by_shape.py --families maxpool,stem_im2col --out gpurun_out/v_by_shape.json 2>&1 | tail -4
