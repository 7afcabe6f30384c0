#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"maxpool|stem_im2col" -c 3 -f -o gpurun_out/q_stem python tools/step_by_shape.py --batch 160 --steps 1 --families maxpool > gpurun_out/q_ncu.log 2>&1; tail -3 gpurun_out/q_ncu.log
ls -la gpurun_out/q_stem.ncu-rep
