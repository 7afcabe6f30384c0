#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none -k regex:"conv_halo_kernel|gemm_tn_kernel|gemm_wgrad_kernel" --launch-skip 200 -c 48 -f -o /tmp/ncu2_tc python tools/step_by_shape.py --batch 160 --steps 1 --families conv3x3 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ncu -i /tmp/ncu2_tc.ncu-rep --page raw --csv > gpurun_out/ncu2_raw.csv 2>/dev/null; ls -la gpurun_out/ncu2_raw.csv
