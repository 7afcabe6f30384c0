#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches.csv python bench.py --steps 1 --warmup 1 --batch 160 --kernels-only > gpurun_out/s_ncu_list.log 2>&1; tail -2 gpurun_out/s_ncu_list.log
python tools/launch_list_summary.py gpurun_out/s_launches.csv | head -70
