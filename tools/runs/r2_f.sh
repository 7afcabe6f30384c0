#!/bin/bash
# ncu captures of the MVF sweep kernels (forward train-mode + backward) at 160 clips, and the step's launch list
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ncu mvf sweep fwd/bwd"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:mvf_sweep -s 6 -c 4 -o gpurun_out/f_mvf_sweep python tools/mvf_microbench.py --clips 160 --iters 2 --only-C 1024 --out gpurun_out/f_tmp.jsonl > gpurun_out/f_ncu_mvf.log 2>&1; tail -3 gpurun_out/f_ncu_mvf.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 1 --batch 160 --kernels-only > gpurun_out/f_ncu_list.log 2>&1; tail -2 gpurun_out/f_ncu_list.log
ls -la gpurun_out/f_*
