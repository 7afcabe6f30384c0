#!/bin/bash
# 8-GPU validation exactly as the driver launches it
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/y_gpus.txt
echo "== 8-GPU bench"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/y_bench_8gpu.json 2> gpurun_out/y_bench_8gpu.err; echo "rc=$?"; tail -c 3000 gpurun_out/y_bench_8gpu.json; tail -6 gpurun_out/y_bench_8gpu.err
