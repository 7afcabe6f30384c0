#!/bin/bash
# 2-GPU validation of the NCCL path (FlatSGD all-reduce, gpu_bar under torchrun, other configs)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/h_gpus.txt
echo "== 2-GPU bench"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/h_bench_2gpu.json 2> gpurun_out/h_bench_2gpu.err; tail -c 2500 gpurun_out/h_bench_2gpu.json; tail -8 gpurun_out/h_bench_2gpu.err
echo "== 2-GPU reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/h_ref_2gpu.json 2> gpurun_out/h_ref_2gpu.err; tail -c 600 gpurun_out/h_ref_2gpu.json; tail -3 gpurun_out/h_ref_2gpu.err
