#!/bin/bash
# 4-GPU check of the bench contract under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/bench_4gpu_r2e.json 2> gpurun_out/bench_4gpu_r2e.err
echo rc $?; grep -v "^\*\|OMP" gpurun_out/bench_4gpu_r2e.err | tail -3; cut -c1-400 gpurun_out/bench_4gpu_r2e.json
