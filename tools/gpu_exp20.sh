#!/bin/bash
# 2-GPU check of the bench contract under torchrun (env groups + NCCL all-reduce of the info vector)
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_2gpu_r2c.json 2> gpurun_out/bench_2gpu_r2c.err
echo rc $?; tail -3 gpurun_out/bench_2gpu_r2c.err; cut -c1-700 gpurun_out/bench_2gpu_r2c.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 | cut -c1-300
