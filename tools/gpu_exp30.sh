#!/bin/bash
# the headline batch as ONE group (all 4,096 envs resident in one launch of k_run_frames_1): the all-resident issue picture
cd "$(dirname "$0")/.."
ARGS="--groups 1 --steps 4 --warmup 3 --preroll 60 --cpu-baseline-seconds 0.5 --e2e-steps 3 --legs none --also-envs 0"
timeout 300 python bench.py $ARGS > gpurun_out/plain_g1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 64 -c 1 -o gpurun_out/prof_r2e_g1 python bench.py $ARGS > gpurun_out/ncu_full_g1.log 2>&1
ls -la gpurun_out/prof_r2e_g1.ncu-rep
