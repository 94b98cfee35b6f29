#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="envs_32768:16 custom:pokelike,94720,12,4,60,0:32 custom:pokelike,16384,20,5,100,0:8" tools/gpu_variants.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 112 -c 1 -o gpurun_out/prof_r2f_32k python bench.py --only-leg envs_32768 > gpurun_out/ncu_r2f.log 2>&1
GBENV_LANES=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 66 -c 1 -o gpurun_out/prof_r2f_div32 python bench.py --only-leg divergent_32768 > gpurun_out/ncu_r2f_div.log 2>&1
GBENV_LANES=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 112 -c 1 -o gpurun_out/prof_r2f_conv32 python bench.py --only-leg envs_32768 > gpurun_out/ncu_r2f_conv.log 2>&1
tail -1 gpurun_out/ncu_r2f_conv.log
