#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch list + one full capture of the dominant kernel.
set -e
mkdir -p gpurun_out
ARGS="--steps 4 --warmup 3 --cpu-baseline-seconds 0.5 --e2e-steps 3 ${BENCH_EXTRA}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG:-r1}.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 5 -c 1 -o gpurun_out/prof_${TAG:-r1} python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-400
ls -la gpurun_out
