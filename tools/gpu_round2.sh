#!/bin/bash
# Round-2 evidence run (under gpurun, one GPU): GPU tests, default bench line, reference arm, ncu launch list, one
# --set full capture of the emulation kernel at the headline batch, one at 32,768 envs and one on the divergence stress.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
TAG=${TAG:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo bench rc $?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ARGS="--steps 4 --warmup 3 --preroll 60 --cpu-baseline-seconds 0.5 --e2e-steps 3 --legs none --also-envs 0"
timeout 300 python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 64 -c 1 -o gpurun_out/prof_${TAG} python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 110 -c 1 -o gpurun_out/prof_r2_envs_32768 python bench.py --only-leg envs_32768 > gpurun_out/ncu_full32k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 70 -c 1 -o gpurun_out/prof_r2_divergent_32768 python bench.py --only-leg divergent_32768 > gpurun_out/ncu_fulldiv.log 2>&1
cut -c1-1500 gpurun_out/bench_${TAG}.json
