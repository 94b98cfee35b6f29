#!/bin/bash
# Round evidence (run under gpurun, one GPU): default bench line, ncu launch list, one --set full capture of the
# dominant kernel.  Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarise_profile.py.
set -e
TAG=${TAG:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ARGS="--steps 4 --warmup 3 --cpu-baseline-seconds 0.5 --e2e-steps 3"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 5 -c 1 -o gpurun_out/prof_${TAG} python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
cut -c1-300 gpurun_out/bench_${TAG}.json
