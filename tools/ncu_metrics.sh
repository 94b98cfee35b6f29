#!/bin/bash
# quick ncu metric grab for one k_run_frames launch: usage  LANES=4 ENVS=4096 TAG=x bash tools/ncu_metrics.sh
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --cpu-baseline-seconds 0.2 --e2e-steps 3 --envs-per-gpu ${ENVS:-4096}"
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
GBENV_LANES=${LANES:-32} python bench.py $ARGS > gpurun_out/plain_${TAG}.log 2>&1 &&
GBENV_LANES=${LANES:-32} ncu --metrics $M --clock-control none -k regex:k_run_frames -s 5 -c 1 --csv --log-file gpurun_out/metrics_${TAG}.csv python bench.py $ARGS > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/metrics_${TAG}.csv')) if len(r)>10]
for r in rows[1:]:
    print('${TAG}', r[-3].replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), r[-1])
PY
