#!/bin/bash
cd "$(dirname "$0")/.."
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio
for v in old new; do
for D in 0 1; do
GBENV_DEFER=$D GBENV_LANES=16 GBENV_LIB=$PWD/pokegym_b200/csrc/variants/libgbenv_$v.so timeout 300 ncu --metrics $M --clock-control none -k regex:k_run_frames -s 110 -c 1 --csv --log-file gpurun_out/m17_${v}_$D.csv python bench.py --only-leg envs_32768 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/m17_${v}_$D.csv')) if len(r)>10]
print('$v defer=$D', ' '.join(r[-3].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').replace('.avg.pct_of_peak_sustained_active','')+'='+r[-1] for r in rows[1:]))
PY
done
done
