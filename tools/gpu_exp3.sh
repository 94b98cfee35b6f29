#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # leg lanes
  GBENV_LANES=$2 timeout 300 python bench.py --only-leg $1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 L=$2', round(d['value']), round(d['ms_per_step'],2), 'faults', d['faults'])
"
}
run main_4096 1
GBENV_NO_SINGLE=1 run main_4096 1
run envs_32768 16
run divergent_4096 1
run n72 1
run n1 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 112 -c 1 -o gpurun_out/prof_r2e_32k python bench.py --only-leg envs_32768 > gpurun_out/ncu_r2e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 160 -c 1 -o gpurun_out/prof_r2e_4k python bench.py --only-leg main_4096 > gpurun_out/ncu_r2e4k.log 2>&1
tail -1 gpurun_out/ncu_r2e4k.log
