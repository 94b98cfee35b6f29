#!/bin/bash
# lanes sweep at 32,768 envs (steady state) + ncu capture of the 32,768-env kernel
cd "$(dirname "$0")/.."
for L in 8 16 32; do
  GBENV_LANES=$L timeout 300 python bench.py --only-leg envs_32768 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('envs_32768 L=$L', round(d['value']), round(d['ms_per_step'],2))
"
done
for L in 8 32; do
  GBENV_LANES=$L timeout 300 python bench.py --only-leg divergent_32768 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('divergent_32768 L=$L', round(d['value']), round(d['ms_per_step'],2))
"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_frames -s 112 -c 1 -o gpurun_out/prof_r2d_32k python bench.py --only-leg envs_32768 > gpurun_out/ncu_r2d.log 2>&1
tail -2 gpurun_out/ncu_r2d.log
