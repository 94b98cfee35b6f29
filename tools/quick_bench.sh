#!/bin/bash
# quick_bench.sh -- short throughput sweep used while tuning k_run_frames (not the judged bench line).
# usage: tools/quick_bench.sh "<lanes list>" "<envs list>"
cd "$(dirname "$0")/.."
LANES=${1:-"1 2 4"}
ENVS=${2:-"4096"}
for N in $ENVS; do for L in $LANES; do
  GBENV_LANES=$L timeout 300 python bench.py --envs-per-gpu $N --steps 12 --warmup 3 --cpu-baseline-seconds 0.5 \
      --e2e-steps 2 --also-envs 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('N=$N L=$L', round(d['value']), d['ms_per_step'], round(d['emulated_instr_per_s'] / 1e9, 2))
"
done; done
