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
run main_4096 2
run envs_32768 8
run envs_32768 16
run envs_32768 32
run divergent_32768 16
run custom:pokelike,94720,12,4,60,0 32
