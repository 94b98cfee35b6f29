#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # leg lanes defer
  GBENV_DEFER=$3 GBENV_LANES=$2 timeout 300 python bench.py --only-leg $1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 L=$2 defer=$3', round(d['value']), round(d['ms_per_step'],2), 'faults', d['faults'])
"
}
for D in 1 0; do
run main_4096 1 $D
run envs_32768 16 $D
run divergent_32768 16 $D
run n1 1 $D
done
GBENV_DEFER=1 timeout 200 python tools/exp_groups.py 4096 2 40 2>&1 | tail -1
GBENV_DEFER=0 timeout 200 python tools/exp_groups.py 4096 2 40 2>&1 | tail -1
