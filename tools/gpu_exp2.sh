#!/bin/bash
# parity + quick sweeps after a kernel change
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # leg lanes
  GBENV_LANES=$2 timeout 300 python bench.py --only-leg $1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 L=$2', round(d['value']), round(d['ms_per_step'],2))
"
}
for L in ${LANES_32K:-8 16 32}; do run envs_32768 $L; done
for L in ${LANES_4K:-1 2}; do run main_4096 $L; done
run divergent_4096 1
run n1 1
