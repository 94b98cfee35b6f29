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
run timer_4096 1
run main_4096 1
for L in 2 4 8; do run custom:pokelike,8192,20,5,100,0 $L; done
for L in 4 8 16; do run custom:pokelike,16384,20,5,100,0 $L; done
for L in 16 32; do run custom:pokelike,65536,12,4,60,0 $L; done
run custom:pokelike,94720,12,4,60,0 32
run custom:pokelike,131072,10,3,50,0 32
