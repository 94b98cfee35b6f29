#!/bin/bash
# runs the legs given in $LEGS ("leg:lanes ...") for every library under pokegym_b200/csrc/variants
cd "$(dirname "$0")/.."
LEGS=${LEGS:-"main_4096:1 envs_32768:16"}
for so in pokegym_b200/csrc/variants/libgbenv_*.so; do
  v=$(basename $so .so); v=${v#libgbenv_}
  for ll in $LEGS; do
    leg=${ll%%:*}; L=${ll##*:}
    GBENV_LIB=$PWD/$so GBENV_LANES=$L timeout 300 python bench.py --only-leg $leg 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v $leg L=$L', round(d['value']), round(d['ms_per_step'],2))
"
  done
done
