#!/bin/bash
cd "$(dirname "$0")/.."
for G in 1 2 4; do timeout 200 python tools/exp_groups.py 4096 $G 40 2>&1 | tail -1; done
for G in 1 2 4; do timeout 200 python tools/exp_groups.py 32768 $G 20 16 2>&1 | tail -1; done
for G in 2 4; do timeout 200 python tools/exp_groups.py 32768 $G 20 8 2>&1 | tail -1; done
