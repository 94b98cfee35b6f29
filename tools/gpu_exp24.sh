#!/bin/bash
cd "$(dirname "$0")/.."
for G in 1 2 4 8; do timeout 200 python tools/exp_groups.py 4096 $G 60 2>&1 | tail -1; done
