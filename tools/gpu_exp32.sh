#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="main_4096:1 busy_4096:1 divergent_4096:1" tools/gpu_variants.sh
for v in a_base b_hotfirst; do GBENV_LIB=$PWD/pokegym_b200/csrc/variants/libgbenv_$v.so timeout 200 python tools/exp_groups.py 4096 2 40 2>&1 | tail -1; done
