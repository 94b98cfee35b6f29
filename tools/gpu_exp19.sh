#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
LEGS="main_4096:1 n1:1" tools/gpu_variants.sh
for v in new s28; do GBENV_LIB=$PWD/pokegym_b200/csrc/variants/libgbenv_$v.so timeout 200 python tools/exp_groups.py 4096 2 40 2>&1 | tail -1; done
GBENV_LIB=$PWD/pokegym_b200/csrc/variants/libgbenv_new.so timeout 200 python tools/exp_groups.py 94720 1 12 32 2>&1 | tail -1
