#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="envs_32768:16 envs_32768:32 divergent_32768:16" tools/gpu_variants.sh
GBENV_LIB=$PWD/pokegym_b200/csrc/variants/libgbenv_simtcls.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "config5 or full_size" 2>&1 | tail -2
