#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "run_action or step_reward or deferred" 2>&1 | tail -2
LEGS="n1:1 main_4096:1 busy_4096:1" tools/gpu_variants.sh
