#!/bin/bash
# BASELINE.json config 2 at full length (72 envs x 10,000 steps, bit-exact vs 72 oracle instances) and config 1 (test.py protocol)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( echo "# GBENV_LONG_STEPS=10000 python -m pytest tests/test_gpu_parity.py -k long_run -x -q   (B200, final round-2 kernel: deferred PPU, 80-register lock-step build, single-thread-block build)"; GBENV_LONG_STEPS=10000 timeout 900 python -m pytest tests/test_gpu_parity.py -k long_run -x -q 2>&1 | tail -3 ) > gpurun_out/r2_parity_72envs_10000steps.log
cat gpurun_out/r2_parity_72envs_10000steps.log
timeout 600 python bench.py --config single > gpurun_out/bench_config1_r2c.json 2> gpurun_out/bench_config1_r2c.err; echo rc $?; cut -c1-600 gpurun_out/bench_config1_r2c.json
