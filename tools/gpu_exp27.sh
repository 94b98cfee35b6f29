#!/bin/bash
# steady-state check: the headline value over 20 and over 2,000 timed steps (after the 200 pre-roll steps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A="--legs none --also-envs 0 --cpu-baseline-seconds 1 --e2e-steps 5"
timeout 300 python bench.py --steps 20 --warmup 5 $A > gpurun_out/bench_steps20_r2e.json 2>/dev/null
timeout 600 python bench.py --steps 2000 --warmup 5 $A > gpurun_out/bench_steps2000_r2e.json 2>/dev/null
python - <<'PY'
import json
for n in (20, 2000):
    d = json.load(open(f'gpurun_out/bench_steps{n}_r2e.json')); print(n, 'steps:', round(d['value']), 'env-steps/s', round(d['ms_per_step'], 3), 'ms/step', 'issue' in d, d['roofline']['traffic'])
PY
