#!/bin/bash
# counters at other lane counts: is the 8-lane build at 32,768 envs issue bound or latency bound?
cd "$(dirname "$0")/.."
LANES=8 ENVS=32768 TAG=l8_32k bash tools/ncu_metrics.sh
LANES=32 ENVS=32768 TAG=l32_32k bash tools/ncu_metrics.sh
LANES=2 ENVS=4096 TAG=l2_4k bash tools/ncu_metrics.sh
