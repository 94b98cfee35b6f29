#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="envs_32768:16 divergent_32768:16" tools/gpu_variants.sh
