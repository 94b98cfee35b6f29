#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="envs_32768:16 envs_32768:12" tools/gpu_variants.sh
