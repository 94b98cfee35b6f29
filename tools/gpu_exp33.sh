#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="envs_32768:16 main_4096:1" tools/gpu_variants.sh
