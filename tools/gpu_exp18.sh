#!/bin/bash
cd "$(dirname "$0")/.."
LEGS="envs_32768:16 custom:pokelike,94720,12,4,60,0:32 divergent_32768:16" tools/gpu_variants.sh
