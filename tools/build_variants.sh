#!/bin/bash
# A/B libraries of the emulation kernel for tuning runs (tools/gpu_variants.sh); not part of the shipped build.
# usage: tools/build_variants.sh <file with lines "name -Dflag ...">
cd "$(dirname "$0")/../pokegym_b200/csrc"
mkdir -p variants
while read n f; do
  [ -z "$n" ] && continue
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $f -o variants/libgbenv_$n.so gbenv.cu 2>&1 | grep -i "error" &
done < "$1"
wait
ls variants
