#!/bin/bash
# tuning helper: build the CUDA library to /tmp/libgbenv_new.so with ptxas statistics and dump k_run_frames' SASS to /tmp/krf_new.sass
cd /root/repo/pokegym_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v $EXTRA -o /tmp/libgbenv_new.so gbenv.cu 2>&1 | grep -E "error|k_run_frames|stack frame|Used" | grep -A2 "k_run_frames" | grep -E "error|stack|Used"
cuobjdump -sass /tmp/libgbenv_new.so > /tmp/sass_new.txt 2>&1
for f in _Z12k_run_frames9RunParams _Z14k_run_frames_19RunParams; do
L=$(grep -n "Function : $f" /tmp/sass_new.txt | cut -d: -f1)
sed -n "$L,\$p" /tmp/sass_new.txt | awk 'NR>1 && /Function :/ {exit} {print}' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/krf_$f.sass
done
mv /tmp/krf__Z12k_run_frames9RunParams.sass /tmp/krf_new.sass; mv /tmp/krf__Z14k_run_frames_19RunParams.sass /tmp/krf1_new.sass
wc -l /tmp/krf_new.sass /tmp/krf1_new.sass
