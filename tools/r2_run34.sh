#!/bin/bash
mkdir -p gpurun_out
AB_LIB=/root/repo/tools/experiments/libloosb200_k64.so timeout 300 python tools/ab_check.py 2>&1 | tail -7
for cfg in "100000 1000" "20000 5000"; do set -- $cfg
for rep in 1 2; do
echo "== F=$1 N=$2 cur"; PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
echo "== F=$1 N=$2 k64"; AB_LIB=/root/repo/tools/experiments/libloosb200_k64.so PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
done; done > gpurun_out/r2_ab_k64.txt 2>&1
cat gpurun_out/r2_ab_k64.txt
AB_LIB=/root/repo/tools/experiments/libloosb200_k64.so LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4
