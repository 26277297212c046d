#!/bin/bash
mkdir -p gpurun_out
for cfg in "100000 1000" "20000 3000"; do set -- $cfg
for rep in 1 2; do
for v in base cur in2 in3; do
echo "== F=$1 N=$2 $v"; if [ $v = cur ]; then PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1; else AB_LIB=/root/repo/tools/experiments/libloosb200_$v.so PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1; fi
done
echo "== F=$1 N=$2 cur RELEASE_AT=1"; LOOSB200_RELEASE_AT=1 PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
done; done > gpurun_out/r2_ab_n.txt 2>&1
cat gpurun_out/r2_ab_n.txt
