#!/bin/bash
# CTA-pair shape (TMEM held through the fp64 part) against the default at larger selections
mkdir -p gpurun_out
for cfg in "20000 3000" "20000 5000" "12000 10000" "40000 2000"; do set -- $cfg
for rep in 1 2; do
echo "== F=$1 N=$2 default"; PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 4 2>&1 | tail -2
echo "== F=$1 N=$2 CLUSTER=2 RELEASE_AT=1"; PS_FRAMES=$1 PS_ATOMS=$2 LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=1 timeout 300 python tools/prof_step.py 4 2>&1 | tail -2
done; done > gpurun_out/r2_ab_pair_largeN.txt 2>&1
cat gpurun_out/r2_ab_pair_largeN.txt
