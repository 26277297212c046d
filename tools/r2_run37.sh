#!/bin/bash
mkdir -p gpurun_out
for cfg in "100000 1000" "20000 3000" "20000 5000"; do set -- $cfg
for rep in 1 2; do
echo "== F=$1 N=$2 CLUSTER=1"; LOOSB200_CLUSTER=1 PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
echo "== F=$1 N=$2 CLUSTER=2 RELEASE_AT=1"; LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=1 PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
echo "== F=$1 N=$2 CLUSTER=2 RELEASE_AT=0"; LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=0 PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
done; done > gpurun_out/r2_ab_o.txt 2>&1
cat gpurun_out/r2_ab_o.txt
LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_o.txt
LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=1 LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_o.txt
