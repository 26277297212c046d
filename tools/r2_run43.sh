#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_rmsds.py tests/test_gpu_rmsds_align.py -m gpu -x -q > gpurun_out/r2_pytest_j.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_j.log | tail -4
for cfg in "100000 1000" "20000 5000"; do set -- $cfg
for rep in 1 2; do
echo "== F=$1 N=$2 base"; AB_LIB=/root/repo/tools/experiments/libloosb200_base.so PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
echo "== F=$1 N=$2 cur (warp-uniform producer)"; PS_FRAMES=$1 PS_ATOMS=$2 timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
done; done > gpurun_out/r2_ab_p.txt 2>&1
cat gpurun_out/r2_ab_p.txt
