#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rmsds.py -m gpu -x -q > gpurun_out/r2_pytest_f.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_f.log | tail -8
for rep in 1 2; do
for v in base cur; do
echo "== $v"; if [ $v = cur ]; then timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; else AB_LIB=/root/repo/tools/experiments/libloosb200_$v.so timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; fi
done; done > gpurun_out/r2_ab_l.txt 2>&1
cat gpurun_out/r2_ab_l.txt
echo "== N=60000 F=1500: tensor (segments) vs fp64 engine"
PS_FRAMES=1500 PS_ATOMS=60000 timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
timeout 300 python tools/prof_fp64.py 1500 60000 2>&1 | tail -2
