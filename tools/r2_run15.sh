#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rmsds.py tests/test_gpu_rmsds_align.py -m gpu -x -q > gpurun_out/r2_pytest_d.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_d.log
for rep in 1 2; do
for v in base cur; do
echo "== $v"; if [ $v = cur ]; then timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; else AB_LIB=/root/repo/tools/experiments/libloosb200_$v.so timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; fi
done; done > gpurun_out/r2_ab_g.txt 2>&1
cat gpurun_out/r2_ab_g.txt
for RA in 0 1; do echo "== RELEASE_AT=$RA"; LOOSB200_RELEASE_AT=$RA LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4; done > gpurun_out/r2_prof_g.txt 2>&1
cat gpurun_out/r2_prof_g.txt
