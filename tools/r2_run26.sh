#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in cur r1late; do
echo "== $v"; if [ $v = cur ]; then timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; else AB_LIB=/root/repo/tools/experiments/libloosb200_$v.so timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; fi
done; done > gpurun_out/r2_ab_k.txt 2>&1
echo "== r1late prof" >> gpurun_out/r2_ab_k.txt
AB_LIB=/root/repo/tools/experiments/libloosb200_r1late.so LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4 >> gpurun_out/r2_ab_k.txt
cat gpurun_out/r2_ab_k.txt
