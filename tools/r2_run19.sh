#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in cur spin spin3 l2; do
echo "== $v"; if [ $v = cur ]; then timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; else AB_LIB=/root/repo/tools/experiments/libloosb200_$v.so timeout 300 python tools/prof_step.py 3 2>&1 | tail -2; fi
done; done > gpurun_out/r2_ab_j.txt 2>&1
cat gpurun_out/r2_ab_j.txt
