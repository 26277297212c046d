#!/bin/bash
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== cur"; timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
echo "== spin (no back-off in the hand-over waits)"; AB_LIB=/root/repo/tools/experiments/libloosb200_spin.so timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
done
for R in 24 48 96; do for C in 16 32; do
echo "== SB_R=$R SB_C=$C"; LOOSB200_SB_R=$R LOOSB200_SB_C=$C timeout 300 python tools/prof_step.py 3 2>&1 | tail -1
done; done
} > gpurun_out/r2_ab_q.txt 2>&1
cat gpurun_out/r2_ab_q.txt
