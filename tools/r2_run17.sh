#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for RA in 0 2; do
echo "== RELEASE_AT=$RA"; LOOSB200_RELEASE_AT=$RA timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
done; done > gpurun_out/r2_ab_h.txt 2>&1
for RA in 0 2; do echo "== RELEASE_AT=$RA prof"; LOOSB200_RELEASE_AT=$RA LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4; done >> gpurun_out/r2_ab_h.txt 2>&1
cat gpurun_out/r2_ab_h.txt
