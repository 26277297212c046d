#!/bin/bash
mkdir -p gpurun_out
(PART=0 timeout 300 python tools/prof_part.py 4 2>&1 | tail -4; PART=7 timeout 300 python tools/prof_part.py 4 2>&1 | tail -2) > gpurun_out/r2_prof_part.txt 2>&1
cat gpurun_out/r2_prof_part.txt
for rep in 1 2; do
echo "== CLUSTER=1 RELEASE_AT=0"; timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
echo "== CLUSTER=2 RELEASE_AT=1"; LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=1 timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
done > gpurun_out/r2_ab_i.txt 2>&1
cat gpurun_out/r2_ab_i.txt
