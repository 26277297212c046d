#!/bin/bash
mkdir -p gpurun_out
for CL in 1 2; do for RA in 0 1; do
echo "== CLUSTER=$CL RELEASE_AT=$RA"
LOOSB200_CLUSTER=$CL LOOSB200_RELEASE_AT=$RA LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4
done; done > gpurun_out/r2_prof_variants_b1.txt 2>&1
cat gpurun_out/r2_prof_variants_b1.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for CS in 1 2; do
LOOSB200_COPY_STREAMS=$CS timeout 600 python tools/prof_e2e.py 2>&1 | tail -4
done > gpurun_out/r2_e2e_copy_streams.txt 2>&1
cat gpurun_out/r2_e2e_copy_streams.txt
