#!/bin/bash
# setmaxnreg split (control 32 / epilogue 160 registers), mirror-canonical rectangles: correctness, then phase counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rmsds.py tests/test_gpu_multi.py tests/test_gpu_tools.py -m gpu -x -q > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_b.log
for CL in 1 2; do for RA in 0 1; do
echo "== CLUSTER=$CL RELEASE_AT=$RA"
LOOSB200_CLUSTER=$CL LOOSB200_RELEASE_AT=$RA LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4
done; done > gpurun_out/r2_prof_variants_c.txt 2>&1
cat gpurun_out/r2_prof_variants_c.txt
