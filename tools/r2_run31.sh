#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rmsds.py tests/test_gpu_rmsds_align.py -m gpu -x -q > gpurun_out/r2_pytest_h.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_h.log | tail -6
timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
PART=0 timeout 300 python tools/prof_part.py 4 2>&1 | tail -2
