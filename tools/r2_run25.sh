#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rmsds.py tests/test_gpu_multi.py tests/test_gpu_tools.py -m gpu -x -q > gpurun_out/r2_pytest_e.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_e.log
(LOOSB200_STRIPS=1 python tools/prof_e2e_bands.py 16 2>&1 | tail -1 | sed 's/^/strips /'; python tools/prof_e2e_bands.py 16 32 64 128 64 2>&1 | tail -5 | sed 's/^/columns /') > gpurun_out/r2_e2e_column_blocks.txt 2>&1
cat gpurun_out/r2_e2e_column_blocks.txt
