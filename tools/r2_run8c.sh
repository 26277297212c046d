#!/bin/bash
# N GPUs in one process: multi-GPU tests, then the delivery modes of the host-output call side by side
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest_multi_${N}gpu.log | tail -6
LOOSB200_VERBOSE=1 timeout 600 python tools/prof_e2e_multi.py $N > gpurun_out/r2_e2e_modes_${N}gpu.txt 2>&1; cat gpurun_out/r2_e2e_modes_${N}gpu.txt
