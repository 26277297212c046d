#!/bin/bash
# rmsds-align on the tensor engine: parity tests, then the 4000 x 4000 x (500 + 500) micro-benchmark of both engines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rmsds_align.py tests/test_gpu_tools.py -m gpu -x -q > gpurun_out/r2_pytest_align.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2_pytest_align.log
timeout 600 python tools/bench_misc.py rmsds_align > gpurun_out/r2_rmsds_align_bench.txt 2>&1; cat gpurun_out/r2_rmsds_align_bench.txt
