#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rmsds_align.py -m gpu -x -q > gpurun_out/r2_pytest_g.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_g.log | tail -8
