#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python tools/sanitize_smoke.py 2>&1 | tail -4
timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
