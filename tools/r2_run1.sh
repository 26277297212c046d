#!/bin/bash
# round 2, GPU call 1 (one GPU): tests, default bench line, phase counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_1gpu.err
LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 > gpurun_out/r2_prof_step.txt 2>&1
tail -8 gpurun_out/r2_prof_step.txt
timeout 300 python tools/prof_d2h_multi.py > gpurun_out/r2_d2h_1gpu.txt 2>&1; cat gpurun_out/r2_d2h_1gpu.txt
