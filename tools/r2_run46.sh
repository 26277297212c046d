#!/bin/bash
mkdir -p gpurun_out
ALIGN_FRAMES=4000 timeout 300 python tools/bench_misc.py rmsds_align > gpurun_out/r2_rmsds_align_bench.txt 2>&1; cat gpurun_out/r2_rmsds_align_bench.txt
timeout 300 python tools/prof_text.py 2>&1 | tail -3
