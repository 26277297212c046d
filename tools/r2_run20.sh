#!/bin/bash
# round 2, profiling call: launch list of the default bench command, ncu --set full of the C4 kernel and of the two
# rmsds-align launches (each only after the same command has exited 0 without ncu)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-also --no-cpu --peak-seconds 0.5 > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-also --no-cpu --peak-seconds 0.5 > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/prof_step.py 1 > /dev/null 2>&1; echo "prof_step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_tensor -c 1 -f -o gpurun_out/r2_k_pairs_tensor_C4_full \
  python tools/prof_step.py 1 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ALIGN_FRAMES=4000 timeout 300 python tools/bench_misc.py rmsds_align > /dev/null 2>&1; echo "align rc=$?"
ALIGN_FRAMES=4000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_tensor -c 2 -f -o gpurun_out/r2_k_pairs_tensor_align_F4000_full \
  python tools/bench_misc.py rmsds_align > gpurun_out/r2_ncu_align.log 2>&1; echo "ncu align rc=$?"
ls -la gpurun_out/*.ncu-rep
