#!/bin/bash
# launch list of the bench command restricted to this library's kernels (the synthetic-data torch kernels and the
# peak micro-benchmark would fill the -c budget otherwise); the same command first without ncu
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-also --no-cpu --peak-seconds 0 > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2_launches_c4_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-also --no-cpu --peak-seconds 0 > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
grep -c k_pairs_tensor gpurun_out/r2_launches_c4_bench.csv
