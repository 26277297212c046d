#!/bin/bash
mkdir -p gpurun_out
for R in 12 24 48 96; do for C in 16 32; do
echo "== SB_R=$R SB_C=$C"
LOOSB200_SB_R=$R LOOSB200_SB_C=$C timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
done; done > gpurun_out/r2_superblock_sweep.txt 2>&1
cat gpurun_out/r2_superblock_sweep.txt
for R in 12 48; do
LOOSB200_SB_R=$R timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_pairs_tensor -c 1 --csv --log-file gpurun_out/r2_ncu_dram_sbr$R.csv python tools/prof_step.py 1 > gpurun_out/r2_ncu_dram_sbr$R.log 2>&1
grep -i "dram__\|duration\|hit_rate\|imma" gpurun_out/r2_ncu_dram_sbr$R.csv | cut -c1-260
done
