#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum
for v in "LOOSB200_CLUSTER=1 LOOSB200_RELEASE_AT=0" "LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=1"; do
echo "== $v"
env $v PS_FRAMES=20000 PS_ATOMS=5000 timeout 600 ncu --metrics $M --clock-control none -k regex:k_pairs_tensor -c 1 --csv python tools/prof_step.py 1 2>/dev/null | grep -E "^\"[0-9]" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done > gpurun_out/r2_ncu_shapes_N5000.txt 2>&1
cat gpurun_out/r2_ncu_shapes_N5000.txt
