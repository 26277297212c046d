#!/bin/bash
# round 2, GPU call 3 (one GPU): 32 dense "b" frames per column tile
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tmem_ld tools/ubench_tmem_ld.cu && /tmp/ubench_tmem_ld > gpurun_out/r2_ubench_tmem_ld.txt 2>&1; cat gpurun_out/r2_ubench_tmem_ld.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_b1.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_b1.log
for CL in 1 2; do
LOOSB200_CLUSTER=$CL LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 > gpurun_out/r2_prof_step_b1_cl$CL.txt 2>&1
tail -4 gpurun_out/r2_prof_step_b1_cl$CL.txt
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-also --no-e2e --no-cpu > gpurun_out/r2_bench_b1.json 2> gpurun_out/r2_bench_b1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_b1.json") if l.startswith("{")][-1])
print("value %.4e ms %.1f kernel %.1f frac %.3f proxy %.3f clocks %s parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["roofline"]["frac_of_proxy"], d["clocks"]["sm_mhz"], d["parity"]["max_abs_err"]))
PY
