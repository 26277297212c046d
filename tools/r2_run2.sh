#!/bin/bash
# round 2, GPU call 2 (N GPUs): the one-process multi-GPU path
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_multi_${N}gpu.log
timeout 300 python tools/prof_d2h_multi.py > gpurun_out/r2_d2h_${N}gpu.txt 2>&1; cat gpurun_out/r2_d2h_${N}gpu.txt
for G in store copy; do
LOOSB200_GATHER=$G timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-also > gpurun_out/r2_bench_${N}gpu_$G.json 2> gpurun_out/r2_bench_${N}gpu_$G.err; echo "bench $G rc=$?"
tail -c 400 gpurun_out/r2_bench_${N}gpu_$G.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${N}gpu_$G.json"))
    e=d["e2e"]; print("value %.3e ms %.1f | e2e ms %.1f call %.1f d2h %.1f GB/s | hbm %s" % (d["value"], d["ms_per_step"], e["ms_per_step"], e["call_ms"], e["d2h_gbs_over_call"], e["to_rank0_hbm"]))
except Exception as ex: print("no json", ex)
PY
done
