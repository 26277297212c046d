#!/bin/bash
# final build on 2 GPUs: multi-GPU tests, the bench line under torchrun, the reference arm under torchrun
N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_multi_${N}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_${N}gpu.json") if l.startswith("{")][-1])
e=d["e2e"]; print("value %.3e ms %.1f kernel %.1f frac %.3f | e2e ms %.1f stage %.1f call %.1f d2h %.1f GB/s | hbm %.1f ms" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], e["ms_per_step"], e["stage_ms"], e["call_ms"], e["d2h_gbs_over_call"], e["to_rank0_hbm"]["ms_per_step"]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 1 --warmup 1 --ref-sample 1500 2>/dev/null | tail -1 | cut -c1-300
