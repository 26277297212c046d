#!/bin/bash
# full-size C5 (200,000 frames x 5,000 atoms) through the torchrun path on 2 ranks (80 GB of row blocks per rank)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --workload c5 --no-e2e --no-cpu --steps 1 --warmup 1 --peak-seconds 0.5 > gpurun_out/r2_bench_c5_2gpu.json 2> gpurun_out/r2_bench_c5_2gpu.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_c5_2gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_c5_2gpu.json") if l.startswith("{")][-1])
print("c5 on 2 ranks: value %.3e ms %.1f kernel %.1f frac %.3f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["parity"]))
PY
