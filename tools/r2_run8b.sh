#!/bin/bash
# 8-GPU call 2: host delivery with link-weighted shares against equal shares, then the bench line
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_multi_${N}gpu.log
LOOSB200_VERBOSE=1 timeout 600 python tools/prof_e2e_multi.py $N > gpurun_out/r2_e2e_shares_${N}gpu.txt 2>&1; cat gpurun_out/r2_e2e_shares_${N}gpu.txt
#timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
#  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_${N}gpu.json") if l.startswith("{")][-1])
    e=d["e2e"]; print("value %.3e ms %.1f frac %.3f | e2e ms %.1f stage %.1f call %.1f d2h %.1f GB/s | hbm %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], e["ms_per_step"], e["stage_ms"], e["call_ms"], e["d2h_gbs_over_call"], e["to_rank0_hbm"]))
    for k,v in (d.get("also") or {}).items(): print(k, "value %.3e ms %.1f frac %.3f parity %s" % (v["value"], v["ms_per_step"], v["roofline"]["frac"], v["parity"]))
except Exception as ex: print("no json", ex)
PY
