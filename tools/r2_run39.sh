#!/bin/bash
# final build: default bench line, then ncu --set full of one C4 launch (after the same command ran without ncu)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_1gpu.json") if l.startswith("{")][-1])
e=d["e2e"]; print("value %.3e ms %.1f kernel %.1f frac %.3f (proxy %.3f) | e2e ms %.1f | cpu %.3e | parity %s | clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["roofline"]["frac_of_proxy"], e["ms_per_step"], d["cpu_baseline"]["value"], d["parity"]["max_abs_err"], d["clocks"]["sm_mhz"]))
for k,v in (d.get("also") or {}).items(): print(k, "value %.3e ms %.2f frac %.3f parity %s" % (v["value"], v["ms_per_step"], v["roofline"]["frac"], v["parity"]["max_abs_err"]))
PY
timeout 300 python tools/prof_step.py 1 > /dev/null 2>&1; echo "prof_step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_tensor -c 1 -f -o gpurun_out/r2_k_pairs_tensor_C4_final_full \
  python tools/prof_step.py 1 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4
