#!/bin/bash
# validation: the whole GPU suite, smoke, a reduced C5 through the default (CTA-pair) shape, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log
timeout 600 python bench.py --workload c5 --frames 24000 --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_reduced_1gpu.json 2> gpurun_out/r2_bench_c5_reduced.err; echo "c5 reduced rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_c5_reduced_1gpu.json") if l.startswith("{")][-1])
print("c5 reduced: value %.3e ms %.1f frac %.3f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity"]))
PY
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_1gpu.json") if l.startswith("{")][-1])
e=d["e2e"]; print("value %.3e ms %.1f kernel %.1f frac %.3f (proxy %.3f) | e2e ms %.1f | cpu %.3e | parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["roofline"]["frac_of_proxy"], e["ms_per_step"], d["cpu_baseline"]["value"], d["parity"]["max_abs_err"]))
for k,v in (d.get("also") or {}).items(): print(k, "value %.3e ms %.2f frac %.3f parity %s" % (v["value"], v["ms_per_step"], v["roofline"]["frac"], v["parity"]["max_abs_err"]))
PY
