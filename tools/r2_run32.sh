#!/bin/bash
# is the large-N kernel power-limited?  SM clock under load for the two launch shapes at 20,000 x 5,000
mkdir -p gpurun_out
for v in "LOOSB200_CLUSTER=1 LOOSB200_RELEASE_AT=0" "LOOSB200_CLUSTER=2 LOOSB200_RELEASE_AT=1"; do for rep in 1 2; do
env $v timeout 300 python bench.py --workload c3 --atoms 5000 --no-e2e --no-cpu --peak-seconds 0 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v: step %.2f kernel %.2f frac_of_proxy %.3f clocks %s %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac_of_proxy'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done; done > gpurun_out/r2_power_shapes.txt 2>&1
cat gpurun_out/r2_power_shapes.txt
nvidia-smi --query-gpu=power.limit,power.max_limit,power.default_limit --format=csv
