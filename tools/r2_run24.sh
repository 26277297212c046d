#!/bin/bash
mkdir -p gpurun_out
for v in "" "LOOSB200_NO_SAMPLER=1"; do for rep in 1 2; do
env $v timeout 300 python bench.py --workload c3 --no-e2e --no-cpu --peak-seconds 0 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v c3 step %.2f kernel %.2f launches %d clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['gpu_launches'], (d['clocks'] or {}).get('nvml_query_ms')))"
done; done
PS_FRAMES=20000 PS_ATOMS=3000 timeout 300 python tools/prof_step.py 6 2>&1 | tail -3
