#!/bin/bash
for v in "LOOSB200_CLUSTER=2" "LOOSB200_CLUSTER=1" "LOOSB200_CLUSTER=2 LOOSB200_NO_SAMPLER=1"; do
env $v timeout 300 python bench.py --no-e2e --no-cpu --no-also --peak-seconds 0 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v: step %.2f kernel %.2f diff %.2f' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['ms_per_step']-d['roofline']['kernel_ms_per_launch']))"
done
