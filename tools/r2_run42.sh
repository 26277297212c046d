#!/bin/bash
for v in "" "--no-also"; do
timeout 900 python bench.py $v --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('bench $v: step %.2f kernel %.2f diff %.2f e2e %.1f' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['ms_per_step']-d['roofline']['kernel_ms_per_launch'], d['e2e']['ms_per_step']))
for k,v in (d.get('also') or {}).items(): print('  ', k, 'step %.2f kernel %.2f' % (v['ms_per_step'], v['roofline']['kernel_ms_per_launch']))"
done
