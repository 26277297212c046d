#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free, nothing charged)
# usage: tools/gpurun_retry.sh [--gpus N] --timeout S -- 'command'
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_last.txt 2>&1
  if grep -q "status=transient" /tmp/gpurun_last.txt; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.txt
