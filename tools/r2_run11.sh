#!/bin/bash
mkdir -p gpurun_out
for SP in 1 0; do for RA in 0 1; do
echo "== REGSPLIT=$SP RELEASE_AT=$RA (prof)"
LOOSB200_REGSPLIT=$SP LOOSB200_RELEASE_AT=$RA LOOSB200_PROF=1 timeout 300 python tools/prof_step.py 2 2>&1 | tail -4
done
echo "== REGSPLIT=$SP (no prof)"
LOOSB200_REGSPLIT=$SP timeout 300 python tools/prof_step.py 3 2>&1 | tail -2
done > gpurun_out/r2_prof_variants_d.txt 2>&1
cat gpurun_out/r2_prof_variants_d.txt
