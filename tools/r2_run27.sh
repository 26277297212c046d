#!/bin/bash
# every kernel once at awkward small sizes, round-2 paths included (compute-sanitizer is closed on this pool, so this is
# the plain run: the script compares engines, launch shapes and delivery modes with each other)
mkdir -p gpurun_out
timeout 300 python tools/sanitize_smoke.py > gpurun_out/r2_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -6 gpurun_out/r2_sanitize_plain.log
