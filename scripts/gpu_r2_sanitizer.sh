#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitizer_probe.py 2>&1 | tail -12 | tee gpurun_out/sanitizer_memcheck.txt
timeout 700 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 5 python scripts/sanitizer_probe.py 2>&1 | tail -14 | tee gpurun_out/sanitizer_racecheck.txt
