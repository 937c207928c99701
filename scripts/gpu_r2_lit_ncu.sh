#!/bin/bash
# one ncu capture of se_step_lit (scripts/ncu_summary.py reads it)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:se_step_lit -s 16 -c 1 -f -o gpurun_out/prof_r2_lit4 python scripts/light_probe.py 8192 12 > /dev/null 2>&1
ls -la gpurun_out/prof_r2_lit4.ncu-rep
