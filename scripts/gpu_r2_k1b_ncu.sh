#!/bin/bash
# one ncu --set full capture of the timed 20-step se_step_tiles launch of bench.py (final tile geometry)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 140 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -f -o gpurun_out/prof_r2_tiles_final python bench.py --steps 20 --warmup 5 --reps 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_r2_tiles_final.ncu-rep
