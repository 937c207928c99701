#!/bin/bash
# how long do the timed repetitions of bench.py take to settle? (per-repetition device times, 25 repetitions)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --reps 25 --no-cpu-baseline > gpurun_out/ramp.json 2> gpurun_out/ramp.err
grep "device times" gpurun_out/ramp.err | cut -c1-400
SE_TILE_PH=192 timeout 300 python bench.py --steps 20 --warmup 5 --reps 25 --no-cpu-baseline > gpurun_out/ramp192.json 2> gpurun_out/ramp192.err
grep "device times" gpurun_out/ramp192.err | cut -c1-400
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
