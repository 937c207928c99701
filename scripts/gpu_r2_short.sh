#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/short_probe.py 16384 2048
python scripts/short_probe.py 16384 4096
python scripts/short_probe.py 16384 16384
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:se_step_tiles -c 12 --csv --log-file gpurun_out/short_launches.csv python scripts/short_probe.py 16384 2048 > /dev/null 2>&1
grep se_step_tiles gpurun_out/short_launches.csv | awk -F, '{print $(NF)}' | tr '\n' ' '
