#!/bin/bash
# fused lit kernel: parity subset + probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "light or lit or lighting or mods or modification or ref_goldens or goldens" 2>&1 | tail -5
timeout 300 python scripts/light_probe.py 8192 48
timeout 300 python scripts/light_probe.py 4096 100
timeout 300 python scripts/run_configs.py 4 2>&1 | tail -2
