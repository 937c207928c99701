#!/bin/bash
# K1c variants: parity subset + probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "k1c or census or unknown or kats or eligible or per_frame or odd" 2>&1 | tail -3
timeout 300 python scripts/k1c_probe.py 2>&1 | grep k1c_probe
