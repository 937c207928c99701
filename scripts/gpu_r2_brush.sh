#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "brush or modifications or census or strips_with_mod or kats or unknown" 2>&1 | tail -4
timeout 300 python scripts/brush_probe.py 16384 40
