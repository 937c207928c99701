#!/bin/bash
# fused lit kernel: parity subset + probe, tile height / buffer count variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "light or lit or lighting or mods or modification or ref_goldens or goldens" 2>&1 | tail -3
for v in "32 2" "16 3" "16 2"; do
  set -- $v
  echo "TH=$1 NBUF=$2"
  SE_LF_TH=$1 SE_LF_NBUF=$2 timeout 300 python scripts/light_probe.py 8192 48
  SE_LF_TH=$1 SE_LF_NBUF=$2 timeout 300 python scripts/light_probe.py 4096 100
done
SE_LF_TH=16 SE_LF_NBUF=3 timeout 900 python -m pytest tests -m gpu -x -q -k "light or lit or lighting" 2>&1 | tail -3
