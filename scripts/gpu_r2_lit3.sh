#!/bin/bash
# fused lit kernel: parity subset + probe; threads per half / rows per thread / buffer count variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "light or lit or lighting or mods or modification or ref_goldens or goldens" 2>&1 | tail -3
for v in ${LIT_VARIANTS:-"384 4 2" "512 4 2" "512 2 3"}; do
  set -- $v
  echo "HALF=$1 ROWS=$2 NBUF=$3"
  SE_LF_HALF=$1 SE_LF_ROWS=$2 SE_LF_NBUF=$3 timeout 300 python scripts/light_probe.py 8192 48
  SE_LF_HALF=$1 SE_LF_ROWS=$2 SE_LF_NBUF=$3 timeout 300 python scripts/light_probe.py 4096 100
done
