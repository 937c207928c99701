#!/bin/bash
# fused lit kernel: parity subset + probe; variants NG_HALF_ROWS_NBUF (groups per CTA, threads per group, rows per thread, TMA buffers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "light or lit or lighting or mods or modification or ref_goldens or goldens" 2>&1 | tail -3
for v in ${LIT_VARIANTS:-3_256_4_2}; do
  IFS=_ read ng half rows nbuf <<< "$v"
  echo "NG=$ng HALF=$half ROWS=$rows NBUF=$nbuf"
  SE_LF_NG=$ng SE_LF_HALF=$half SE_LF_ROWS=$rows SE_LF_NBUF=$nbuf timeout 300 python scripts/light_probe.py 8192 48
  SE_LF_NG=$ng SE_LF_HALF=$half SE_LF_ROWS=$rows SE_LF_NBUF=$nbuf timeout 300 python scripts/light_probe.py 4096 100
done
