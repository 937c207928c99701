#!/bin/bash
# A/B of kernel source versions on ONE box: build/ab/kern_<X>.cuh (git-ignored, written before the call), alternating
cd "$(dirname "$0")/.."
for round in $(seq ${AB_ROUNDS:-3}); do
  for v in ${AB_VARIANTS:-A B}; do
    echo -n "$v: "
    SE_KERNEL_SOURCE_FILE=build/ab/kern_$v.cuh timeout 300 python scripts/light_probe.py ${AB_SIZE:-8192} ${AB_FRAMES:-48}
  done
done
