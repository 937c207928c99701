#!/bin/bash
# what a short K1b launch costs on a strip-sized grid, over T-block lengths and tile heights
cd "$(dirname "$0")/.."
python scripts/short_probe.py 16384 2048
for T in 20 10 7 5; do python scripts/short_probe.py 16384 2048 $T; done
for PH in 218 184 240 282; do SE_TILE_PH=$PH python scripts/short_probe.py 16384 2048; done
python scripts/short_probe.py 16384 4096
SE_TILE_PH=286 python scripts/short_probe.py 16384 4096
