for PH in 254 200 138 100 64; do SE_TILE_PH=$PH timeout 300 python scripts/strip_probe.py 1 0 2>&1 | grep strip_probe | sed "s/^/PH=$PH /"; done
for PH in 254 190 138 100; do SE_TILE_PH=$PH timeout 300 python scripts/strip_probe.py 8 34 2>&1 | grep strip_probe | sed "s/^/PH=$PH /"; done
for T in 4 16; do timeout 300 python scripts/strip_probe.py 8 34 $T 2>&1 | grep strip_probe | sed "s/^/T=$T /"; done
