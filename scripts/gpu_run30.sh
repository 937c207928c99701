for NS in 0 10000 25000 50000; do SE_TILE_STAGGER_NS=$NS timeout 300 python scripts/strip_probe.py 1 0 2>&1 | grep strip_probe | sed "s/^/stagger=$NS /"; done
for NS in 0 8000 16000 30000; do SE_TILE_STAGGER_NS=$NS timeout 300 python scripts/strip_probe.py 8 34 2>&1 | grep strip_probe | sed "s/^/stagger=$NS /"; done
