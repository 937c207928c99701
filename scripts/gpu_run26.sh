timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or strips or default_rules" 2>&1 | tail -2
timeout 300 python scripts/strip_probe.py 1 0 2>&1 | grep strip_probe
for H in 32 34; do timeout 300 python scripts/strip_probe.py 8 $H 2>&1 | grep strip_probe; done
timeout 300 python scripts/strip_probe.py 4 34 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 2 34 2>&1 | grep strip_probe
