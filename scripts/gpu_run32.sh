mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or strips or default_rules or chunking or phase" > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python scripts/strip_probe.py 1 0 2>&1 | grep strip_probe
timeout 120 python scripts/strip_probe.py 8 34 2>&1 | grep strip_probe
timeout 120 python scripts/strip_probe.py 4 34 2>&1 | grep strip_probe
timeout 120 python scripts/strip_probe.py 2 34 2>&1 | grep strip_probe
