mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/strip_probe.py 1 0 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 8 64 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 8 34 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 4 34 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 2 34 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 8 34 6 2>&1 | grep strip_probe
timeout 300 python scripts/strip_probe.py 8 34 12 2>&1 | grep strip_probe
