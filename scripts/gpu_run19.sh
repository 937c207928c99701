mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/run_configs.py 4 2>gpurun_out/cfg4.err | tee gpurun_out/cfg4.json
