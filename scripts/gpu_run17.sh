mkdir -p gpurun_out
timeout 600 python scripts/run_configs.py 2 2>gpurun_out/cfg2.err | tee gpurun_out/cfg2.json
timeout 900 python scripts/run_configs.py 4 2>gpurun_out/cfg4.err | tee gpurun_out/cfg4.json
SE_CFG5_SIZE=16384 timeout 600 python scripts/run_configs.py 5 2>gpurun_out/cfg5_1gpu.err | tee gpurun_out/cfg5_1gpu_16384.json
tail -3 gpurun_out/cfg2.err gpurun_out/cfg4.err gpurun_out/cfg5_1gpu.err
