mkdir -p gpurun_out
free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/run_configs.py 5 > gpurun_out/cfg5_8gpu.json 2> gpurun_out/cfg5_8gpu.err
cat gpurun_out/cfg5_8gpu.json | grep "^{"; tail -3 gpurun_out/cfg5_8gpu.err | grep -v "^\*\*\*\|OMP_NUM"
