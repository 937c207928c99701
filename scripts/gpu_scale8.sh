# one 8-GPU box: multi-process parity, the driver's scaling run (--steps 20 --warmup 5) at 8 and 4 ranks, a long run at 8, config 5 at 65536^2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29510 scripts/check_strips_multi.py 2>&1 | grep check_strips
bash scripts/gpu_scale_quick.sh 8 4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/run_configs.py 5 2>&1 | grep '"config"' | tee gpurun_out/config5_8gpu.json
