mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/check_strips_multi.py 2>&1 | grep check_strips
bash scripts/gpu_scale.sh 8 4 2 1
