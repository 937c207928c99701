mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 512 --warmup 64 --no-cpu-baseline --debug-no-exchange > gpurun_out/bench14_noex.json 2> gpurun_out/bench14_noex.err
grep "\[bench\] rank" gpurun_out/bench14_noex.err | sort
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 512 --warmup 64 --no-cpu-baseline > gpurun_out/bench14_ex.json 2> gpurun_out/bench14_ex.err
grep "\[bench\] rank" gpurun_out/bench14_ex.err | sort
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
