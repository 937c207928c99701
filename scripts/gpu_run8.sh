set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 300 python bench.py --steps 240 --warmup 24 --temporal-block 8 --no-cpu-baseline > gpurun_out/bench8_n1.json 2> gpurun_out/bench8_n1.err; tail -2 gpurun_out/bench8_n1.err
for H in 16 32 64; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 32 --temporal-block 8 --halo $H --no-cpu-baseline > gpurun_out/bench8_n2_h$H.json 2> gpurun_out/bench8_n2_h$H.err; tail -3 gpurun_out/bench8_n2_h$H.err
done
for f in gpurun_out/bench8_*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"; done
