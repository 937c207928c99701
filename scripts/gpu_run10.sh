set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strips" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_strips_multi.py 2>&1 | grep check_strips
timeout 300 python bench.py --steps 256 --warmup 32 --temporal-block 8 --no-cpu-baseline > gpurun_out/bench10_n1.json 2> gpurun_out/bench10_n1.err; tail -2 gpurun_out/bench10_n1.err
for H in 32 64; do for HS in "" "--host-sync"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 32 --temporal-block 8 --halo $H $HS --no-cpu-baseline > gpurun_out/bench10_n2_h$H$HS.json 2> gpurun_out/bench10_n2.err; tail -3 gpurun_out/bench10_n2.err | grep -v "^\*\*\*\|OMP_NUM"
done; done
for f in gpurun_out/bench10_*.json; do python -c "
import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"; done
