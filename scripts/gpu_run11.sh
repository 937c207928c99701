set -x
mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
timeout 300 python bench.py --steps 256 --warmup 32 --no-cpu-baseline > gpurun_out/bench11_n1.json 2> gpurun_out/bench11_n1.err; tail -2 gpurun_out/bench11_n1.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 256 --warmup 32 --no-cpu-baseline > gpurun_out/bench11_n$N.json 2> gpurun_out/bench11_n$N.err; tail -3 gpurun_out/bench11_n$N.err | grep -v "^\*\*\*\|OMP_NUM"
fi
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/bench11_n$N.json') if l.startswith('{')][-1]); print('N=$N', d['n_gpus'], d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"
