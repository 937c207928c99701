mkdir -p gpurun_out
for N in "$@"; do
if [ "$N" = "1" ]; then
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
fi
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/scale_n$N.json') if l.startswith('{')][-1]); print('N=$N', d['n_gpus'], d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
