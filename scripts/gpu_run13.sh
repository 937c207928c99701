mkdir -p gpurun_out
N=8
for cfg in "16 8" "32 8" "64 8" "32 4" "16 4" "32 16"; do set -- $cfg; H=$1; T=$2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 512 --warmup 64 --halo $H --temporal-block $T --no-cpu-baseline > gpurun_out/bench13_n${N}_h${H}_T$T.json 2> gpurun_out/bench13.err
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/bench13_n${N}_h${H}_T$T.json') if l.startswith('{')][-1]); print('N=$N halo=$H T=$T', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"
done
