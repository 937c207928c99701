# quick multi-GPU timing: bench.py --steps 20 --warmup 5 (the driver's flags) at the given GPU counts, then a long run at the largest
mkdir -p gpurun_out
for N in "$@"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_short_n$N.json 2> gpurun_out/scale_short_n$N.err
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/scale_short_n$N.json') if l.startswith('{')][-1]); print('short N=$N', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/scale_short_n$N.err
grep "device times" gpurun_out/scale_short_n$N.err | head -8
done
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1000 --warmup 64 --reps 3 --no-cpu-baseline > gpurun_out/scale_long_n$N.json 2> gpurun_out/scale_long_n$N.err
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/scale_long_n$N.json') if l.startswith('{')][-1]); print('long N=$N', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], d['parity']['checksum'], 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/scale_long_n$N.err
