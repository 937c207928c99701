# bench.py at the given GPU counts on one box, the driver's way (--steps 20 --warmup 5) and a long run (--steps 1000):
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_scale.sh 1 2 4 8'      -> gpurun_out/scale_{short,long}_n*.json
# and, first, the multi-process parity check at the largest count (scripts/check_strips_multi.py -> gpurun_out/check_strips_N.log)
mkdir -p gpurun_out
MAXN=1; for N in "$@"; do [ "$N" -gt "$MAXN" ] && MAXN=$N; done
if [ "$MAXN" -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $MAXN --master-addr 127.0.0.1 --master-port 29510 scripts/check_strips_multi.py 2>&1 | grep check_strips
fi
for N in "$@"; do
for MODE in short long; do
if [ "$MODE" = "short" ]; then ARGS="--steps 20 --warmup 5"; else ARGS="--steps 1000 --warmup 64 --reps 3"; fi
if [ "$N" = "1" ]; then
timeout 600 python bench.py $ARGS --no-cpu-baseline > gpurun_out/scale_${MODE}_n1.json 2> gpurun_out/scale_${MODE}_n1.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $ARGS --no-cpu-baseline > gpurun_out/scale_${MODE}_n$N.json 2> gpurun_out/scale_${MODE}_n$N.err
fi
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/scale_${MODE}_n$N.json') if l.startswith('{')][-1]); print('$MODE N=$N', d['n_gpus'], d['value'], d['ms_per_step'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], d['parity']['checksum'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/scale_${MODE}_n$N.err
done
done
