# one 8-GPU box: multi-process parity, the driver's scaling run (--steps 20 --warmup 5) at 8, 4, 2 and 1 ranks, a long run at 8, config 5 at 65536^2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29510 scripts/check_strips_multi.py 2>&1 | grep check_strips
bash scripts/gpu_scale_quick.sh 8 4 2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_short_n1.json 2> gpurun_out/scale_short_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/scale_short_n1.json') if l.startswith('{')][-1]); print('short N=1', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/scale_short_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/run_configs.py 5 2>&1 | grep '"config"' | tee gpurun_out/config5_8gpu.json
