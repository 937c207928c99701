#!/bin/bash
# the driver's scaling run only: bench.py --steps 20 --warmup 5 at the given GPU counts on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_short_n1.json 2> gpurun_out/scale_short_n1.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_short_n$N.json 2> gpurun_out/scale_short_n$N.err
  fi
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/scale_short_n$N.json') if l.startswith('{')][-1]); print('short N=$N', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/scale_short_n$N.err
done
