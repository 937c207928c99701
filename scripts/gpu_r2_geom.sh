#!/bin/bash
# new tile-geometry cost model: parity of the tile kernel paths, bench at the driver's flags, strip-sized launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long_horizon.py -m gpu -x -q -k "tiled or tile or strips or long or horizon or full or 16384 or two_table or eligible or global" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/geom_bench.json 2> gpurun_out/geom_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/geom_bench.json') if l.startswith('{')][-1]); print('bench 20/5', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/geom_bench.err
timeout 300 python bench.py --steps 1000 --warmup 64 --reps 3 --no-cpu-baseline > gpurun_out/geom_bench_long.json 2> gpurun_out/geom_bench_long.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/geom_bench_long.json') if l.startswith('{')][-1]); print('bench 1000/64', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'])" || tail -5 gpurun_out/geom_bench_long.err
python scripts/short_probe.py 16384 2048
python scripts/short_probe.py 16384 4096
SE_CFG5_SIZE=16384 python scripts/run_configs.py 5 2>&1 | tail -1 | cut -c1-400
