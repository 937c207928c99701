#!/bin/bash
# round-2 final single-GPU validation: the whole GPU suite, smoke(), both bench arms, launch list + lit capture for profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_final_bench.json') if l.startswith('{')][-1]); print('bench 20/5', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'launches', d['gpu_launches'], d['clocks'])" || tail -5 gpurun_out/r2_final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; tail -c 600 gpurun_out/r2_final_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:se_step_lit -s 16 -c 1 -f -o gpurun_out/prof_r2_lit_final python scripts/light_probe.py 8192 12 > /dev/null 2>&1
python scripts/light_probe.py 8192 48
python scripts/light_probe.py 4096 100
python scripts/run_configs.py 4 2>&1 | tail -1
python scripts/run_configs.py 1 2>&1 | tail -1
python scripts/run_configs.py 2 2>&1 | tail -1
SE_CFG5_SIZE=16384 python scripts/run_configs.py 5 2>&1 | tail -1
python scripts/k1c_probe.py 2>&1 | grep k1c_probe
