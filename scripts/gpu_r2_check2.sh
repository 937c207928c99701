mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1]); print('20/5', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/r2_bench.err
python scripts/k1c_probe.py
SE_CFG5_SIZE=16384 timeout 300 python scripts/run_configs.py 5 2>&1 | tail -2
SE_LUT_FORCE_MODE=0 SE_CFG5_SIZE=16384 timeout 300 python scripts/run_configs.py 5 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
