mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1c or census or strips or unknown or kats or eligible or two_table or light or lit" 2>&1 | tail -4
python scripts/k1c_probe.py
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1]); print('20/5', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/r2_bench.err
python scripts/light_probe.py 8192 48
python scripts/light_probe.py 4096 100
