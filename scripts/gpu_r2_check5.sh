mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_goldens.py tests/test_gpu_long_horizon.py -x -q -m gpu 2>&1 | tail -4
for A in "--steps 20 --warmup 5" "--steps 1000 --warmup 64 --reps 3"; do
timeout 200 python bench.py $A --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1]); print('$A', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/r2_bench.err
done
python scripts/k1c_probe.py
python scripts/light_probe.py 8192 48
python scripts/light_probe.py 4096 100
timeout 200 ncu --set full --clock-control none --import-source on -k regex:se_step_lit -s 16 -c 1 -o gpurun_out/prof_r2_lit python scripts/light_probe.py 8192 12 > /dev/null 2>&1
