mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "strips or census or k1c or tiled" 2>&1 | tail -4
for A in "--steps 20 --warmup 5" "--steps 20 --warmup 5 --temporal-block 10" "--steps 20 --warmup 5 --temporal-block 6"; do
timeout 200 python bench.py $A --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1]); print('$A', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/r2_bench.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_r2_tiles python bench.py --steps 64 --warmup 8 --reps 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_tiles.err; tail -2 gpurun_out/ncu_tiles.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu-baseline > /dev/null 2>&1
