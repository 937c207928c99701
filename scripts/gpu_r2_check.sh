# Round-2 validation of the rewritten tile kernel (K1b v2): parity first, then timing, then profiles.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_check.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled or kats or default_rules or chunking or k1c or strips or census or 4096" 2>&1 | tail -5
for A in "--steps 64 --warmup 8" "--steps 20 --warmup 5" "--steps 1000 --warmup 64 --reps 3"; do
timeout 200 python bench.py $A --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1]); print('$A', d['value'], d['timing']['ms_per_repetition'], 'parity', d['parity']['status'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'])" || tail -5 gpurun_out/r2_bench.err
done
if [ "$1" = "ncu" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_r2_tiles python bench.py --steps 64 --warmup 8 --reps 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_tiles.err; tail -3 gpurun_out/ncu_tiles.err
fi
