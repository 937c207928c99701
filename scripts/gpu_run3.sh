set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for T in 1 2 4 6 8; do
timeout 600 python bench.py --steps 240 --warmup 24 --temporal-block $T --no-cpu-baseline > gpurun_out/bench3_T$T.json 2> gpurun_out/bench3_T$T.err; tail -2 gpurun_out/bench3_T$T.err; python -c "
import json; d=json.load(open('gpurun_out/bench3_T$T.json')); print('T=$T', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['gpu_launches'])"
done
nvidia-smi --help-query-gpu | grep -iE "reasons|throttle" | head -20
