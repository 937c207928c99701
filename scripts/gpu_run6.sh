set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for T in 2 4 6 8 12; do
timeout 600 python bench.py --steps 240 --warmup 24 --temporal-block $T --no-cpu-baseline > gpurun_out/bench6_T$T.json 2> gpurun_out/bench6_T$T.err; tail -2 gpurun_out/bench6_T$T.err; python -c "
import json; d=json.load(open('gpurun_out/bench6_T$T.json')); print('T=$T', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"
done
