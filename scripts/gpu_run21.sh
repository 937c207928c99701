mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench21.json 2> gpurun_out/bench21.err; tail -3 gpurun_out/bench21.err
python -c "
import json; d=json.load(open('gpurun_out/bench21.json')); print('N=1', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'], d['clocks'])"
timeout 900 python scripts/run_configs.py 4 2>gpurun_out/cfg4.err | tee gpurun_out/cfg4.json
