set -x
mkdir -p gpurun_out
for TT in 512 768 1024; do for T in 4 8; do
SE_TILE_THREADS=$TT timeout 600 python bench.py --steps 240 --warmup 24 --temporal-block $T --no-cpu-baseline > gpurun_out/bench9_t${TT}_T$T.json 2> gpurun_out/bench9.err; tail -2 gpurun_out/bench9.err; python -c "
import json; d=json.load(open('gpurun_out/bench9_t${TT}_T$T.json')); print('threads=$TT T=$T', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['job_roundtrip']['value'], d['gpu_launches'])"
done; done
SE_TILE_THREADS=1024 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or strips or default_rules" 2>&1 | tail -2
