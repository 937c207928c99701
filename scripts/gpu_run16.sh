set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench16.json 2> gpurun_out/bench16.err; tail -3 gpurun_out/bench16.err; cat gpurun_out/bench16.json
timeout 600 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench16_ref.json 2> gpurun_out/bench16_ref.err; cat gpurun_out/bench16_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 32 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_r1_k1b python bench.py --steps 16 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_k1b.log 2>&1
