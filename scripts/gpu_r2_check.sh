# Round-2 validation of the rewritten tile kernel (K1b v2): parity first, then timing, then profiles.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_check.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled or kats or default_rules or chunking or k1c or strips or census or 4096" 2>&1 | tail -15
timeout 200 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench64.json 2> gpurun_out/r2_bench64.err; cat gpurun_out/r2_bench64.json | head -c 300; tail -3 gpurun_out/r2_bench64.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err; cat gpurun_out/r2_bench20.json | head -c 300; tail -3 gpurun_out/r2_bench20.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_r2_tiles python bench.py --steps 64 --warmup 8 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_tiles.err; tail -3 gpurun_out/ncu_tiles.err
