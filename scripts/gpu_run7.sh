set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_tiles_T8c python bench.py --steps 16 --warmup 8 --temporal-block 8 --no-cpu-baseline > gpurun_out/ncu_tiles.log 2>&1
tail -3 gpurun_out/ncu_tiles.log
