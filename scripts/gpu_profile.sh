# The ncu captures summarised under profiles/ (run under gpurun on ONE B200; never wrap a multi-rank command in ncu).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh'
# then summarise gpurun_out/*.ncu-rep / *.csv with `ncu -i ... --page raw --csv` (see profiles/*.txt headers).
set -x
mkdir -p gpurun_out
# every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 32 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# K1b: the second se_step_tiles launch = the timed 64-step launch (8 T-blocks of 8 Margolus steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_r1_k1b_final \
    python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_k1b.log 2>&1
# K1c: a per-frame launch of the e2e leg
timeout 900 ncu --set full --clock-control none --import-source on -k regex:se_step_lut_global -s 4 -c 1 -o gpurun_out/prof_r1_k1c_final \
    python bench.py --steps 16 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_k1c.log 2>&1
# K1a (generic generated-code kernel): force it with --temporal-block 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:se_step_inplace -s 5 -c 2 -o gpurun_out/prof_r1_k1a \
    python bench.py --steps 5 --warmup 3 --temporal-block 1 --no-cpu-baseline > gpurun_out/ncu_k1a.log 2>&1
