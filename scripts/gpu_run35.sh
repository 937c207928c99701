mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 32 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:se_step_tiles -s 1 -c 1 -o gpurun_out/prof_r1_k1b_final python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_k1b.log 2>&1
timeout 300 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; cat gpurun_out/bench_final_n1.json | cut -c1-400
