# Round-2 opener: validate and time the experimental, default-off paths in ONE gpurun call.
#   gpurun --timeout 600 -- 'bash scripts/gpu_experiments.sh'
mkdir -p gpurun_out
# 0. GPU tests written after round 1's GPU budget was spent (marker first_gpu_run: CUDA path vs reference-shader goldens,
#    long horizons at full size, the C++ mirror).  Once green: drop the marker from those files.
timeout 600 python -m pytest tests -q -m "gpu and first_gpu_run" 2>&1 | tail -5
# 1. running census (SE_FLAG_RUNNING_CENSUS): correctness, then the e2e leg with and without it
SE_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "running_census or lit_strips or fused_light or random_eligible or two_table or strips_with_modifications" 2>&1 | tail -5
timeout 200 python bench.py --steps 400 --warmup 16 --no-cpu-baseline > gpurun_out/exp_e2e_base.json 2> gpurun_out/exp_e2e_base.err
timeout 200 python bench.py --steps 400 --warmup 16 --no-cpu-baseline --running-census > gpurun_out/exp_e2e_running.json 2> gpurun_out/exp_e2e_running.err
python - <<'PY'
import json
for n in ("base", "running"):
    try:
        d = json.load(open(f"gpurun_out/exp_e2e_{n}.json"))
        print(n, "value", d["value"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(n, "failed:", e)
PY
SE_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_c_abi_example.py -q -m gpu 2>&1 | tail -3
# 2. lighting path after the se_light rewrite: launch list + one full capture of the new kernel
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/light_launches_r2.csv python scripts/light_probe.py 8192 12 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:se_light -s 16 -c 1 -o gpurun_out/prof_r2_light python scripts/light_probe.py 8192 12 > /dev/null 2>&1
python scripts/light_probe.py 8192 48
# 3. se_light tile-height / occupancy variants (rule-compile-time env; all three tile heights are host-checked)
#    ptxas on the build host: default 64 registers / 0 spills (also what MINCTAS=3 gives: identical code, dropped here);
#    ROWS=2: 56 registers, 10.8 KB smem; ROWS=8: 64 registers, 36.9 KB smem; MINCTAS=5: 48 registers with 24 B of spills
#    (worth one timing); MINCTAS >= 6 spills 200-450 B (not worth GPU time).
for v in "SE_LT_ROWS=2" "SE_LT_ROWS=8" "SE_LT_MINCTAS=5"; do echo "$v"; env $v python scripts/light_probe.py 8192 48; done
# 4. fused step + lighting kernel (K3f) against the two-kernel path
SE_FUSED=1 python scripts/light_probe.py 8192 48
SE_FUSED=1 python scripts/light_probe.py 4096 100
SE_FUSED=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:se_light_fused -s 16 -c 1 -o gpurun_out/prof_r2_light_fused python scripts/light_probe.py 8192 12 > /dev/null 2>&1
# 5. K1c occupancy (no code change: rule-compile-time defines + grid override).  Default: 54 registers, 2 CTAs of 512
#    threads per SM (50 %).  Checked with ptxas on the build host: MINCTAS=3 (<= 42 registers) and MINCTAS=4 with BATCH=2
#    (<= 32 registers) compile WITHOUT spills; MINCTAS=4 with BATCH=4 spills.  Staged table 43.6 KB per CTA fits 4 per SM.
python scripts/k1c_probe.py
SE_NVRTC_DEFS="-DSE_K1C_MINCTAS=3" SE_K1C_GRID=444 python scripts/k1c_probe.py
SE_NVRTC_DEFS="-DSE_K1C_MINCTAS=4 -DSE_K1C_BATCH=2" SE_K1C_GRID=592 python scripts/k1c_probe.py
SE_NVRTC_DEFS="-DSE_K1C_MINCTAS=3 -DSE_K1C_BATCH=2" SE_K1C_GRID=444 python scripts/k1c_probe.py
# 6. BASELINE configs[0] on the GPU (timing + comparison with the reference shader output)
python scripts/run_configs.py 1 > gpurun_out/config1.json
