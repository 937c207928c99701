mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_goldens.py -x -q -m gpu -k "light or lit or ref_goldens or modifications" 2>&1 | tail -5
python scripts/light_probe.py 8192 48
python scripts/light_probe.py 4096 100
SE_NO_FUSED_LIT=1 python scripts/light_probe.py 8192 48
SE_NO_FUSED_LIT=1 python scripts/light_probe.py 4096 100
timeout 200 ncu --set full --clock-control none --import-source on -k regex:se_step_lit -s 16 -c 1 -o gpurun_out/prof_r2_lit python scripts/light_probe.py 8192 12 > /dev/null 2>&1
python scripts/k1c_probe.py
