set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/env.txt; nproc >> gpurun_out/env.txt; ldconfig -p | grep -E 'EGL|OSMesa|GLX|libGL' >> gpurun_out/env.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
