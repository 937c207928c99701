#!/bin/bash
# two GPUs: multi-process strip parity (incl. the strip snapshot case); one of them: the newest GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29510 scripts/check_strips_multi.py 2>&1 | grep check_strips
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "brush or snapshot" 2>&1 | tail -3
