#!/bin/bash
# the whole GPU suite + smoke() on one B200 (gpurun -- 'bash scripts/gpu_final_check.sh')
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
