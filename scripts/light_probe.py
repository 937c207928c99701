"""Per-frame path with lighting + modifications (BASELINE configs[3] shape), short: for ncu launch lists / captures.
  python scripts/light_probe.py [S] [frames]          (env SE_NO_FUSED_LIT=1: the two-kernel path instead of se_step_lit)
Prints CUDA-event time per frame; under ncu the kernel of interest is se_step_lit (two-kernel path: se_step_pingpong_mods, se_light).
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "scripts"))
import sandengine_b200 as se  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402
from run_configs import frame_mods  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
FUSED = os.environ.get("SE_NO_FUSED_LIT") is None
rules = se.parse_path(REPO / "data" / "materials.yaml")
sel = [m.id for m in rules.selectable_materials]
sim = se.Simulation(rules, (S, S), lighting=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
sim.upload_cells(synthetic_grid(S, S, 4)); sim.upload_light(np.zeros((S, S, 4), np.float32)); sim.params.frame = 1
mods = [frame_mods(k, S, S, sel) for k in range(K + 8)]
for k in range(8):
    sim.push_modifications(mods[k]); sim.run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(st)
for k in range(8, K + 8):
    sim.push_modifications(mods[k]); sim.run()
e1.record(st); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 1e3
print(json.dumps({"S": S, "frames": K, "fused": FUSED, "ms_per_frame": round(t / K * 1e3, 4), "gcell_per_s": round(S * S * K / t / 1e9, 1),
                  "frac_40B": round(40.0 * S * S * K / t / 1e9 / 6537.3, 4)}))
