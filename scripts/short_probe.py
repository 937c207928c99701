"""Fixed cost of one K1b launch on a strip-sized grid: time step(n) for n = 20, 40, 80 on W x H (default 16384 x 2048 = one of eight
strips of the bench grid); 2 t(20) - t(40) is what a launch costs beyond its steps.
  python scripts/short_probe.py [W] [H] [temporal_block]          (env SE_TILE_PH: tile height override)"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sandengine_b200 as se  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
T = int(sys.argv[3]) if len(sys.argv) > 3 else 0        # temporal_block (0: the library's default)
rules = se.parse_path(REPO / "data" / "materials.yaml")
sim = se.Simulation(rules, (W, H), temporal_block=T)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
sim.upload_cells(synthetic_grid(W, H, 3)); sim.params.frame = 1
sim.step(5)
out = {"W": W, "H": H, "T": T, "SE_TILE_PH": __import__("os").environ.get("SE_TILE_PH")}
for n in (10, 20, 40, 80, 20):
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(st); sim.step(n); e1.record(st); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out[f"ms_step{n}"] = round(ts[len(ts) // 2], 4)
out["fixed_ms"] = round(2 * out["ms_step20"] - out["ms_step40"], 4)
out["gcell_per_s_20"] = round(W * H * 20 / out["ms_step20"] / 1e6, 1)
out["gcell_per_s_marginal"] = round(W * H * 40 / (out["ms_step80"] - out["ms_step40"]) / 1e6, 1)
print(json.dumps(out))
