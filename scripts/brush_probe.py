"""Per-frame path with a brush held down, lighting off: every frame pushes 4 stamps + 1 explosion (the generator of BASELINE configs[3]) and
runs ONE step -- se_step_lut_global[_census]_mods.  Prints CUDA-event time per frame with and without the running census, and the same loop
without modifications for comparison.
  python scripts/brush_probe.py [S] [frames]          (env SE_TEMPORAL_BLOCK=1: the generated-code kernel se_step_inplace_mods instead)"""
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

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rules = se.parse_path(REPO / "data" / "materials.yaml")
sel = [m.id for m in rules.selectable_materials]
g = synthetic_grid(S, S, 3)
mods = [frame_mods(k, S, S, sel) for k in range(K + 8)]
out = {"S": S, "frames": K}
for label, kw, with_mods in [("brush", {}, True), ("brush_running_census", {"running_census": True}, True), ("no_brush", {}, False),
                             ("brush_generated_code", {"temporal_block": 1}, True)]:
    sim = se.Simulation(rules, (S, S), **kw)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
    sim.upload_cells(g); sim.params.frame = 1
    for k in range(8):
        if with_mods: sim.push_modifications(mods[k])
        sim.run()
    if kw.get("running_census"): sim.census()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for k in range(8, K + 8):
        if with_mods: sim.push_modifications(mods[k])
        sim.run()
    e1.record(st); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3
    out[label] = {"us_per_frame": round(t / K * 1e6, 1), "gcell_per_s": round(S * S * K / t / 1e9, 1)}
    sim.close()
print(json.dumps(out))
