"""Tile-geometry sweep of K1b on W x H: time step(n) over T-block lengths and tile heights (SE_TILE_PH is read at every launch).
One JSON line per (H, n).
  python scripts/geom_sweep.py W n H [H ...]         (n = steps per launch: 20 = the driver's flags)"""
import json
import os
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sandengine_b200 as se  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402

W, N = int(sys.argv[1]), int(sys.argv[2])
rules = se.parse_path(REPO / "data" / "materials.yaml")
for H in [int(a) for a in sys.argv[3:]]:
    g = synthetic_grid(W, H, 3)
    res = {}
    for T in ((7, 10) if N <= 20 else (8, 10, 12, 16)):
        sim = se.Simulation(rules, (W, H), temporal_block=T)
        st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
        sim.upload_cells(g); sim.params.frame = 1
        sim.step(5)
        nblk = (N + T - 1) // T
        ts_eff = (N + nblk - 1) // nblk
        hy = ((ts_eff // 2 + 1) + 1) & ~1
        for ph in (0, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256, 282):
            if ph:
                if ph - 2 * hy < 2 * ts_eff + 8: continue
                os.environ["SE_TILE_PH"] = str(ph)
            else:
                os.environ.pop("SE_TILE_PH", None)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(st); sim.step(N); e1.record(st); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res[f"T{T}_ph{ph}"] = round(ts[2], 4)
        os.environ.pop("SE_TILE_PH", None)
        sim.close()
    best = min(res, key=res.get)
    print(json.dumps({"W": W, "H": H, "steps": N, "best": best, "best_ms": res[best], "gcell_per_s_best": round(W * H * N / res[best] / 1e6, 1), "all": res}), flush=True)
