"""Times ONE strip of an N-way split on one GPU without neighbours (ghost rows go stale -> results are not
meaningful, only the per-rank kernel time is): cheap stand-in for the compute part of an N-GPU run.
  python scripts/strip_probe.py N [halo] [T]"""
import sys
from pathlib import Path
import numpy as np
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sandengine_b200 as se
from sandengine_b200.distributed import StripPlan
from sandengine_b200.grids import synthetic_grid

N = int(sys.argv[1]); halo = int(sys.argv[2]) if len(sys.argv) > 2 else 34; T = int(sys.argv[3]) if len(sys.argv) > 3 else 0
S, K = 16384, 512
plan = StripPlan(S, S, N, halo if N > 1 else 0)
r = N // 2
b, e = plan.rows(r)
rules = se.parse_path(REPO / "data" / "materials.yaml")
sim = se.Simulation(rules, (S, S), row_begin=b, row_end=e, halo_rows=plan.halo_rows, temporal_block=T)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
sim.upload_cells(synthetic_grid(S, S, 3, row_begin=b, row_end=e)); sim.params.frame = 1
sim.step(64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(st); sim.step(K); e1.record(st); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 1e3
print(f"strip_probe N={N} halo={halo} T={T or 8}: rows {b}..{e} (+ghosts {plan.ghosts(r)}): {t / K * 1e6:.2f} us/step -> {N}x-equivalent {S * S * K / t / 1e9:.0f} Gcell/s if perfectly coupled")
