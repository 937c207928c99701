"""Times K1c (one se_sim_step(1) per frame) at 16384^2 on one GPU."""
import sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sandengine_b200 as se
from sandengine_b200.grids import synthetic_grid
S, K = 16384, 200
rules = se.parse_path(REPO / "data" / "materials.yaml")
sim = se.Simulation(rules, (S, S))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
sim.upload_cells(synthetic_grid(S, S, 3)); sim.params.frame = 1
sim.step(64)
for _ in range(8): sim.step(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(st)
for _ in range(K): sim.step(1)
e1.record(st); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 1e3
print(f"k1c_probe: {t / K * 1e6:.1f} us/step -> {S * S * K / t / 1e9:.0f} Gcell/s")
