"""Times the per-frame path (one se_sim_step(1) per frame) at 16384^2 on one GPU: the bare kernel K1c, K1c keeping the
running census, and the e2e loop of bench.py (push a modification record, step, asynchronous census read-back)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sandengine_b200 as se  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402

S, K = 16384, 200
rules = se.parse_path(REPO / "data" / "materials.yaml")
g = synthetic_grid(S, S, 3)
for running in (False, True):
    sim = se.Simulation(rules, (S, S), running_census=running)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
    sim.upload_cells(g); sim.params.frame = 1
    sim.step(64)
    if running:
        sim.census()
    for _ in range(8): sim.step(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(K): sim.step(1)
    e1.record(st); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3
    print(f"k1c_probe running_census={running}: kernel loop {t / K * 1e6:.1f} us/step -> {S * S * K / t / 1e9:.0f} Gcell/s")
    term = np.zeros(1, dtype=se.MOD_DTYPE)
    ring = torch.zeros((4, 256), dtype=torch.int64).pin_memory()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(K):
        sim.push_modifications(term); sim.step(1); sim.census_async(ring[k % 4].data_ptr())
        if k % 4 == 3: sim.census_wait()
    sim.census_wait(); torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"k1c_probe running_census={running}: e2e loop    {t / K * 1e6:.1f} us/step -> {S * S * K / t / 1e9:.0f} Gcell/s")
    sim.close()
