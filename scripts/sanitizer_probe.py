"""Small runs of the round-2 kernels for compute-sanitizer (memcheck / racecheck), each compared with the oracle:
  compute-sanitizer --tool racecheck python scripts/sanitizer_probe.py
lit path with modifications on a grid with interior AND edge tiles (se_step_lit), per-frame steps with modifications and the running
census (se_step_lut_global_census[_mods]), a run of steps (se_step_tiles)."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "scripts"))
import sandengine_b200 as se  # noqa: E402
from oracle.build_oracle import load_oracle  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402
from run_configs import frame_mods  # noqa: E402

rules = se.parse_path(REPO / "data" / "materials.yaml")
sel = [m.id for m in rules.selectable_materials]
orc = load_oracle()
# lit: 448 x 200 -> 7 x 13 tiles of 64 x 16, interior ones included
w, h, frames = 448, 200, 6
g = synthetic_grid(w, h, 71)
L0 = np.random.default_rng(5).random((h, w, 4), dtype=np.float32)
mods = [frame_mods(k, w, h, sel) for k in range(frames)]
sim = se.Simulation(rules, (w, h), lighting=True)
sim.upload_cells(g); sim.upload_light(L0); sim.params.frame = 1
for k in range(frames):
    sim.push_modifications(mods[k]); sim.run()
ref, refL, _ = orc.run(g, 1, frames, light=L0, mods_per_step=mods)
assert np.array_equal(sim.download_cells(), ref) and float(np.abs(sim.download_light() - refL).max()) <= 1e-6
sim.close()
print("lit ok")
# per-frame table kernel with modifications and the running census, then a run of steps
w, h, frames = 516, 96, 6
g = synthetic_grid(w, h, 72)
mods = [frame_mods(k, w, h, sel) for k in range(frames)]
sim = se.Simulation(rules, (w, h), running_census=True)
sim.upload_cells(g); sim.params.frame = 1
sim.run(); sim.census()
for k in range(frames):
    sim.push_modifications(mods[k]); sim.run()
ref, _, frame = orc.run(g, 1, 1)
ref, _, frame = orc.run(ref, frame, frames, mods_per_step=mods)
assert np.array_equal(sim.download_cells(), ref)
assert np.array_equal(sim.census(), np.bincount(np.minimum(ref, 255).ravel(), minlength=256))
sim.step(20)
ref, _, frame = orc.run(ref, frame, 20, blocks=True)
assert np.array_equal(sim.download_cells(), ref)
sim.close()
print("per-frame + tiles ok")
