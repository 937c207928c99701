"""The CUDA path at the full sizes / long horizons of BASELINE configs[1] and configs[2] against committed oracle hashes
(tests/golden/oracle_long_horizon.json, generator make_oracle_long_horizon.py).  The oracle is pinned bit for bit by the
reference's own shader (tests/test_ref_shader.py); only the product runs here.
  4096^2, seed 2: 1000 steps in one se_sim_step call (K1b, 125 temporal blocks) and in chunks of 1 (K1c) + 7 + 64 ...
  16384^2, seed 3: 64 steps (K1b, 8 temporal blocks over the full grid)
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from sandengine_b200.grids import synthetic_grid

pytestmark = [pytest.mark.gpu, pytest.mark.first_gpu_run]

GOLDEN = json.loads((Path(__file__).parent / "golden" / "oracle_long_horizon.json").read_text())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def se(native_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sandengine_b200
    return sandengine_b200


@pytest.mark.parametrize("name,chunks", [("default_4096x4096_seed2_1000steps", None),
                                         ("default_4096x4096_seed2_1000steps", (1, 7, 64, 200, 1, 3, 724)),
                                         ("default_16384x16384_seed3_64steps", None)])
def test_cuda_path_matches_the_oracle_at_full_size(se, default_rules, name, chunks):
    want = GOLDEN[name]
    size, steps = want["size"], want["steps"]
    g = synthetic_grid(size, size, want["seed"])
    assert sha(g) == want["init"], "input generator drifted"
    sim = se.Simulation(default_rules, (size, size))
    sim.upload_cells(g)
    sim.params.frame = 1
    for n in (chunks or (steps,)):
        sim.step(n)
    assert sum(chunks or (steps,)) == steps and sim.params.frame == want["final_frame"]
    out = sim.download_cells()
    sim.close()
    assert np.bincount(out.ravel(), minlength=11).tolist() == want["histogram"]
    assert sha(out) == want["final"], f"{name}: cell ids differ from the oracle after {steps} steps"
