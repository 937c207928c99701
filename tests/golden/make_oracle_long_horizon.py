"""Regenerates oracle_long_horizon.json: SHA-256 of the ORACLE's grid at the long horizons / full sizes SURVEY.md 8d names,
which the reference's shader compiled for the CPU cannot reach in reasonable time (it runs at ~2 Mcell/s):
  configs[1]  4096 x 4096, seed 2, after 1000 steps          (the shader itself pins steps 1-5 and 100: ref_shader_goldens.json)
  configs[2]  16384 x 16384, seed 3, after 64 steps           (8 temporal blocks of the tile kernel over the full grid)
The oracle is pinned bit for bit by that shader on 19 cases (tests/test_ref_shader.py), so these are oracle outputs, labelled
as such.  The CPU suite does not recompute them (minutes of CPU); tests/test_gpu_long_horizon.py compares the CUDA path.
  python tests/golden/make_oracle_long_horizon.py        (~5 minutes on 8 cores, ~3 GiB)
"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))
from oracle.build_oracle import load_oracle  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402

CASES = [("default_4096x4096_seed2_1000steps", 4096, 2, 1000), ("default_16384x16384_seed3_64steps", 16384, 3, 64)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


if __name__ == "__main__":
    orc = load_oracle()
    out = {}
    for name, size, seed, steps in CASES:
        t = time.time()
        g = synthetic_grid(size, size, seed)
        init = sha(g)
        frame = orc.run_blocks(g, 1, steps)          # in place, lighting off, no modifications
        out[name] = {"size": size, "seed": seed, "steps": steps, "init": init, "final": sha(g), "final_frame": frame,
                     "histogram": np.bincount(g.ravel(), minlength=11).tolist()}
        print(name, f"{time.time() - t:.0f} s", flush=True)
    (Path(__file__).parent / "oracle_long_horizon.json").write_text(json.dumps(out, indent=1) + "\n")
