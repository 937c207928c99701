"""Generator of tests/golden/bench_checksums.json: the ORACLE's full-grid checksum and census after n steps of
bench.py's workload (counter-hash grid seed 3, default rule set, lighting off, frame 1 start), for the step totals
bench.py reaches with the driver's flags (--warmup 5 --steps 20 --reps 5 -> 105 steps) and a few cheaper ones.
bench.py compares the sharding-independent checksum of the GPU run (se_sim_checksum, summed over ranks) with these.

  python tests/golden/make_bench_checksums.py [SIZE:STEPS ...]      (default: the list below; 16384^2 takes minutes)
"""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from oracle.build_oracle import load_oracle  # noqa: E402
from sandengine_b200.grids import grid_checksum, synthetic_grid  # noqa: E402

SEED = 3
DEFAULT = ["256:105", "1024:105", "2048:105", "16384:25", "16384:72", "16384:105"]


def main():
    out_path = Path(__file__).with_name("bench_checksums.json")
    table = json.loads(out_path.read_text()) if out_path.exists() else {}
    orc = load_oracle()
    todo = {}
    for spec in (sys.argv[1:] or DEFAULT):
        size, steps = (int(x) for x in spec.split(":"))
        todo.setdefault(size, []).append(steps)
    for size, steps_list in todo.items():
        g = synthetic_grid(size, size, SEED)
        frame, done = 1, 0
        for steps in sorted(steps_list):
            frame = orc.run_blocks(g, frame, steps - done)
            done = steps
            table[f"{size}x{size}:seed{SEED}:steps{steps}"] = {
                "checksum": grid_checksum(g), "census": [int(x) for x in np.bincount(np.minimum(g, 255).ravel(), minlength=256)[:11]]}
            out_path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
            print(size, steps, f"{table[f'{size}x{size}:seed{SEED}:steps{steps}']['checksum']:016x}", flush=True)


if __name__ == "__main__":
    main()
