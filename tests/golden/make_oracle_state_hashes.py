"""Regenerates oracle_state_sha256.json: SHA-256 of oracle (C restatement) outputs for fixed seeded inputs.
These pin the ORACLE against drift (they are not reference outputs: the reference ships no state fixtures and
cannot run in this image -- SURVEY.md 8c); the survey's independent KATs are in tests/test_oracle_kat.py."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import yaml_cases as Y  # noqa: E402
from oracle.build_oracle import MOD_DTYPE, load_oracle  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402
from sandengine_b200.synth_rules import synthetic_rule_set  # noqa: E402


def cases():
    default = (REPO / "data" / "materials.yaml").read_text()
    s_text, s_ids, s_mix = synthetic_rule_set(64, 28, seed=5)
    yield "default_256x256_seed1_1000steps", default, synthetic_grid(256, 256, 1), 1000, {}
    yield "default_130x66_seed11_257steps", default, synthetic_grid(130, 66, 11), 257, {}
    yield "rich_128x96_seed5_300steps", Y.RICH_YAML, synthetic_grid(128, 96, 5, mix=Y.RICH_MIX, ids=Y.RICH_IDS), 300, {}
    yield "expr_96x64_seed5_120steps", Y.EXPR_YAML, synthetic_grid(96, 64, 5, mix=Y.EXPR_MIX, ids=Y.EXPR_IDS), 120, {}
    yield "synth64_160x128_seed5_200steps", s_text, synthetic_grid(160, 128, 5, mix=s_mix, ids=s_ids), 200, {}


def compute():
    out = {}
    for name, text, g, steps, _ in cases():
        c, _, _ = load_oracle(text).run(g, 1, steps, blocks=True)
        out[name] = {"init": hashlib.sha256(g.astype(np.uint32).tobytes()).hexdigest(), "final": hashlib.sha256(c.astype(np.uint32).tobytes()).hexdigest(),
                     "histogram": np.bincount(c.ravel()).tolist()}
    # lighting + modifications (literal per-cell form)
    g = synthetic_grid(64, 48, 9)
    mods = []
    for s in range(30):
        m = np.zeros(2, MOD_DTYPE)
        m[0] = ((7 * s % 64, 5 * s % 48), s % 2, 2 + s % 5, 3 + s % 8, (0, 0, 0))
        m[1] = ((63 - 3 * s % 64, 11 * s % 48), 0, 3, 0, (0, 0, 0))
        mods.append(m)
    c, L, _ = load_oracle((REPO / "data" / "materials.yaml").read_text()).run(g, 1, 30, light=np.zeros((48, 64, 4), np.float32), mods_per_step=mods)
    out["default_64x48_seed9_light_mods_30steps"] = {"final": hashlib.sha256(c.astype(np.uint32).tobytes()).hexdigest(),
                                                      "light_sum": [round(float(x), 3) for x in L.sum(axis=(0, 1))]}
    return out


if __name__ == "__main__":
    (Path(__file__).parent / "oracle_state_sha256.json").write_text(json.dumps(compute(), indent=1) + "\n")
