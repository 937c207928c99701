"""Triangulates the C oracle (generated C rule text) with the independent pure-Python restatement
(oracle/pyoracle.py: eval of the reference-shaped GLSL conditions) -- small grids, CPU only."""
import hashlib

import numpy as np
import pytest

import yaml_cases as Y
from conftest import DEFAULT_YAML
from oracle.build_oracle import load_oracle
from oracle.pyoracle import PyOracle, hash43
from sandengine_b200.grids import kat_grid, synthetic_grid


def test_hash_kat_pure_python():
    r = hash43(0, 0, 2)
    assert abs(float(r[0]) - 0.9144064784) < 1e-9 and abs(float(r[1]) - 0.8206900358) < 1e-9
    assert bool(hash43(2, 0, 2)[0] < np.float32(0.5)) and not bool(hash43(-1, -1, 5)[0] < np.float32(0.5))


def test_default_rules_state_kat_pure_python():
    py = PyOracle(DEFAULT_YAML.read_text())
    c, f = py.run(kat_grid(16), 1, 40)
    assert f == 41 and hashlib.sha256(c.astype(np.uint8).tobytes()).hexdigest()[:16] == "e4ab9d8c55d01623"


@pytest.mark.parametrize("name", ["default", "rich", "synth64"])
def test_c_oracle_equals_python_oracle(name):
    if name == "default":
        text, kw = DEFAULT_YAML.read_text(), {}
    elif name == "rich":
        text, kw = Y.RICH_YAML, dict(mix=Y.RICH_MIX, ids=Y.RICH_IDS)
    else:
        from sandengine_b200.synth_rules import synthetic_rule_set
        text, ids, mix = synthetic_rule_set(64, 28, seed=5)
        kw = dict(mix=mix, ids=ids)
    c_orc, py = load_oracle(text), PyOracle(text)
    for (w, h, seed, steps, f0) in [(24, 18, 3, 24, 1), (9, 7, 4, 12, 1002)]:
        g = synthetic_grid(w, h, seed, **kw)
        a, _, _ = c_orc.run(g, f0, steps, blocks=True)
        b, _ = py.run(g, f0, steps)
        assert np.array_equal(a, b), name
        assert not np.array_equal(a, g)
