"""Pins the C restatement of the compute shader (oracle/sand_oracle.c) against the cross-check vectors of
SURVEY.md section 8c -- the reference itself holds no state-level fixtures.  CPU only."""
import hashlib

import numpy as np
import pytest

from sandengine_b200.grids import kat_grid, synthetic_grid
from oracle.build_oracle import MOD_DTYPE


def sha16(g):
    return hashlib.sha256(g.astype(np.uint8).tobytes()).hexdigest()[:16]


HASH_KATS = [  # (px, py, frame) -> seed, lane0, lane1, mirror
    ((0, 0, 2), 0x008A95D2, 0xEA168B23, 0xD218BE2F, False),
    ((2, 0, 2), 0x008A996C, 0x584201C0, 0x5BD60E64, True),
    ((-1, -1, 5), 0x015A6C6D, 0xDB2CF413, 0xFBF8349D, False),
    ((0, -1, 2), 0x008A8D7F, 0x9DA58A51, 0x903F1D5A, False),
    ((255, 255, 1000), 0x0EB6B408, 0x232B76F2, 0x88E4ECB4, True),
    ((16382, 16382, 1001), 0x1179D4D1, 0x643CDF59, 0x3FC0D005, True),
    ((65535, 65535, 4), 0x0B352184, 0xC412C61D, 0x95B473D7, False),
]


@pytest.mark.parametrize("p,seed,l0,l1,mirror", HASH_KATS)
def test_hash43_kat(oracle, p, seed, l0, l1, mirror):
    s, lanes, r = oracle.hash43(*p)
    assert (s, lanes[0], lanes[1]) == (seed, l0, l1)
    assert bool(r[0] < np.float32(0.5)) == mirror


def test_hash_values(oracle):
    _, lanes, r = oracle.hash43(0, 0, 2)
    assert lanes == [0xEA168B23, 0xD218BE2F, 0x512A0588, 0x1CA4752E]
    assert abs(float(r[0]) - 0.9144064784) < 1e-9 and abs(float(r[1]) - 0.8206900358) < 1e-9


def test_mirror_boundary():
    # rand.x < 0.5  <=>  u0 <= 2^31 - 65 (u0 = 2^31 - 64 ties-to-even up to 2^31)
    f = lambda u: np.float32(u) / np.float32(4294967296.0)
    assert f(2**31 - 65) < np.float32(0.5) and not (f(2**31 - 64) < np.float32(0.5))
    assert np.float32(0.1).view(np.uint32) == 0x3DCCCCCD and np.float32(0.004).view(np.uint32) == 0x3B83126F
    assert np.float32(0.001).view(np.uint32) == 0x3A83126F and np.float32(0.3).view(np.uint32) == 0x3E99999A


STATE_KATS = [(16, 40, "9d3592bdc9197ce4", "e4ab9d8c55d01623", [99, 0, 0, 32, 31, 30, 0, 26, 0, 1, 37]),
              (32, 100, "fb250fd4366337af", "d2f2ab20d7d3a8bc", [400, 0, 0, 133, 128, 126, 0, 109, 0, 0, 128]),
              (20, 60, None, "298014800d6a9b1b", [168, 0, 0, 51, 44, 46, 0, 42, 0, 0, 49])]


@pytest.mark.parametrize("n,steps,init_sha,final_sha,hist", STATE_KATS)
def test_state_kat(oracle, n, steps, init_sha, final_sha, hist):
    g = kat_grid(n)
    if init_sha:
        assert sha16(g) == init_sha
    for blocks in (False, True):   # literal per-cell form and per-block in-place form
        c, _, frame = oracle.run(g, 1, steps, blocks=blocks)
        assert frame == 1 + steps
        assert sha16(c) == final_sha
        assert np.bincount(c.ravel(), minlength=11).tolist() == hist


def test_light_kat(oracle):
    g = kat_grid(16)
    g[5][5] = 6
    g[9][12] = 8
    assert sha16(g) == "b9e9a99a089ece58"
    c, L, _ = oracle.run(g, 1, 40, light=np.zeros((16, 16, 4), np.float32))
    assert sha16(c) == "ca0e8b7d36edd6b2"
    np.testing.assert_allclose(L.sum(axis=(0, 1)), [37.89108, 39.94216, 34.04128, 159.99199], rtol=2e-6)
    probes = {(0, 0): (1, 1, 1, 0.999999), (3, 1): (0.3963899, 0.3967239, 0.3963899, 0.9295824), (5, 5): (0.05, 0.7, 0.05, 0.9),
              (8, 8): (0.01085952, 0.06828480, 0.01085950, 0.5489540), (15, 15): (0, 0, 0, 0.6501103)}
    for (x, y), v in probes.items():
        np.testing.assert_allclose(L[y][x], v, rtol=2e-6, atol=1e-7)


def test_cells_form_equals_blocks_form_random(oracle):
    for (w, h, seed) in [(64, 48, 1), (33, 21, 7), (1, 9, 3), (9, 1, 4), (2, 2, 5)]:
        g = synthetic_grid(w, h, seed)
        a, _, _ = oracle.run(g, 1, 30)
        b, _, _ = oracle.run(g, 1, 30, blocks=True)
        assert np.array_equal(a, b)


def test_frame1_clears_and_ignores_mods(oracle):
    g = synthetic_grid(16, 16, 1)
    mods = np.zeros(1, MOD_DTYPE)
    mods[0]["position"] = (8, 8); mods[0]["mod_size"] = 3; mods[0]["mod_matID"] = 3
    c, _, f = oracle.run(g, 0, 1, mods_per_step=[mods])
    assert f == 1 and not c.any()


def test_modification_semantics(oracle):
    g = np.zeros((24, 24), np.uint32)
    mods = np.zeros(5, MOD_DTYPE)
    mods[0] = ((5, 5), 0, 3, 3, (0, 0, 0))      # circle of sand
    mods[1] = ((15, 5), 1, 2, 4, (0, 0, 0))     # square of rock
    mods[2] = ((5, 5), 0, 1, 99, (0, 0, 0))     # unknown id => NULL => last match wins but cancels the override
    mods[3] = ((20, 20), 0, 0, 5, (0, 0, 0))    # mod_size == 0 terminates the scan
    mods[4] = ((20, 20), 1, 2, 5, (0, 0, 0))    # never reached
    c, _, _ = oracle.run(g, 1, 1, mods_per_step=[mods])
    assert c[5, 5] == 0 and c[5, 4] == 0 and c[5, 8] == 3 and c[8, 5] == 3 and c[5, 9] == 0   # centre cancelled by the NULL mod
    assert c[2, 5] == 3 and c[3, 3] == 3 and c[2, 3] == 0           # circle radius 3: (2,2) offset has dist 2.83, (3,2) 3.6
    assert (c[3:8, 13:18] == 4).all() and c[2, 15] == 0
    assert c[20, 20] == 0
    # explosion = circle of EMPTY (id 0 passes the != NULL test)
    g2 = np.full((16, 16), 4, np.uint32)
    m2 = np.zeros(1, MOD_DTYPE); m2[0] = ((8, 8), 0, 2, 0, (0, 0, 0))
    c2, _, _ = oracle.run(g2, 1, 1, mods_per_step=[m2])
    assert c2[8, 8] == 0 and c2[8, 10] == 0 and c2[8, 11] == 4 and (c2 == 0).sum() == 13


def test_unknown_ids_become_null(oracle):
    g = np.zeros((4, 4), np.uint32)
    g[1, 1] = 200
    c, _, _ = oracle.run(g, 1, 1)
    assert c[1, 1] == 1


def test_oracle_state_fixtures_do_not_drift():
    """tests/golden/oracle_state_sha256.json (made by make_oracle_state_hashes.py): the oracle still produces them."""
    import importlib.util
    import json
    from pathlib import Path
    here = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("make_oracle_state_hashes", here / "make_oracle_state_hashes.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.loads((here / "oracle_state_sha256.json").read_text())
    got = mod.compute()
    assert set(got) == set(want)
    for k in want:
        for field in ("init", "final", "histogram"):
            if field in want[k]:
                assert got[k][field] == want[k][field], (k, field)
        if "light_sum" in want[k]:
            np.testing.assert_allclose(got[k]["light_sum"], want[k]["light_sum"], rtol=1e-5)
