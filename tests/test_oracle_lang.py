"""Pins the oracle's Python restatement of sandengine-lang (oracle/oracle_lang.py) -- CPU only."""
import hashlib
import json
from pathlib import Path

import pytest

from conftest import DEFAULT_YAML, REFERENCE
from oracle import oracle_lang as L
import yaml_cases as Y

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_gen_sha256.json").read_text())


def _sha(s: str):
    b = s.encode()
    return {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}


def test_glsl_emitter_matches_reference_goldens():
    """The only golden fixtures the reference holds for this path: its checked-in gen/*.glsl."""
    res = L.parse_path(DEFAULT_YAML)
    assert _sha(L.emit_glsl_materials(res)) == GOLD["shaders/compute/gen/materials.glsl"]
    assert _sha(L.emit_glsl_rules(res)) == GOLD["shaders/compute/gen/rules.glsl"]


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree only exists in the build container")
def test_against_reference_tree_bytes():
    for f, meta in GOLD.items():
        if f.startswith("_"):
            continue
        b = (REFERENCE / f).read_bytes()
        assert {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()} == meta
    ref = L.parse_path(REFERENCE / "data" / "materials.yaml")
    own = L.parse_path(DEFAULT_YAML)
    assert L.emit_glsl_materials(ref) == (REFERENCE / "shaders/compute/gen/materials.glsl").read_text()
    assert L.emit_glsl_rules(ref) == (REFERENCE / "shaders/compute/gen/rules.glsl").read_text()
    # our re-typed data/materials.yaml is the same rule set
    assert L.emit_c_rules(ref) == L.emit_c_rules(own)


def test_default_rule_set_tables():
    res = L.parse_path(DEFAULT_YAML)
    assert [m.name for m in res.materials] == ["EMPTY", "NULL", "WALL", "sand", "rock", "water", "radioactive", "smoke",
                                               "toxic_sludge", "vine", "dirt"]
    assert [t.name for t in res.types] == ["EMPTY", "NULL", "WALL", "solid", "movable_solid", "liquid", "gas", "plant"]
    assert res.types[3].children == ["movable_solid"]
    used = [r.name for r in res.rules if r.used]
    assert used == ["fall_slide", "fall_slide_dirt", "horizontal_slide", "rise_up", "dissolve", "grow", "grow_up", "die_off"]
    assert all(r.ruletype == "Mirrored" for r in res.rules)
    by = {r.name: r for r in res.rules}
    assert by["fall_slide"].precondition == "isType_liquid(self) || self.mat == MAT_sand"
    assert by["rise_up"].precondition is None
    assert not res.materials[9].selectable and res.materials[3].selectable


def test_f32_display():
    for v, s in [(1.0, "1"), (1.5, "1.5"), (9999.0, "9999"), (87 / 255, "0.34117648"), (0.1, "0.1"), (0.004, "0.004"),
                 (0.99999, "0.99999"), (0.0, "0")]:
        import numpy as np
        assert L.f32_display(np.float32(v)) == s


@pytest.mark.parametrize("name,text,kind", Y.ERROR_CASES, ids=[c[0] for c in Y.ERROR_CASES])
def test_error_classes(name, text, kind):
    with pytest.raises(L.ParsingErr) as ei:
        L.parse_string(text)
    assert ei.value.kind == kind
    assert kind in str(ei.value)


def test_rich_rule_set_parses():
    res = L.parse_string(Y.RICH_YAML)
    kinds = {r.name: r.effective_type for r in res.rules}
    assert kinds["drift_left"] == "Left" and kinds["creep_left"] == "Left" and kinds["drift_right"] == "Right"
    assert kinds["sink"] == "Mirrored"
    # reference classification is dead code: every non-mirrored rule is "Right" 
    assert all(r.ruletype in ("Mirrored", "Right") for r in res.rules)
    # inheritance: granular's child fine_granular is registered in static's and granular's children
    t = {x.name: x for x in res.types}
    assert t["static"].children == ["granular", "fine_granular"] and t["granular"].children == ["fine_granular"]
