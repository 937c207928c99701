"""The reference's documented quirks (SURVEY.md Appendix B "README vs. code: trust the code" and section 8a P1-P4), each
pinned by name through BOTH front ends: the C++ one behind the ABI and the Python restatement (oracle_lang).  A front end
that "fixed" one of these would no longer be a drop-in: the same YAML would mean something else than in the reference."""
import numpy as np
import pytest

import yaml_cases as Y
from oracle import oracle_lang


@pytest.fixture(scope="module")
def se(native_lib):
    import sandengine_b200
    return sandengine_b200


def both(se, text):
    """-> (native ParsingResult, oracle ParsingResult); generated text must agree."""
    nat = se.parse_string(text, compile=False)
    orc = oracle_lang.parse_string(text)
    assert nat.glsl_rules == oracle_lang.emit_glsl_rules(orc) and nat.glsl_materials == oracle_lang.emit_glsl_materials(orc)
    return nat, orc


def both_fail(se, text, kind):
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(text, compile=False)
    assert ei.value.kind == kind, str(ei.value)
    with pytest.raises(oracle_lang.ParsingErr) as eo:
        oracle_lang.parse_string(text)
    assert eo.value.kind == kind, str(eo.value)


def test_function_syntax_swap_is_not_recognised(se):
    """README.md:201 shows `do: swap(SELF, DOWN)`; only `SWAP a b` / `SET a m` exist (rules.rs:364-417)."""
    both_fail(se, Y.BASE_OK.replace("do: SWAP SELF DOWN\n", "do: swap(SELF, DOWN)\n", 1), "NotRecognized")


def test_lowercase_actions_are_not_recognised(se):
    both_fail(se, Y.BASE_OK.replace("do: SWAP SELF DOWN\n", "do: swap SELF DOWN\n", 1), "NotRecognized")


def test_istype_empty_breaks_but_istype_EMPTY_works(se):
    """README.md:210 `isType_empty(SELF)`: the literal replace "empty" -> "MAT_EMPTY" (rules.rs:343) turns it into
    isType_MAT_EMPTY, which is not a type (rules.rs:262-272)."""
    both_fail(se, Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", "if: isType_empty(DOWN)", 1), "NotFound")
    nat, _ = both(se, Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", "if: isType_EMPTY(DOWN)", 1))
    assert "isType_EMPTY(down)" in nat.glsl_rules


def test_lowercase_empty_on_the_right_of_a_material_comparison_is_not_found(se):
    """`SELF.mat == empty`: "empty" becomes MAT_EMPTY first, and MAT_EMPTY is not a material name (rules.rs:246-260)."""
    both_fail(se, Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", "if: DOWN.mat == empty", 1), "NotFound")
    nat, _ = both(se, Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", "if: DOWN.mat == EMPTY", 1))
    assert "down.mat == MAT_EMPTY" in nat.glsl_rules


def test_chance_key_is_ignored_only_probability_counts(se):
    """README.md:181-182 uses `chance:`; the parser reads `probability` only (rules.rs:310)."""
    with_chance = Y.BASE_OK.replace("    mirrored: true\n", "    mirrored: true\n    chance: 0.25\n", 1)
    nat, _ = both(se, with_chance)
    assert "rand.y" not in nat.glsl_rules
    nat, _ = both(se, Y.BASE_OK.replace("    mirrored: true\n", "    mirrored: true\n    probability: 0.25\n", 1))
    assert "rand.y <= 0.25 && " in nat.glsl_rules          # `<=`, not `<` (README.md:172-174 vs rules.rs:60)


def test_probability_one_emits_no_rand_test(se):
    nat, _ = both(se, Y.BASE_OK.replace("    mirrored: true\n", "    mirrored: true\n    probability: 1.0\n", 1))
    assert "rand.y" not in nat.glsl_rules


def test_only_used_rules_are_emitted_in_yaml_order(se):
    """sandengine-lang/src/lib.rs:83-86: a rule nobody lists in base_rules / extra_rules is dropped."""
    text = Y.BASE_OK.replace("types:", "  never_used:\n    if: SELF.mat == sand\n    do: SWAP SELF RIGHT\ntypes:", 1)
    nat, _ = both(se, text)
    assert "rule_never_used" not in nat.glsl_rules
    assert nat.glsl_rules.index("rule_gravity") < nat.glsl_rules.index("rule_slide_diagonally")
    assert [r.name for r in nat.rules if not r.used] == ["never_used"]


def test_unknown_extra_rules_are_silently_ignored_but_unknown_base_rules_are_not(se):
    """materials.rs:147-165 vs types.rs:121-147."""
    both(se, Y.BASE_OK.replace("    selectable: true", "    selectable: true\n    extra_rules: [no_such_rule]"))
    both_fail(se, Y.BASE_OK.replace("base_rules: [gravity, slide_diagonally]", "base_rules: [gravity, no_such_rule]"), "NotFound")


def test_colour_components_integer_one_is_one_255th(se):
    """parser.rs:219-235: integers 1..255 and floats > 1 are divided by 255, floats in [0, 1] are taken as they are --
    so the integer 1 is 1/255 while 1.0 is full intensity; three components get the default alpha (1 for colour)."""
    nat, _ = both(se, Y.BASE_OK.replace("color: [1.0, 1.0, 0.0, 1.0]", "color: [1, 1.0, 128, 200.0]"))
    m = [m for m in nat.materials if m.name == "sand"][0]
    assert np.allclose(m.color, [1 / 255, 1.0, 128 / 255, 200 / 255], rtol=0, atol=1e-7)
    nat, _ = both(se, Y.BASE_OK.replace("color: [1.0, 1.0, 0.0, 1.0]", "color: [10, 20, 30]"))
    m = [m for m in nat.materials if m.name == "sand"][0]
    assert np.allclose(m.color, [10 / 255, 20 / 255, 30 / 255, 1.0], rtol=0, atol=1e-7)


def test_selectable_non_bool_silently_becomes_false(se):
    """materials.rs:123-131."""
    nat, orc = both(se, Y.BASE_OK.replace("selectable: true", "selectable: 1"))
    assert [m.selectable for m in nat.materials if m.name == "sand"] == [False]
    assert [m.selectable for m in orc.materials if m.name == "sand"] == [False]


def test_builtin_ids_and_densities(se):
    """EMPTY 0 (density 1), NULL 1 (density 0), WALL 2 (density 9999), user materials from 3 (materials.rs:52-83)."""
    nat, _ = both(se, Y.BASE_OK)
    assert [(m.id, m.name, m.density) for m in nat.materials] == [(0, "EMPTY", 1.0), (1, "NULL", 0.0), (2, "WALL", 9999.0), (3, "sand", 1.5)]


def test_non_mirrored_rules_are_all_classified_right_by_the_reference(se):
    """rules.rs:152-163: `do_actions[0].contains("LEFT")` runs on already lower-cased text, so the reference's own
    classification never yields Left; the known-answer GLSL keeps that, the CUDA back end uses the LEFT definition."""
    text = Y.BASE_OK.replace("do: SWAP SELF DOWN\n    mirrored: false", "do: SWAP SELF LEFT\n    mirrored: false")
    orc = oracle_lang.parse_string(text)
    r = [r for r in orc.rules if r.name == "gravity"][0]
    assert r.ruletype == "Right" and r.effective_type == "Left"
    nat = se.parse_string(text, compile=False)
    assert nat.glsl_rules == oracle_lang.emit_glsl_rules(orc)
    assert "swap(self, left);" in nat.glsl_rules and "inout Cell left" not in nat.glsl_rules      # the uncompilable text of the reference


def test_duplicate_mapping_keys_are_a_yaml_error(se):
    """serde_yaml 0.9 (unsafe-libyaml 0.2.11) refuses a repeated key while building the Value that parser.rs:95-98
    deserialises into ("duplicate entry with key ..."); PyYAML would keep the last one.  Found by scripts/diff_frontends.py."""
    both_fail(se, Y.BASE_OK.replace("    do: SWAP SELF DOWN\n", "    do: SWAP SELF DOWN\n    do: SWAP SELF RIGHT\n", 1), "Yaml")
    both_fail(se, Y.BASE_OK + "types:\n  other:\n    base_rules: [gravity]\n", "Yaml")


def test_wrapped_plain_scalars_are_folded(se):
    """A long condition wrapped over several deeper-indented lines is ONE plain scalar in YAML (line breaks fold into
    spaces); both front ends must read it like libyaml does.  Found by scripts/diff_frontends.py."""
    wrapped = Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density",
                                "if: DOWN.mat.density < SELF.mat.density\n        and isType_EMPTY(DOWN)\n        or SELF.mat == sand", 1)
    flat = Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density",
                             "if: DOWN.mat.density < SELF.mat.density and isType_EMPTY(DOWN) or SELF.mat == sand", 1)
    nat_w, _ = both(se, wrapped)
    nat_f, _ = both(se, flat)
    assert nat_w.glsl_rules == nat_f.glsl_rules and "&& isType_EMPTY(down) || self.mat == MAT_sand" in nat_w.glsl_rules
    # a folded value is re-typed as a whole: `mirrored: false` + a continuation line is a string, not a bool
    both_fail(se, Y.BASE_OK.replace("    mirrored: false\n", "    mirrored: false\n      - SWAP SELF DOWN\n", 1), "InvalidType")
    # ... also inside a sequence item: `- SET SELF sand` + a deeper `- SWAP SELF DOWN` is ONE item "SET SELF sand - SWAP SELF DOWN"
    seq = Y.BASE_OK.replace("    do: SWAP SELF DOWN\n", "    do:\n    - SET SELF sand\n      - SWAP SELF DOWN\n", 1)
    try:
        nat = ("ok", se.parse_string(seq, compile=False).glsl_rules)
    except se.SandEngineError as e:
        nat = (e.kind,)
    try:
        orc = ("ok", oracle_lang.emit_glsl_rules(oracle_lang.parse_string(seq)))
    except oracle_lang.ParsingErr as e:
        orc = (e.kind,)
    assert nat == orc
    # a continuation line that looks like `key: value` is not a continuation (libyaml: "mapping values are not allowed")
    both_fail(se, Y.BASE_OK.replace("    mirrored: false\n", "    mirrored: false\n      extra: 1\n", 1), "Yaml")


def test_front_ends_agree_on_random_mutants(se):
    """Fixed-seed slice of scripts/diff_frontends.py: same ParsingErr class or byte-identical generated text."""
    import importlib.util
    import random
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("diff_frontends", Path(__file__).resolve().parent.parent / "scripts" / "diff_frontends.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = random.Random(20261017)
    outcomes = set()
    for _ in range(400):
        text = mod.mutate(rng.choice(mod.SOURCES), rng)
        nat, orc = mod.run(text)
        assert nat == orc, (nat[0], orc[0], text)
        outcomes.add(nat[0])
    assert {"ok", "Yaml", "MissingField", "InvalidType", "NotFound"} <= outcomes
