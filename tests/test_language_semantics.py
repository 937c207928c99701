"""Expression semantics of rule conditions beyond what data/materials.yaml uses: the three evaluators -- the
product's generated CUDA code (compiled for the host), the C oracle (rule text rewritten to C) and the Python
oracle (eval of the GLSL-shaped text) -- must agree bit for bit.  CPU only."""
import numpy as np
import pytest

import yaml_cases as Y
from oracle.build_oracle import load_oracle
from oracle.pyoracle import PyOracle
from sandengine_b200.grids import synthetic_grid
from test_codegen_host_emulation import build_emu, compare


def test_expression_grammar_three_way(native_lib, tmp_path_factory):
    import sandengine_b200 as se
    rules = se.parse_string(Y.EXPR_YAML)
    hdr = rules.cuda_header
    assert "#define SE_RAND_LANES 15u" in hdr           # rand.x (mirror + spark_fade), .y (probabilities), .z, .w
    assert "#define SE_LUT_ELIGIBLE 0" in hdr           # pos / frame / rand.z make the transition table impossible
    assert "#define SE_HAVE_RIGHT_RULES 1" in hdr
    lib = build_emu(tmp_path_factory, "expr", rules)
    c_orc, py = load_oracle(Y.EXPR_YAML), PyOracle(Y.EXPR_YAML)
    g = synthetic_grid(40, 28, 5, mix=Y.EXPR_MIX, ids=Y.EXPR_IDS)
    out = compare(lib, c_orc, g, 60)                     # generated CUDA code == C oracle, every step
    assert not np.array_equal(out, g) and len(np.unique(out)) >= 5
    b, _ = py.run(g, 1, 60)                              # == Python oracle
    assert np.array_equal(out, b)
    compare(lib, c_orc, synthetic_grid(17, 9, 6, mix=Y.EXPR_MIX, ids=Y.EXPR_IDS), 40, frame=997)


def test_glsl_builtins_bit_operators_and_conditional_three_way(native_lib, tmp_path_factory):
    """The reference hands conditions to the GLSL compiler verbatim, so scalar built-ins (abs min max clamp mod floor ceil
    fract sign step sqrt float() int()), integer & | ^ ~ << >>, ?: and vector == / != (`mat.emission.rgb != vec3(0.0)`) are legal in a rule file.  Generated CUDA code (host)
    == C oracle step by step; with the reference present also == its own shader compiled for the CPU."""
    import sandengine_b200 as se
    from oracle import build_ref, oracle_lang
    rules = se.parse_string(Y.FUNC_YAML)                  # incl. NVRTC for sm_100a
    hdr = rules.cuda_header
    for needle in ("floorf(", "ceilf(", "__fsqrt_rn(", " >> 1)", " & 1)", "(~px)", " ? ", "se_emission_table[SE_ID(s) * 4 + 2]) == 0x0p+0f"):
        assert needle in hdr, needle
    lib = build_emu(tmp_path_factory, "func", rules)
    c_orc = load_oracle(Y.FUNC_YAML)
    g = synthetic_grid(64, 40, 5, mix=Y.FUNC_MIX, ids=Y.FUNC_IDS)
    out = compare(lib, c_orc, g, 120)
    assert int((out != g).sum()) > 500 and len(np.unique(out)) >= 5
    compare(lib, c_orc, synthetic_grid(33, 17, 6, mix=Y.FUNC_MIX, ids=Y.FUNC_IDS), 40, frame=2001)
    if build_ref.reference_available():
        res = oracle_lang.parse_string(Y.FUNC_YAML)
        assert rules.glsl_rules == oracle_lang.emit_glsl_rules(res)
        ref = build_ref.load_ref(oracle_lang.emit_glsl_materials(res), oracle_lang.emit_glsl_rules(res))
        ref.create(64, 40); ref.upload_ids(g); ref.frame = 1
        ref.step(120)
        assert np.array_equal(ref.download_ids(), out)


@pytest.mark.parametrize("cond,kind", [
    ("sin(rand.x) > 0.5", "NotFound"),          # transcendentals have no bit-exact meaning: refused, not approximated
    ("abs(1, 2) > 0", "NotRecognized"),
    ("min(SELF, 1) > 0", "NotRecognized"),
    ("1 & 2.0", "NotRecognized"),
    ("pos.x << 1.5 > 0", "NotRecognized"),
    ('"(pos.x > 1 ? 1 : true) == 1"', "NotRecognized"),       # (a ?: needs YAML quotes: ': ' cannot occur in a plain scalar)
    ("(pos.x > 1 ? 1 : true) == 1", "Yaml"),
    ("pos.x > 1 ? true", "NotRecognized"),
    ("SELF.mat.emission.rgb == vec2(0.0)", "NotRecognized"),   # sizes differ
    ("SELF.mat.emission.rgb < vec3(0.0)", "NotRecognized"),    # only == / != have a scalar result
    ("SELF.mat.emission.rgb", "NotRecognized"),
    ("SELF.mat.density +", "NotRecognized"),
    ("SELF.density < 1.0", "NotRecognized"),
    ("isType_movable_solid(SELF", "NotRecognized"),
    ("isType_movable_solid(3)", "NotRecognized"),
    ("SELF.mat == 3", "NotFound"),           # the reference's material regex rejects it at parse time
    ("SELF.mat < DOWN.mat", "NotRecognized"),
    ("UP.mat.density < 1.0", "NotFound"),
    ("SELF.mat.density", "NotRecognized"),
    ("rand.q < 0.5", "NotRecognized"),
    ("SELF.mat.type == TYPE_nothing", "NotFound"),
    ("3 % 2.0 == 1", "NotRecognized"),
])
def test_malformed_conditions_are_clean_errors(native_lib, cond, kind):
    """Text the reference would hand to the GLSL compiler (and panic on, simulation.rs:133-137) is an error code here."""
    import sandengine_b200 as se
    y = Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", f"if: {cond}")
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(y)
    assert ei.value.kind == kind, str(ei.value)


def test_density_literal_folding_and_type_masks(native_lib):
    import re
    import sandengine_b200 as se
    y = Y.RICH_YAML
    hdr = se.parse_string(y).cuda_header
    cool = hdr.split("se_rule_cool(")[1].split("\n}\n")[0]
    assert "SE_RANK(s) >= 8u" in cool and "se_density_table" not in cool      # density > 2.0 decided on ranks
    sink = hdr.split("se_rule_sink(")[1].split("\n}\n")[0]
    # granular (4) + fine_granular (5) -> 0x30, static (3) incl. children -> 0x38, fluid (6) + thick_fluid (7) -> 0xc0
    assert "SE_ISTYPE32(s, 0x00000030u)" in sink and "SE_ISTYPE32(d, 0x00000038u)" in sink and "SE_ISTYPE32(s, 0x000000c0u)" in sink
