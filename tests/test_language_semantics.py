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


@pytest.mark.parametrize("cond,kind", [
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
