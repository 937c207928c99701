"""Native (C++) rule-language front end + CUDA C back end through the C ABI.  No GPU needed: NVRTC
cross-compiles for sm_100a on the build host."""
import ctypes as C
import hashlib
import json
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import DEFAULT_YAML, REPO
from oracle import oracle_lang as L
import yaml_cases as Y

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_gen_sha256.json").read_text())


def _sha(s: str):
    b = s.encode()
    return {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}


def test_library_exports_every_declared_symbol(native_lib):
    header = (REPO / "include" / "sandengine_b200.h").read_text()
    declared = set(re.findall(r"\b(se_[a-z0-9_]+)\s*\(", header))
    declared -= {"se_status"}
    assert len(declared) >= 25
    from sandengine_b200 import _capi
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(native_lib, name), name
    assert b"sm_100a" in native_lib.se_version()


def test_abi_struct_layout():
    from sandengine_b200 import _capi
    assert C.sizeof(_capi.se_modification) == 32      # simulation.rs:45-56, std140 stride of the UBO
    assert C.sizeof(_capi.se_create_params) == 36     # 9 x uint32 (device_share appended in round 2)


def test_glsl_known_answer(default_rules):
    assert _sha(default_rules.glsl_materials) == GOLD["shaders/compute/gen/materials.glsl"]
    assert _sha(default_rules.glsl_rules) == GOLD["shaders/compute/gen/rules.glsl"]


def test_parsing_result_surface(default_rules):
    r = default_rules
    assert [m.name for m in r.materials] == ["EMPTY", "NULL", "WALL", "sand", "rock", "water", "radioactive", "smoke",
                                             "toxic_sludge", "vine", "dirt"]
    assert [m.id for m in r.materials] == list(range(11))
    assert r.material_id("water") == 5
    assert [m.name for m in r.selectable_materials] == ["EMPTY", "sand", "rock", "water", "radioactive", "smoke", "toxic_sludge", "dirt"]
    assert abs(r.materials[8].density - 1.49) < 1e-6 and r.materials[6].emission[1] == np.float32(0.7)
    assert [x.name for x in r.rules if x.used] == ["fall_slide", "fall_slide_dirt", "horizontal_slide", "rise_up", "dissolve",
                                                   "grow", "grow_up", "die_off"]
    assert len(r.cubin) > 1000 and r.cubin[:4] == b"\x7fELF"


@pytest.mark.parametrize("name,text,kind", Y.ERROR_CASES, ids=[c[0] for c in Y.ERROR_CASES])
def test_error_classes_match_oracle_parser(native_lib, name, text, kind):
    import sandengine_b200 as se
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(text, compile=False)
    assert ei.value.kind == kind
    assert f"({kind})" in str(ei.value)


@pytest.mark.parametrize("text,want", [("0b11", 3.0), ("007", None), ("017", None), ("-012", None), ("-0x1F", -31.0), ("+0x10", 16.0), ("0o17", 15.0),
                                       ("-0", 0.0), ("+12", 12.0), ("1e3", 1000.0), ("00.5", 0.5), ("0x", None), ("0b2", None), ("-.5", -0.5),
                                       ("1.", 1.0), ("0", 0.0)])
def test_plain_scalar_numbers_resolve_like_serde_yaml(native_lib, text, want):
    """serde_yaml 0.9.34 (the reference's Cargo.lock) reads signed hex / octal / binary integers and treats a decimal
    with a leading zero as a string (-> InvalidType for a density).  Product and oracle parser must agree case by case."""
    import sandengine_b200 as se
    doc = Y.BASE_OK.replace("density: 1.5", f"density: {text}")
    if want is None:
        with pytest.raises(se.SandEngineError) as ei:
            se.parse_string(doc, compile=False)
        assert ei.value.kind == "InvalidType"
        with pytest.raises(L.ParsingErr) as eo:
            L.parse_string(doc)
        assert eo.value.kind == "InvalidType"
    else:
        assert se.parse_string(doc, compile=False).materials[-1].density == want
        assert float(L.parse_string(doc).materials[-1].density) == want


def test_unused_conflicting_rule_is_accepted(native_lib):
    """the LEFT / RIGHT conflict is an error of a USED rule only (the reference emits used rules only)."""
    import sandengine_b200 as se
    r = se.parse_string(Y.UNUSED_LEFT_CONFLICT_OK, compile=False)
    assert [x.name for x in r.rules if x.used] == ["gravity", "slide_diagonally"]
    o = L.parse_string(Y.UNUSED_LEFT_CONFLICT_OK)
    assert [x.name for x in o.rules if x.used] == ["gravity", "slide_diagonally"]
    # ... and the same draft rule is rejected as soon as a type refers to it
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(Y.UNUSED_LEFT_CONFLICT_OK.replace("base_rules: [gravity, slide_diagonally]", "base_rules: [gravity, draft_never_used]"), compile=False)
    assert ei.value.kind == "NotRecognized"
    with pytest.raises(L.ParsingErr) as eo:
        L.parse_string(Y.UNUSED_LEFT_CONFLICT_OK.replace("base_rules: [gravity, slide_diagonally]", "base_rules: [gravity, draft_never_used]"))
    assert eo.value.kind == "NotRecognized"


@pytest.mark.parametrize("text", [Y.BASE_OK, Y.RICH_YAML], ids=["base", "rich"])
def test_front_ends_agree_on_text(native_lib, text):
    """Two independent parsers (C++ product, Python oracle) must build the same rule text."""
    import sandengine_b200 as se
    n = se.parse_string(text)
    o = L.parse_string(text)
    assert n.glsl_materials == L.emit_glsl_materials(o)
    assert n.glsl_rules == L.emit_glsl_rules(o)
    assert [(r.name, r.used, r.ruletype) for r in n.rules] == [(r.name, r.used, r.effective_type) for r in o.rules]


def test_yaml_subset_reader(native_lib):
    import sandengine_b200 as se
    # comments, quoted scalars, flow + block sequences, YAML-1.2 booleans (`yes` is a string, not a bool)
    y = Y.BASE_OK.replace("mirrored: true", "mirrored: yes")
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(y, compile=False)
    assert ei.value.kind == "InvalidType"
    y = Y.BASE_OK.replace("do: SWAP SELF DOWN\n", "do: 'SWAP SELF DOWN'  # quoted\n").replace("density: 1.5", "density: 2")
    r = se.parse_string(y, compile=False)
    assert r.materials[3].density == 2.0
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string("rules:\n\tbad: 1\n", compile=False)
    assert ei.value.kind == "Yaml"
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string("rules: {a: 1, a: 2}\ntypes: {}\nmaterials: {}\n", compile=False)
    assert ei.value.kind == "Yaml"


def test_probability_thresholds_are_exact(default_rules):
    """`rand.y <= p` is lowered to `u <= U`: check U against the float expression at and around the boundary."""
    hdr = default_rules.cuda_header
    found = {}
    for body in hdr.split("static __device__ __forceinline__ void se_rule_")[1:]:
        m = re.search(r"rnd\.u\[1\] <= (0x[0-9a-f]+)u", body.split("\n}\n")[0])
        if m:
            found[body.split("(")[0]] = m.group(1)
    expect = {"fall_slide_dirt": 0.1, "dissolve": 0.004, "grow": 0.001, "grow_up": 0.004, "die_off": 0.3}
    assert set(found) == set(expect)
    f = lambda u: np.float32(np.uint32(u)) / np.float32(4294967296.0)
    for name, p in expect.items():
        U = int(found[name], 16)
        p32 = np.float32(p)
        assert f(U) <= p32 and not (f(U + 1) <= p32)
        for u in range(max(0, U - 300), U + 300):
            assert (f(u) <= p32) == (u <= U)
    m = re.search(r"#define SE_MIRROR_UMAX (0x[0-9a-f]+)u", hdr)
    assert int(m.group(1), 16) == 2**31 - 65


def test_fat_cell_tables(default_rules):
    hdr = default_rules.cuda_header
    fat = [int(x, 16) for x in re.search(r"se_fat_table\[256\] = \{(.*?)\};", hdr, re.S).group(1).replace("u", "").replace("\n", " ").split(",") if x.strip()]
    assert len(fat) == 256
    o = L.parse_path(DEFAULT_YAML)
    dens = sorted({float(m.density) for m in o.materials})
    for m in o.materials:
        w = fat[m.id]
        assert w & 0xFF == m.id
        assert (w >> 8) & 0xFF == next(t.id for t in o.types if t.name == m.mattype)
        assert w >> 24 == dens.index(float(m.density))
        assert bool(w & (1 << 16)) == (m.mattype in ("WALL", "NULL"))
        emissive = any(float(e) != 0 for e in m.emission[:3])
        assert bool(w & (1 << 17)) == emissive
        assert bool(w & (1 << 18)) == ((not emissive) and m.mattype != "EMPTY")
    assert all(w == fat[1] for w in fat[11:])   # unknown ids read as NULL


def test_rich_rule_set_compiles_for_sm100a(native_lib):
    import sandengine_b200 as se
    r = se.parse_string(Y.RICH_YAML)
    assert r.cubin[:4] == b"\x7fELF"
    assert "#define SE_HAVE_LEFT_RULES 1" in r.cuda_header and "#define SE_HAVE_RIGHT_RULES 1" in r.cuda_header
    assert "#define SE_RAND_LANES 3u" in r.cuda_header


def test_compile_time_errors_for_text_the_reference_would_pass_to_glsl(native_lib):
    import sandengine_b200 as se
    # SET target is not validated by the reference parser (rules.rs:406-409); the shader compile would fail.
    y = Y.BASE_OK.replace("do: SWAP SELF DOWN\n", "do: SET SELF lava\n")
    assert se.parse_string(y, compile=False) is not None      # parse_string accepts it, like the reference
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(y, compile=True)
    assert ei.value.kind == "NotFound"
    # same material compared twice: the reference's global replace yields MAT_MAT_sand
    y = Y.BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", "if: DOWN.mat == sand or SELF.mat == sand")
    with pytest.raises(se.SandEngineError) as ei:
        se.parse_string(y)
    assert ei.value.kind == "NotFound" and "MAT_MAT_sand" in str(ei.value)


def test_no_device_fails_loudly(default_rules):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sandengine_b200 as se
    with pytest.raises(se.SandEngineError) as ei:
        se.Simulation(default_rules, (64, 64))
    assert ei.value.kind == "Cuda" and "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_geometry_before_touching_a_device(default_rules):
    """Argument validation of se_sim_create runs before any CUDA call (same errors with and without a GPU)."""
    import sandengine_b200 as se
    for size, kw in (((0, 64), {}), ((64, 0), {}), ((4, 600_000), {}),
                     ((64, 64), dict(row_begin=3, row_end=33)),        # odd strip boundary
                     ((64, 64), dict(row_begin=32, row_end=16)),       # empty strip
                     ((64, 64), dict(row_begin=0, row_end=128)),       # beyond the grid
                     ((64, 64), dict(row_begin=0, row_end=32, halo_rows=3))):   # ghost zones start on even rows
        with pytest.raises(se.SandEngineError) as ei:
            se.Simulation(default_rules, size, **kw)
        assert ei.value.kind == "InvalidArg", (size, kw, str(ei.value))


def test_yaml_syntax_variants_parse_identically(native_lib):
    """Block / flow / quoted / commented spellings of one document give the same parse (and the same GLSL text)."""
    import sandengine_b200 as se
    ref = se.parse_string(Y.BASE_OK, compile=False)
    variants = [
        # flow mappings everywhere, trailing commas, comments
        """
rules: {gravity: {if: DOWN.mat.density < SELF.mat.density, do: SWAP SELF DOWN, mirrored: false},   # falls
        slide_diagonally: {if: DOWNRIGHT.mat.density < SELF.mat.density, do: SWAP SELF DOWNRIGHT, mirrored: true,}}
types: {movable_solid: {base_rules: [gravity, slide_diagonally,]}}
materials: {sand: {color: [1.0, 1.0, 0.0, 1.0], type: movable_solid, density: 1.5, selectable: true}}
""",
        # block sequences, quoted scalars and keys, a document marker, odd indentation widths
        """---
"rules":
    gravity:
        'if': "DOWN.mat.density < SELF.mat.density"
        do: 'SWAP SELF DOWN'
        mirrored: False
    slide_diagonally:
        if: DOWNRIGHT.mat.density < SELF.mat.density
        do:
        - SWAP SELF DOWNRIGHT
        mirrored: TRUE
types:
    movable_solid:
        base_rules:
            - gravity
            - "slide_diagonally"
materials:
    sand:
        color:
            - 1.0
            - 1.0
            - 0.0
            - 1.0
        type: movable_solid
        density: 1.5e0
        selectable: true
""",
    ]
    for v in variants:
        r = se.parse_string(v, compile=False)
        assert r.glsl_materials == ref.glsl_materials and r.glsl_rules == ref.glsl_rules
        assert L.emit_glsl_rules(L.parse_string(v)) == ref.glsl_rules       # PyYAML-based oracle parser agrees
