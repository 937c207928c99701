"""The GENERATED CUDA rule code, compiled for the host (tests/emu/host_emu.cpp shims the device intrinsics)
and stepped on the CPU, must reproduce the oracle bit for bit.  This checks the typed expression compiler,
the fat-cell tables, the integer RAND thresholds, Left/Right routing and operator precedence of the CUDA
back end without a GPU; the kernels around it are covered by the -m gpu parity tests."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import yaml_cases as Y
from conftest import DEFAULT_YAML, REPO
from sandengine_b200.grids import DEFAULT_IDS, DEFAULT_MIX, synthetic_grid


def build_emu(tmp_path_factory, tag, rules, defs=()):
    d = tmp_path_factory.mktemp(f"emu_{tag}")
    (d / "rules_gen.cuh").write_text(rules.cuda_header)
    so = d / "emu.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", *defs, "-I", str(d),
                           "-I", str(REPO / "sandengine_b200" / "csrc" / "kernels"),
                           str(REPO / "tests" / "emu" / "host_emu.cpp"), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.emu_step_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.emu_step_lut_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    return lib


def compare(lib, orc, g, steps, frame=1, lut=False):
    a, b = g.copy(), g.copy()
    step = lib.emu_step_lut_inplace if lut else lib.emu_step_inplace
    for s in range(steps):
        frame += 1
        orc.step_blocks_inplace(a, frame)
        step(b.ctypes.data, g.shape[1], g.shape[0], frame)
        assert np.array_equal(a, b), f"generated code diverges from the oracle at step {s + 1} (frame {frame})"
    return a


def test_default_rules(native_lib, tmp_path_factory, default_rules, oracle):
    lib = build_emu(tmp_path_factory, "default", default_rules)
    for (w, h, seed, steps) in [(96, 64, 1, 300), (33, 17, 2, 100), (2, 2, 3, 20)]:
        compare(lib, oracle, synthetic_grid(w, h, seed), steps)


def test_transition_table_default_rules(native_lib, tmp_path_factory, default_rules, oracle):
    """The transition table (built from the generated rule code) + its lookup with mirror folding, rand.y
    classes and the WALL/NULL slow path reproduces the oracle -- incl. grids with WALL / NULL / unknown ids."""
    lib = build_emu(tmp_path_factory, "default_lut", default_rules)
    assert lib.emu_lut_eligible() == 1
    n_sens = lib.emu_build_lut()
    assert 0 < n_sens <= 4095
    for (w, h, seed, steps) in [(96, 64, 1, 300), (33, 17, 2, 100), (2, 2, 3, 20), (128, 128, 3, 400)]:
        compare(lib, oracle, synthetic_grid(w, h, seed), steps, lut=True)
    g = synthetic_grid(64, 64, 9)
    rng = np.random.default_rng(3)
    g[rng.integers(0, 64, 60), rng.integers(0, 64, 60)] = 2      # WALL cells inside the grid
    g[rng.integers(0, 64, 20), rng.integers(0, 64, 20)] = 1      # NULL
    g[rng.integers(0, 64, 20), rng.integers(0, 64, 20)] = 9      # vine
    g[5, 5] = 77; g[6, 9] = 4000000000                           # unknown ids read as NULL
    compare(lib, oracle, g, 200, lut=True)


def test_rich_rules_left_right_precedence(native_lib, tmp_path_factory):
    import sandengine_b200 as se
    from oracle.build_oracle import load_oracle
    rules = se.parse_string(Y.RICH_YAML)
    lib = build_emu(tmp_path_factory, "rich", rules)
    assert lib.emu_lut_eligible() == 0      # uses pos.y, rand.x and non-mirrored rules
    orc = load_oracle(Y.RICH_YAML)
    out = compare(lib, orc, synthetic_grid(128, 96, 5, mix=Y.RICH_MIX, ids=Y.RICH_IDS), 300)
    assert len(np.unique(out)) > 5
    compare(lib, orc, synthetic_grid(61, 47, 6, mix=Y.RICH_MIX, ids=Y.RICH_IDS), 150, frame=1001)


def test_synthetic_64_material_rule_set(native_lib, tmp_path_factory):
    """configs[4] rule set (deep isType_* inheritance, non-mirrored LEFT rules) at crop size."""
    import sandengine_b200 as se
    from oracle.build_oracle import load_oracle
    from oracle import oracle_lang as L
    from sandengine_b200.synth_rules import synthetic_rule_set
    text, ids, mix = synthetic_rule_set(64, 28, seed=5)
    rules = se.parse_string(text)
    o = L.parse_string(text)
    assert len(rules.materials) == 64 and rules.glsl_rules == L.emit_glsl_rules(o) and rules.glsl_materials == L.emit_glsl_materials(o)
    kinds = [r.ruletype for r in rules.rules if r.used]
    assert kinds.count("Left") >= 4 and kinds.count("Right") >= 4 and kinds.count("Mirrored") >= 8 and len(kinds) >= 24
    depth = lambda t: 0 if not t.inherits else 1 + depth(next(x for x in o.types if x.name == t.inherits))
    assert max(depth(t) for t in o.types) >= 5
    lib = build_emu(tmp_path_factory, "synth64", rules)
    orc = load_oracle(text)
    g = synthetic_grid(160, 128, 5, mix=mix, ids=ids)
    out = compare(lib, orc, g, 200)
    assert not np.array_equal(out, g)


@pytest.mark.parametrize("rows", [4, 2, 8])      # 4 = the shipped tile height (32 rows); 2 / 8: the SE_LT_ROWS experiments
def test_lighting_kernel_phases_on_host(native_lib, tmp_path_factory, default_rules, oracle, rows):
    """se_light's two per-thread phases (term staging into the tile + ring, sliding-window combine; interior and
    rim CTAs) run CTA by CTA on the host and must give the oracle's light field BIT FOR BIT: tile/ring indexing,
    neighbour order, the x * 0.125 shortcut of the interior path and the general path on the rim."""
    lib = build_emu(tmp_path_factory, f"light{rows}", default_rules, defs=(f"-DSE_LT_ROWS={rows}",))
    lib.emu_light.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]
    rng = np.random.default_rng(11)
    saw_interior = 0
    for (w, h, seed, steps) in [(160, 128, 1, 6), (33, 17, 2, 4), (97, 99, 3, 4), (2, 2, 4, 3), (128, 33, 5, 3), (32, 32, 6, 3), (100, 200, 7, 3)]:
        cells = synthetic_grid(w, h, seed)
        light = rng.random((h, w, 4), dtype=np.float32)
        light[rng.random((h, w)) < 0.2, 3] = 0.0                 # falloff == 0 takes the running maximum
        light[rng.random((h, w)) < 0.1] = 0.0
        frame = 1
        for s in range(steps):
            frame += 1
            new_cells, want, _ = oracle.step_cells(cells, frame, light)
            got = np.full_like(want, np.nan)
            n_int = C.c_int(0)
            lib.emu_light(cells.ctypes.data, new_cells.ctypes.data, light.ctypes.data, got.ctypes.data, w, h, 0, h, C.byref(n_int))
            saw_interior += n_int.value
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{w}x{h} step {s + 1}: light differs from the oracle"
            if w == 160 and s == 0:
                # a strip (local rows [32, 102) of the grid): every row except the first and last local one sees all of
                # its neighbours, so it must equal the full-grid result
                gy0, hl = 32, 70
                part = np.full((hl, w, 4), np.nan, np.float32)
                lib.emu_light(cells[gy0:gy0 + hl].ctypes.data, new_cells[gy0:gy0 + hl].ctypes.data,
                              np.ascontiguousarray(light[gy0:gy0 + hl]).ctypes.data, part.ctypes.data, w, hl, gy0, h, None)
                assert np.array_equal(part[1:-1].view(np.uint32), want[gy0 + 1:gy0 + hl - 1].view(np.uint32))
            cells, light = new_cells, want
    assert saw_interior > 0


def test_running_census_deltas_on_host(native_lib, tmp_path_factory, default_rules, oracle):
    """Running census (SE_FLAG_RUNNING_CENSUS): the flagged table and the per-block deltas that K1c's census variant
    applies, run on the host: after every step the maintained census equals a recount of the owned rows -- full grids,
    ragged sizes, WALL / NULL / unknown ids, and strips (owned rows inside a larger buffer)."""
    lib = build_emu(tmp_path_factory, "census", default_rules)
    lib.emu_step_lut_census.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n_pop = lib.emu_build_census_lut()
    n4 = len(default_rules.materials) ** 4
    assert 0 <= n_pop < n4         # a real filter: few plain outcomes change the population (SET rules fire on rand.y: mostly pool entries)

    def recount(cells, y0, y1):
        return np.bincount(np.minimum(cells[y0:y1], 255).ravel(), minlength=256).astype(np.int64)

    filtered = C.c_longlong(0)
    cases = [(96, 64, 1, 200, 0, 64), (33, 17, 2, 80, 0, 17), (2, 2, 3, 12, 0, 2), (64, 96, 5, 150, 32, 64), (40, 50, 6, 100, 10, 11)]
    for (w, h, seed, steps, own0, own1) in cases:
        g = synthetic_grid(w, h, seed)
        if seed == 5:
            rng = np.random.default_rng(3)
            g[rng.integers(0, h, 40), rng.integers(0, w, 40)] = 2          # WALL inside the grid
            g[rng.integers(0, h, 15), rng.integers(0, w, 15)] = 1          # NULL
            g[40, 5] = 77; g[33, 9] = 4000000000; g[31, 2] = 300           # unknown ids (normalised to NULL when written)
        ref = g.copy()
        census = recount(g, own0, own1)
        frame = 1
        for s in range(steps):
            frame += 1
            oracle.step_blocks_inplace(ref, frame)
            lib.emu_step_lut_census(g.ctypes.data, w, h, frame, own0, own1, census.ctypes.data, C.byref(filtered))
            assert np.array_equal(g, ref), f"{w}x{h} step {s + 1}: cells"
            assert np.array_equal(census, recount(g, own0, own1)), f"{w}x{h} rows {own0}..{own1} step {s + 1}: census"
    assert filtered.value > 1000       # the filter did skip changed blocks (pure swaps)


def test_lit_strip_schedule_on_host(native_lib, tmp_path_factory, default_rules, oracle):
    """Strips WITH lighting: the light stencil invalidates one ghost row per step, so G ghost rows of ids
    and light are good for G steps (StripPlan(lighting=True)).  Emulated rank by rank with the kernels' strip semantics
    (oracle strip step for the ids, se_light's host phases for the light); the rows that the schedule declares invalid
    are POISONED after every step, so agreement with the full-grid oracle proves that owned rows never read them."""
    from sandengine_b200.distributed import StripPlan
    lib = build_emu(tmp_path_factory, "litstrip", default_rules)
    lib.emu_light.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]
    rng = np.random.default_rng(5)
    for (w, h, world, G, steps) in [(96, 120, 3, 4, 19), (70, 64, 2, 2, 9), (64, 200, 4, 6, 14)]:
        plan = StripPlan(w, h, world, G, lighting=True)
        assert plan.steps_per_exchange == G
        cells = synthetic_grid(w, h, 7)
        light = rng.random((h, w, 4), dtype=np.float32)
        light[rng.random((h, w)) < 0.2, 3] = 0.0
        # ranks: local buffers with ghosts
        loc = []
        for r in range(world):
            b, e = plan.rows(r); gt, gb = plan.ghosts(r)
            loc.append(dict(gy0=b - gt, b=b, e=e, gt=gt, gb=gb, cells=cells[b - gt:e + gb].copy(), light=light[b - gt:e + gb].copy()))
        ref_c, ref_l, frame = cells, light, 1
        done = 0
        for chunk in plan.chunks(steps):
            for n in range(1, chunk + 1):
                frame += 1
                ref_c, ref_l, _ = oracle.step_cells(ref_c, frame, ref_l)
                for L in loc:
                    old = L["cells"].copy()
                    oracle.step_blocks_strip(L["cells"], L["gy0"], h, frame)
                    out = np.empty_like(L["light"])
                    hl = L["cells"].shape[0]
                    lib.emu_light(old.ctypes.data, L["cells"].ctypes.data, L["light"].ctypes.data, out.ctypes.data, w, hl, L["gy0"], h, None)
                    L["light"] = out
                    # rows the schedule gives up after n steps since the last exchange: n from each inner edge
                    if L["gt"]:
                        L["cells"][:n] = 3; L["light"][:n] = np.float32(1e9)
                    if L["gb"]:
                        L["cells"][hl - n:] = 3; L["light"][hl - n:] = np.float32(1e9)
            # exchange: owners push their boundary rows (ids + light) into the neighbours' ghosts
            for r, L in enumerate(loc):
                if r > 0:
                    up = loc[r - 1]
                    L["cells"][:L["gt"]] = up["cells"][up["gt"] + (up["e"] - up["b"]) - L["gt"]:up["gt"] + (up["e"] - up["b"])]
                    L["light"][:L["gt"]] = up["light"][up["gt"] + (up["e"] - up["b"]) - L["gt"]:up["gt"] + (up["e"] - up["b"])]
                if r < world - 1:
                    dn = loc[r + 1]
                    hl = L["cells"].shape[0]
                    L["cells"][hl - L["gb"]:] = dn["cells"][dn["gt"]:dn["gt"] + L["gb"]]
                    L["light"][hl - L["gb"]:] = dn["light"][dn["gt"]:dn["gt"] + L["gb"]]
            done += chunk
            for L in loc:
                own = slice(L["gt"], L["gt"] + L["e"] - L["b"])
                assert np.array_equal(L["cells"][own], ref_c[L["b"]:L["e"]]), f"{w}x{h}/{world} ids after {done} steps"
                assert np.array_equal(L["light"][own].view(np.uint32), ref_l[L["b"]:L["e"]].view(np.uint32)), f"{w}x{h}/{world} light after {done} steps"


def _stage_mods(se, mods, n_materials):
    """What se_sim_step stages for the kernels: the first min(len, 256) records up to the first mod_size == 0, unknown
    material ids replaced by NULL (api.cpp)."""
    out = np.zeros(len(mods), se.MOD_DTYPE)
    n = 0
    for m in mods[:256]:
        if m["mod_size"] == 0:
            break
        out[n] = m
        if not (0 <= m["mod_matID"] < n_materials):
            out[n]["mod_matID"] = 1
        n += 1
    return out[:n]


@pytest.mark.parametrize("seed,n_mat,n_rules,kinds", [(201, 7, 23, ("mirrored",)), (216, 12, 22, ("mirrored",)), (110, 20, 28, None)])
def test_random_rule_sets_differential(native_lib, tmp_path_factory, seed, n_mat, n_rules, kinds):
    """Random rule sets (scripts/diff_campaign.py runs many more): oracle == generated code on the host == transition
    table (one shared-memory table for small mirrored-only sets, one table per view otherwise; mode 2 = table in global
    memory for the 20-material Left/Right set)."""
    import sandengine_b200 as se
    from oracle.build_oracle import load_oracle
    from sandengine_b200.synth_rules import synthetic_rule_set
    text, ids, mix = synthetic_rule_set(n_mat, n_rules, seed=seed, **({"kinds": kinds} if kinds else {}))
    rules = se.parse_string(text)
    assert "#define SE_LUT_ELIGIBLE 1" in rules.cuda_header
    assert ("#define SE_LUT_TWO_TABLES 1" in rules.cuda_header) == (kinds is None or n_mat > 11)
    orc = load_oracle(text)
    lib = build_emu(tmp_path_factory, f"rand{seed}", rules)
    g = synthetic_grid(48, 40, seed, mix=mix, ids=ids)
    ref = compare(lib, orc, g, 60)
    assert (ref != g).sum() > 500
    assert lib.emu_lut_eligible() == 1 and lib.emu_build_lut() >= 0
    assert lib.emu_lut_mode() == (1 if n_mat <= 11 and kinds is not None else 2)
    compare(lib, orc, g, 60, lut=True)


def test_two_table_lut_for_left_right_rule_sets(native_lib, tmp_path_factory, monkeypatch):
    """Rule sets with Left/Right rules get one transition table per view (unmirrored / mirrored evaluation) instead of
    falling back to the generated code.  Table build + lookup on the host vs the oracle for random mixed rule sets and
    for the Left/Right test YAMLs that use no pos / rand.x; SE_LUT_FORCE_MODE=0 switches the tables off."""
    import sandengine_b200 as se
    from oracle.build_oracle import load_oracle
    from sandengine_b200.synth_rules import synthetic_rule_set
    cases = []
    for seed, n_mat, n_rules in [(401, 9, 20), (402, 12, 24), (403, 6, 9)]:
        text, ids, mix = synthetic_rule_set(n_mat, n_rules, seed=seed)      # mirrored + right + left rules
        cases.append((f"synth{seed}", text, dict(mix=mix, ids=ids)))
    cases.append(("base_ok", Y.BASE_OK, {}))
    monkeypatch.setenv("SE_LUT_FORCE_MODE", "0")
    off = se.parse_string(cases[0][1])
    assert "#define SE_LUT_ELIGIBLE 0" in off.cuda_header                     # tables switched off: generated code only
    monkeypatch.delenv("SE_LUT_FORCE_MODE")
    n_two = 0
    for name, text, gk in cases:
        rules = se.parse_string(text)
        if "#define SE_LUT_TWO_TABLES 1" not in rules.cuda_header:
            continue                                                           # uses pos / other rand lanes: stays on the generated code
        n_two += 1
        orc = load_oracle(text)
        lib = build_emu(tmp_path_factory, f"lr_{name}", rules)
        assert lib.emu_lut_eligible() == 1 and lib.emu_build_lut() >= 0
        for (w, h, seed, steps) in [(48, 40, 5, 80), (33, 17, 6, 40)]:
            g = synthetic_grid(w, h, seed, **gk) if gk else (synthetic_grid(w, h, seed) % len(rules.materials)).astype(np.uint32)
            ref = compare(lib, orc, g, steps, lut=True)
            compare(lib, orc, g, steps)
            assert (ref != g).sum() > 50, name
    assert n_two >= 3
