"""Cases shared by the reference-shader golden generator (tests/golden/make_ref_shader_goldens.py), the CPU tests
(oracle vs goldens, oracle vs the translated reference shader) and the GPU tests (CUDA path vs goldens).

Every input is a pure function of small integers (counter hashes from sandengine_b200.grids, no RNG library), so the
generator run in the build container and the test run on the GPU box see identical grids, light fields and
modification lists.  An *engine* is anything with the reference's `Simulation` contract (simulation.rs:195-253):
RefEngine = the reference's own shader compiled for the CPU (oracle/build_ref.py), OracleEngine = the C restatement,
GpuEngine (tests/test_gpu_ref_goldens.py) = the product through the C ABI.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

import yaml_cases as Y
from sandengine_b200.grids import hashi, synthetic_grid
from sandengine_b200.synth_rules import synthetic_rule_set

REPO = Path(__file__).resolve().parent.parent
GOLDEN_JSON = Path(__file__).parent / "golden" / "ref_shader_goldens.json"
GOLDEN_NPZ = Path(__file__).parent / "golden" / "ref_shader_arrays.npz"
MOD_DTYPE = np.dtype([("position", "<i4", (2,)), ("mod_shape", "<i4"), ("mod_size", "<i4"), ("mod_matID", "<i4"), ("_pad4", "<i4", (3,))])


@dataclass
class Case:
    name: str
    rules: str = "default"                 # "default" | "expr" | "func" | "rich" | "synthR:<n_materials>:<n_rules>:<seed>" | "synthLR:..."
    W: int = 64
    H: int = 48
    seed: int = 1
    steps: int = 50
    frame0: int = 1                        # params.frame before the first run(); 1 = the uploaded grid is the frame-1 state
    checkpoints: Tuple[int, ...] = ()      # ids hashed after these step counts (the last step always)
    lighting: bool = False
    light0: str = "zero"                   # "zero" | "hash"
    mods: str = "none"                     # "none" | "stamps" | "edge" | "config3" | "huge"
    grid: str = "synthetic"                # "synthetic" | "unknown_ids"
    store_light: str = "none"              # "full" | "sub8" | "none" (what goes into the .npz)
    store_color: bool = False
    slow: bool = False                     # only in the generator + GPU test, not in the quick CPU cross-check

    def all_checkpoints(self):
        return tuple(sorted(set(self.checkpoints) | {self.steps}))


CASES = [
    # BASELINE configs[0]: 256 x 256, default rule set, 1000 steps, seed 1 -- ids every 100 steps, lighting on
    Case("default_256x256_seed1_lit_1000", W=256, H=256, seed=1, steps=1000, checkpoints=tuple(range(100, 1001, 100)),
         lighting=True, store_light="sub8", slow=True),
    # BASELINE configs[1] at full size (4096 x 4096, seed 2): the short horizon SURVEY.md 8d names, steps 1-5 and 100
    Case("default_4096x4096_seed2_100", W=4096, H=4096, seed=2, steps=100, checkpoints=(1, 2, 3, 4, 5), slow=True),
    Case("default_130x66_seed11_257", W=130, H=66, seed=11, steps=257, checkpoints=(1, 2, 3, 4, 5, 100)),
    Case("default_33x17_seed4_90", W=33, H=17, seed=4, steps=90),
    Case("default_1x9_seed3_40", W=1, H=9, seed=3, steps=40),
    Case("default_9x1_seed4_40", W=9, H=1, seed=4, steps=40),
    Case("default_2x2_seed6_24", W=2, H=2, seed=6, steps=24, checkpoints=tuple(range(1, 24))),
    Case("default_64x48_seed9_lit_stamps_30", W=64, H=48, seed=9, steps=30, lighting=True, mods="stamps", store_light="full"),
    Case("default_96x72_seed4_mods_edge_60", W=96, H=72, seed=4, steps=60, mods="edge", checkpoints=(4, 6, 8)),
    Case("default_80x64_seed9_lit_hashlight_stamps_50", W=80, H=64, seed=9, steps=50, lighting=True, light0="hash", mods="stamps",
         store_light="full"),
    # BASELINE configs[3] semantics at a size the CPU shader finishes in seconds: stamps + explosion every frame, lighting on
    Case("default_512x512_seed4_config3_lit_24", W=512, H=512, seed=4, steps=24, lighting=True, mods="config3", checkpoints=(1, 8, 16),
         store_light="sub8"),
    Case("default_32x32_seed2_frame0_lit_3", W=32, H=32, seed=2, steps=3, frame0=0, lighting=True, light0="hash", mods="stamps",
         checkpoints=(1, 2), store_light="full"),
    Case("default_40x30_seed8_unknown_ids_50", W=40, H=30, seed=8, steps=50, grid="unknown_ids", checkpoints=(1, 2)),
    Case("default_64x48_seed23_colour_1", W=64, H=48, seed=23, steps=1, store_color=True),
    Case("expr_96x64_seed5_120", rules="expr", W=96, H=64, seed=5, steps=120),
    Case("synthR12_96x64_seed7_100", rules="synthR:12:20:7", W=96, H=64, seed=7, steps=100),
    Case("synthR64_96x64_seed5_60", rules="synthR:64:28:5", W=96, H=64, seed=5, steps=60, slow=True),
    # LEFT rules: outputs of the reference's shader template with gen/rules.glsl from the PATCHED emitter (see RefEngine)
    Case("rich_patchedleft_128x96_seed5_200", rules="rich", W=128, H=96, seed=5, steps=200, checkpoints=(1, 2, 3, 4)),
    Case("synthLR64_patchedleft_96x64_seed5_60", rules="synthLR:64:28:5", W=96, H=64, seed=5, steps=60, slow=True),
    Case("func_64x40_seed5_120", rules="func", W=64, H=40, seed=5, steps=120, checkpoints=(1, 2, 3, 4, 40)),
    # mod_size up to INT_MAX (kept last: the CUDA path stages such sizes as 2^30, see api.cpp::se_sim_step)
    Case("default_48x40_seed3_mods_huge_10", W=48, H=40, seed=3, steps=10, mods="huge", checkpoints=tuple(range(1, 10))),
]
CASE_BY_NAME = {c.name: c for c in CASES}


# ---- inputs -------------------------------------------------------------------------------------------------------
def rules_for(case: Case):
    """-> (yaml_text, ids, mix); the synthR sets have mirrored + non-mirrored RIGHT rules only: the reference cannot
    compile LEFT rules, so they cannot be pinned by its shader."""
    if case.rules == "default":
        return (REPO / "data" / "materials.yaml").read_text(), None, None
    if case.rules == "expr":
        return Y.EXPR_YAML, Y.EXPR_IDS, Y.EXPR_MIX
    if case.rules == "func":             # scalar GLSL built-ins, bit operators, ?: in conditions
        return Y.FUNC_YAML, Y.FUNC_IDS, Y.FUNC_MIX
    if case.rules == "rich":             # Left + Right rules, precedence, pos, rand.x: runs in the shader only with the patched emitter
        return Y.RICH_YAML, Y.RICH_IDS, Y.RICH_MIX
    kind, nm, nr, seed = case.rules.split(":")
    if kind == "synthLR":                # mirrored + RIGHT + LEFT rules (BASELINE configs[4] rule set): patched emitter
        return synthetic_rule_set(int(nm), int(nr), seed=int(seed))
    assert kind == "synthR"
    return synthetic_rule_set(int(nm), int(nr), seed=int(seed), kinds=("mirrored", "right", "mirrored", "right"))


def grid_for(case: Case) -> np.ndarray:
    _, ids, mix = rules_for(case)
    g = synthetic_grid(case.W, case.H, case.seed) if ids is None else synthetic_grid(case.W, case.H, case.seed, mix=mix, ids=ids)
    if case.grid == "unknown_ids":       # ids the rule set does not define read as NULL (gen/materials.glsl:79-86)
        y, x = np.mgrid[0:case.H, 0:case.W]
        g = g.copy()
        g[(x * 7 + y * 3) % 11 == 0] = 11
        g[(x * 5 + y * 13) % 17 == 0] = 200
        g[(x + y * 2) % 29 == 0] = 1     # NULL itself
        g[(x * 3 + y) % 31 == 0] = 2     # WALL inside the grid
    return g


def _h(*k) -> int:
    x = np.uint32(0x9E3779B9)
    for v in k:
        x = hashi(np.array([(int(x) * 2131 + int(v) * 461 + 1) & 0xFFFFFFFF], np.uint32))[0]
    return int(x)


def light0_for(case: Case) -> Optional[np.ndarray]:
    if not case.lighting:
        return None
    if case.light0 == "zero":
        return np.zeros((case.H, case.W, 4), np.float32)
    idx = np.arange(case.H * case.W * 4, dtype=np.uint64)
    h = hashi(((idx * np.uint64(2654435761) + np.uint64(case.seed * 97 + 5)) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    L = ((h >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)).reshape(case.H, case.W, 4)
    a = L[..., 3]
    a[(h.reshape(case.H, case.W, 4)[..., 0] & np.uint32(7)) == 0] = 0.0       # some alpha == 0 (the falloff special case)
    return np.ascontiguousarray(L)


def mods_for(case: Case, n_materials: int):
    """List (len steps) of MOD_DTYPE arrays (or None)."""
    if case.mods == "none":
        return None
    W, H = case.W, case.H
    out = []
    for s in range(case.steps):
        k = _h(case.seed, s, 1) % 7
        m = np.zeros(k, MOD_DTYPE)
        for i in range(k):
            m[i]["position"] = (_h(case.seed, s, i, 2) % (W + 8) - 4, _h(case.seed, s, i, 3) % (H + 8) - 4)
            m[i]["mod_shape"] = _h(case.seed, s, i, 4) % 2
            m[i]["mod_size"] = 1 + _h(case.seed, s, i, 5) % 11
            m[i]["mod_matID"] = _h(case.seed, s, i, 6) % n_materials        # includes 0 = explosion (EMPTY), 1 = NULL, 2 = WALL
        out.append(m)
    if case.mods == "config3":
        # BASELINE configs[3] (scripts/run_configs.py::frame_mods): per frame 4 brush stamps (CIRCLE / SQUARE, size 3..32,
        # a selectable material) + 1 explosion (CIRCLE of EMPTY, size 16..64)
        selectable = list(range(3, n_materials))
        out = []
        for s in range(case.steps):
            k = s + 1
            m = np.zeros(5, MOD_DTYPE)
            hv = hashi(np.arange(k * 16, k * 16 + 16, dtype=np.uint32))
            for i in range(4):
                m[i]["position"] = (int(hv[3 * i] % W), int(hv[3 * i + 1] % H))
                m[i]["mod_shape"] = int(hv[3 * i + 2] & 1)
                m[i]["mod_size"] = 3 + int((hv[3 * i + 2] >> 1) % 30)
                m[i]["mod_matID"] = selectable[int((hv[3 * i + 2] >> 8) % len(selectable))]
            m[4]["position"] = (int(hv[12] % W), int(hv[13] % H))
            m[4]["mod_shape"] = 0
            m[4]["mod_size"] = 16 + int(hv[14] % 49)
            m[4]["mod_matID"] = 0
            out.append(m)
        return out
    if case.mods == "huge":
        # mod_size up to INT_MAX ("fill everything"): the shader's tests are true for every cell of the grid
        I = 2 ** 31 - 1
        out = [np.zeros(0, MOD_DTYPE) for _ in range(case.steps)]
        for s, (px, py, shape, size, mat) in enumerate([(10, 10, 0, I, 3), (20, 5, 1, I, 5), (W // 2, H // 2, 0, 2 ** 30, 0),
                                                        (7, 9, 1, 2 ** 30 + 12345, 10), (30, 20, 0, 6, 4)]):
            m = np.zeros(1, MOD_DTYPE)
            m[0] = ((px, py), shape, size, mat, (0, 0, 0))
            out[2 * s] = m                     # every other frame, so the material also gets to move in between
        return out
    if case.mods == "edge":
        m = np.zeros(4, MOD_DTYPE)      # a mod_size == 0 record in the middle ends the scan (falling_sand.glsl:139-141)
        m[0] = ((20, 20), 0, 6, 3, (0, 0, 0)); m[1] = ((30, 30), 1, 0, 4, (0, 0, 0)); m[2] = ((40, 40), 1, 5, 4, (0, 0, 0))
        out[3] = m
        m = np.zeros(2, MOD_DTYPE)      # last match wins, and an unknown id (-> NULL) cancels it (falling_sand.glsl:176)
        m[0] = ((50, 20), 0, 6, 4, (0, 0, 0)); m[1] = ((50, 20), 0, 3, 77, (0, 0, 0))
        out[5] = m
        m = np.zeros(300, MOD_DTYPE)    # more than 256: the extras are dropped (simulation.rs:205)
        for i in range(300):
            m[i] = ((i % W, (i * 7) % H), i % 2, 1, 3 + (i % 8), (0, 0, 0))
        out[7] = m
        m = np.zeros(3, MOD_DTYPE)      # negative size never matches but does not end the scan; shape 2 matches nothing
        m[0] = ((10, 60), 0, -3, 4, (0, 0, 0)); m[1] = ((12, 60), 2, 5, 4, (0, 0, 0)); m[2] = ((70, 10), 1, 2, 5, (0, 0, 0))
        out[9] = m
    return out


# ---- engines ------------------------------------------------------------------------------------------------------
class RefEngine:
    """The reference's shader text compiled for the CPU (needs /root/reference or a prebuilt oracle/_ref)."""

    def __init__(self, yaml_text: str, is_default: bool):
        from oracle import oracle_lang
        from oracle.build_ref import load_ref
        if is_default:
            self.ref = load_ref()                      # the reference's own gen/materials.glsl + gen/rules.glsl
        else:
            # the emitter is pinned byte-exact on the reference's gen/*.glsl.  patched_left: identical text for rule sets
            # without LEFT rules; for LEFT rules the minimal emitter patch of oracle_lang.emit_glsl_rules (the reference
            # itself cannot compile them)
            res = oracle_lang.parse_string(yaml_text)
            self.ref = load_ref(oracle_lang.emit_glsl_materials(res), oracle_lang.emit_glsl_rules(res, patched_left=True))

    def start(self, W, H, lighting, grid, light0, frame0):
        self.ref.create(W, H)
        self.ref.upload_ids(grid)
        if light0 is not None:
            self.ref.upload_light(light0)
        self.ref.frame = frame0

    def step(self, mods=None):
        if mods is not None and len(mods):
            self.ref.push_modifications(mods)
        self.ref.step(1)

    def step_many(self, n):
        self.ref.step(n)

    def ids(self):
        return self.ref.download_ids()

    def light(self):
        return self.ref.download_light()

    def color(self):
        return self.ref.download_color()


class OracleEngine:
    """oracle/sand_oracle.c, literal per-cell form."""

    def __init__(self, yaml_text: str, is_default: bool):
        from oracle.build_oracle import load_oracle
        self.orc = load_oracle(None if is_default else yaml_text)
        self.want_color = False

    def start(self, W, H, lighting, grid, light0, frame0):
        self.cells = np.ascontiguousarray(grid, np.uint32).copy()
        self.L = None if light0 is None else light0.copy()
        self.frame = frame0
        self.col = None

    def step(self, mods=None):
        self.frame += 1
        self.cells, self.L, self.col = self.orc.step_cells(self.cells, self.frame, self.L, mods, want_color=self.want_color)

    def step_many(self, n):
        if self.L is None and not self.want_color and self.cells.size >= (1 << 22) and self.frame >= 1:
            # large grids, lighting off, no modifications: the in-place per-block form (4x less work; its equivalence
            # with the literal per-cell form is tested in tests/test_oracle_kat.py)
            self.frame = self.orc.run_blocks(self.cells, self.frame, n)
            return
        for _ in range(n):
            self.step()

    def ids(self):
        return self.cells

    def light(self):
        return self.L

    def color(self):
        return self.col


def sha_ids(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


def run_case(case: Case, engine_cls, want_light=True, want_color=False):
    """-> dict(ids={step: array}, light=array|None, color=array|None).  `engine_cls(yaml_text, is_default)`."""
    text, _, _ = rules_for(case)
    eng = engine_cls(text, case.rules == "default")
    eng.want_color = want_color
    n_materials = n_materials_of(case)
    grid, light0, mods = grid_for(case), light0_for(case), mods_for(case, n_materials)
    eng.start(case.W, case.H, case.lighting, grid, light0, case.frame0)
    ids, done = {}, 0
    for cp in case.all_checkpoints():
        if mods is None:
            eng.step_many(cp - done)
        else:
            for s in range(done, cp):
                eng.step(mods[s])
        done = cp
        ids[cp] = eng.ids().copy()
    return {"ids": ids, "light": eng.light() if (case.lighting and want_light) else None,
            "color": eng.color() if want_color else None}


def n_materials_of(case: Case) -> int:
    from oracle import oracle_lang
    text, _, _ = rules_for(case)
    return len(oracle_lang.parse_string(text).materials)
