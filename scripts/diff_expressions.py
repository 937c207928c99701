"""Tooling (CPU, build container: needs /root/reference): differential test of the typed expression compiler.
Random rule sets whose conditions are random, well-typed scalar GLSL expressions (comparisons of densities / colours /
emissions / rand lanes / positions / frame / ids, arithmetic, the scalar built-ins, integer bit operators, ?:, vector
== / !=) are run
through four evaluators and compared after every step:
   the product's generated CUDA rule code compiled for the host (tests/emu)      -- what NVRTC compiles for the device
   the transition table, when the rule set is table-eligible
   the C oracle (oracle/sand_oracle.c + generated rules_gen.h)
   the reference's own shader compiled for the CPU (oracle/build_ref.py) with gen/*.glsl emitted for the rule set
python scripts/diff_expressions.py <seed_lo> <seed_hi> [eligible]      (eligible: mirrored-only rules without pos / frame / free rand,
                                                                     so that the transition table is built and compared too)
"""
import ctypes as C
import random
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import numpy as np  # noqa: E402
import sandengine_b200 as se  # noqa: E402
from oracle import build_ref, oracle_lang  # noqa: E402
from oracle.build_oracle import load_oracle  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402

# names that are not substrings of any word of the expression vocabulary (the reference replaces material names globally)
MATERIALS = """
materials:
  pebble:  {type: loose, color: [0.5, 0.5, 0.5], density: 2.0, extra_rules: [XR]}
  ashes:   {type: loose, color: [0.3, 0.25, 0.3], density: 1.25}
  ember:   {type: glowing, color: [0.4, 0.1, 0.1], density: 3.0, emission: [1.0, 0.6, 0.1, 0.9]}
  lantern: {type: glowing, color: [0.9, 0.9, 0.2], density: 3.0, emission: [0.2, 0.9, 0.2, 0.8]}
  sparkle: {type: shortlived, color: [1.0, 0.8, 0.2], density: 0.5, emission: [1.0, 0.8, 0.1, 0.7]}
"""
IDS = {"EMPTY": 0, "pebble": 3, "ashes": 4, "ember": 5, "lantern": 6, "sparkle": 7}
MIX = (("EMPTY", 0.45), ("pebble", 0.15), ("ashes", 0.15), ("ember", 0.09), ("lantern", 0.08), ("sparkle", 0.08))
TYPES = ["loose", "glowing", "shortlived", "EMPTY"]
MATS = ["pebble", "ashes", "ember", "lantern", "sparkle", "EMPTY"]
FLITS = ["0.0", "0.25", "0.5", "1.0", "1.25", "1.5", "2.0", "3.0", "0.1", "0.7", "0.9", "2.5", "9999.0", "0.004"]
COMP = "rgba"


class Gen:
    def __init__(self, rng, cells, eligible=False):
        self.r, self.cells, self.eligible = rng, cells, eligible      # eligible: only what a transition table can capture

    def cell(self):
        return self.r.choice(self.cells)

    def F(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.3:
            k = r.randrange(6)
            if self.eligible and k == 3: k = 1
            if k == 0: return f"{self.cell()}.mat.density"
            if k == 1: return f"{self.cell()}.mat.color.{r.choice(COMP)}"
            if k == 2: return f"{self.cell()}.mat.emission.{r.choice(COMP)}"
            if k == 3: return f"rand.{r.choice('xyzw')}"
            if k == 4: return r.choice(FLITS)
            return f"float({self.I(0)})"
        k = r.randrange(16)
        a, b = self.F(d - 1), self.F(d - 1)
        if k == 0: return f"({a} + {b})"
        if k == 1: return f"({a} - {b})"
        if k == 2: return f"({self.Fsmall(d - 1)} * {self.Fsmall(d - 1)})"
        if k == 3: return f"({a} / {r.choice(['2.0', '0.5', '3.0', '4.0', '0.1'])})"
        if k == 4: return f"abs({a})"
        if k == 5: return f"min({a}, {b})"
        if k == 6: return f"max({a}, {b})"
        if k == 7: return f"clamp({a}, {r.choice(['0.0', '0.25', '1.0'])}, {r.choice(['1.5', '2.0', '3.0'])})"
        if k == 8: return f"floor({a})"
        if k == 9: return f"ceil({a})"
        if k == 10: return f"fract({a})"
        if k == 11: return f"sign({a})"
        if k == 12: return f"step({a}, {b})"
        if k == 13: return f"sqrt(abs({a}))"
        if k == 14: return f"mod({a}, {r.choice(['2.0', '0.75', '3.0'])})"
        return f"({self.B(d - 1)} ? {a} : {b})"

    def Fsmall(self, d):          # factors that cannot blow up (products of 9999s overflow nothing, but keep magnitudes modest)
        r = self.r
        return r.choice([r.choice(FLITS[:12]) if self.eligible else f"rand.{r.choice('xyzw')}", r.choice(FLITS[:12]), f"{self.cell()}.mat.color.{r.choice(COMP)}",
                         f"{self.cell()}.mat.emission.{r.choice(COMP)}", f"min({self.cell()}.mat.density, 4.0)"])

    def I(self, d):
        r = self.r
        if d <= 0 or r.random() < 0.35:
            k = r.randrange(7)
            if self.eligible and k in (2, 3, 4): k = r.choice([0, 1, 5])
            if k == 0: return f"{self.cell()}.mat.id"
            if k == 1: return f"{self.cell()}.mat.type"
            if k == 2: return "pos.x"
            if k == 3: return "pos.y"
            if k == 4: return "frame"
            if k == 5: return str(r.randrange(8))
            return f"TYPE_{r.choice(TYPES)}"
        k = r.randrange(14)
        a, b = self.I(d - 1), self.I(d - 1)
        if k == 0: return f"({a} + {b})"
        if k == 1: return f"({a} - {b})"
        if k == 2: return f"({a} * {r.randrange(1, 5)})"
        if k == 3: return f"(abs({a}) % {r.randrange(2, 7)})"
        if k == 4: return f"(abs({a}) / {r.randrange(1, 5)})"
        if k == 5: return f"(abs({a}) >> {r.randrange(0, 4)})"
        if k == 6: return f"((abs({a}) & 1023) << {r.randrange(0, 4)})"
        if k == 7: return f"({a} & {r.randrange(1, 16)})"
        if k == 8: return f"({a} | {b})"
        if k == 9: return f"({a} ^ {b})"
        if k == 10: return f"(~{a})"
        if k == 11: return r.choice([f"min({a}, {b})", f"max({a}, {b})", f"abs({a})", f"sign({a})", f"clamp({a}, 0, {r.randrange(1, 9)})"])
        if k == 12: return f"int(min({self.cell()}.mat.density, 8.0) * {r.choice(['2.0', '0.5', '3.0'])})"
        return f"({self.B(d - 1)} ? {a} : {b})"

    def B(self, d):
        r = self.r
        cmp = r.choice(["<", "<=", ">", ">=", "==", "!="])
        if d <= 0 or r.random() < 0.3:
            k = r.randrange(9)
            if k == 8:                                                                          # vector == / != (one bool)
                sw = r.choice(["rgb", "rg", "rgba", "xyz"])
                args = [r.choice(["0.0", "0.5", "1.0", "0.3", "0.9"]) for _ in range(len(sw))] if r.random() < 0.5 else [r.choice(["0.0", "1.0", "0.5"])]
                return f"{self.cell()}.mat.{r.choice(['color', 'emission'])}.{sw} {r.choice(['==', '!='])} vec{len(sw)}({', '.join(args)})"
            if k == 0: return f"{self.cell()}.mat.density {cmp} {self.cell()}.mat.density"      # rank comparison
            if k == 1: return f"{self.cell()}.mat.density {cmp} {r.choice(FLITS)}"              # folded at code generation
            if k == 2: return f"rand.{'y' if self.eligible else r.choice('xyzw')} {r.choice(['<', '<=', '>', '>='])} {r.choice(FLITS[:12])}"   # integer threshold
            if k == 3: return f"isType_{r.choice(TYPES)}({self.cell()})"
            # (not EMPTY here: the reference replaces a compared material name GLOBALLY in the condition, rules.rs:246-260,
            #  which would turn isType_EMPTY / TYPE_EMPTY elsewhere in it into isType_MAT_EMPTY -- reproduced, but noise here)
            if k == 4: return f"{self.cell()}.mat {r.choice(['==', '!='])} {r.choice(MATS[:-1])}"
            if k == 5: return f"{self.I(1)} {cmp} {self.I(1)}"
            if k == 6: return f"{r.choice(FLITS)} {cmp} {self.cell()}.mat.density"
            return f"{self.F(1)} {cmp} {self.F(1)}"
        k = r.randrange(7)
        if k == 0: return f"{self.F(d - 1)} {cmp} {self.F(d - 1)}"
        if k == 1: return f"{self.I(d - 1)} {cmp} {self.I(d - 1)}"
        if k == 2: return f"({self.B(d - 1)} and {self.B(d - 1)})"
        if k == 3: return f"({self.B(d - 1)} or {self.B(d - 1)})"
        if k == 4: return f"not ({self.B(d - 1)})"
        if k == 5: return f"({self.B(d - 1)} ? {self.B(d - 1)} : {self.B(d - 1)})"
        return f"(({self.B(d - 1)}) == ({self.B(d - 1)}))"


ELIGIBLE = "eligible" in sys.argv[3:]


def make_rules(seed, eligible=None):
    eligible = ELIGIBLE if eligible is None else eligible
    rng = random.Random(seed)
    n = rng.randint(3, 6)
    rules, names = [], []
    for k in range(n):
        mirrored = eligible or rng.random() < 0.6
        g = Gen(rng, ["SELF", "RIGHT", "DOWN", "DOWNRIGHT"], eligible=eligible)
        name = f"q{k}"
        names.append(name)

        def action():
            if rng.random() < 0.6:
                a, b = rng.sample(["SELF", "RIGHT", "DOWN", "DOWNRIGHT"], 2)
                return f"SWAP {a} {b}"
            return f"SET {rng.choice(['SELF', 'RIGHT', 'DOWN', 'DOWNRIGHT'])} {rng.choice(MATS)}"

        txt = f"  {name}:\n    mirrored: {'true' if mirrored else 'false'}\n"
        if rng.random() < 0.3:
            txt += "    precondition: false\n"
        txt += f'    if: "{g.B(rng.randint(1, 3))}"\n    do: {action()}\n'
        if rng.random() < 0.5:
            txt += f"    probability: {rng.choice(['0.1', '0.5', '0.7', '0.004', '0.25'])}\n"
        if rng.random() < 0.4:
            txt += f'    else:\n      if: "{g.B(rng.randint(1, 2))}"\n      do: {action()}\n'
            if rng.random() < 0.5:
                txt += f"      probability: {rng.choice(['0.3', '0.9'])}\n"
        rules.append(txt)
    deal = {t: [] for t in ("loose", "glowing", "shortlived")}
    for nme in names[:-1]:
        deal[rng.choice(list(deal))].append(nme)
    types = "types:\n" + "".join(f"  {t}:\n    base_rules: [{', '.join(v)}]\n" for t, v in deal.items())
    return "rules:\n" + "".join(rules) + types + MATERIALS.replace("XR", names[-1])


def build_emu(rules, d):
    (d / "rules_gen.cuh").write_text(rules.cuda_header)
    so = d / "emu.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(d),
                           "-I", str(REPO / "sandengine_b200" / "csrc" / "kernels"), str(REPO / "tests" / "emu" / "host_emu.cpp"), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.emu_step_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.emu_step_lut_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    return lib


def run_one(seed, steps=40, w=40, h=28, eligible=None):
    text = make_rules(seed, eligible)
    try:
        rules = se.parse_string(text)                     # front end + CUDA code generation + NVRTC (sm_100a)
    except se.SandEngineError as e:
        return "refused", f"{e.kind}: {str(e)[:160]}", text
    res = oracle_lang.parse_string(text)
    mg, rg = oracle_lang.emit_glsl_materials(res), oracle_lang.emit_glsl_rules(res, patched_left=True)
    assert rules.glsl_rules == oracle_lang.emit_glsl_rules(res)
    try:
        ref = build_ref.load_ref(mg, rg)
    except RuntimeError as e:
        return "ref-compile", str(e)[-300:], text
    orc = load_oracle(text)
    g = synthetic_grid(w, h, seed, mix=MIX, ids=IDS)
    with tempfile.TemporaryDirectory() as td:
        lib = build_emu(rules, Path(td))
        lut = lib.emu_lut_eligible() == 1 and lib.emu_build_lut() >= 0
        ref.create(w, h); ref.upload_ids(g); ref.frame = 1
        a, b, c = g.copy(), g.copy(), g.copy()
        frame, status = 1, ("ok", "")
        for s in range(steps):
            frame += 1
            ref.step(1)
            orc.step_blocks_inplace(a, frame)
            lib.emu_step_inplace(b.ctypes.data, w, h, frame)
            if lut:
                lib.emu_step_lut_inplace(c.ctypes.data, w, h, frame)
            r = ref.download_ids()
            if not (np.array_equal(a, r) and np.array_equal(b, r) and (not lut or np.array_equal(c, r))):
                status = ("MISMATCH", f"step {s + 1}: oracle==ref {np.array_equal(a, r)}, generated code==ref {np.array_equal(b, r)}, "
                                      f"table==ref {np.array_equal(c, r) if lut else None}")
                break
        changed = int((r != g).sum())
    shutil.rmtree(build_ref.REF_OUT / build_ref._key_for(mg, rg), ignore_errors=True)
    return status[0], status[1] + f" lut {lut} changed {changed}", text


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    tally = {}
    for seed in range(lo, hi):
        st, info, text = run_one(seed)
        tally[st] = tally.get(st, 0) + 1
        print(seed, st, info, flush=True)
        if st in ("MISMATCH", "ref-compile", "refused"):
            (REPO / "gpurun_out").mkdir(exist_ok=True)
            (REPO / "gpurun_out" / f"expr_{st}_{seed}.yaml").write_text(text)
    print("tally", tally)
