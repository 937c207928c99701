"""Tooling: differential campaign over RANDOM rule sets (sandengine_b200.synth_rules): the oracle's C restatement vs the
generated CUDA rule code compiled for the host (tests/emu), plus the transition table when the set is eligible; every
set also goes through NVRTC.   python scripts/diff_campaign.py <seed_lo> <seed_hi> [mirrored|small] [ref]
With `ref` (build container only: needs /root/reference) the REFERENCE'S OWN SHADER, compiled for the CPU by
oracle/build_ref.py with gen/materials.glsl + gen/rules.glsl emitted for the rule set (LEFT rules through the minimal
emitter patch), is the fourth party and is compared after every step as well.
Round 1: 36 mixed sets (LEFT / RIGHT rules, 6-29 materials) and 36 mirrored-only, table-eligible sets (5-12 materials),
80 steps each on a 48 x 40 grid: no mismatch; with `ref`: see DESIGN.md section 7."""
import sys, ctypes as C, subprocess, tempfile, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / 'tests'))
import numpy as np
import sandengine_b200 as se
from sandengine_b200.synth_rules import synthetic_rule_set
from sandengine_b200.grids import synthetic_grid
from oracle.build_oracle import load_oracle

def build_emu(rules, d):
    (d / "rules_gen.cuh").write_text(rules.cuda_header)
    so = d / "emu.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(d),
                           "-I", str(REPO / "sandengine_b200" / "csrc" / "kernels"), str(REPO / "tests" / "emu" / "host_emu.cpp"), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.emu_step_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.emu_step_lut_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    return lib

lo, hi = int(sys.argv[1]), int(sys.argv[2])
MIRRORED_ONLY = len(sys.argv) > 3 and sys.argv[3] == 'mirrored'
SMALL = MIRRORED_ONLY or (len(sys.argv) > 3 and sys.argv[3] == 'small')     # 'small' + env SE_LUT_LR=1: two-table sets
WITH_REF = 'ref' in sys.argv[3:]
if WITH_REF:
    import shutil
    from oracle import oracle_lang, build_ref
bad = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    n_mat = int(rng.integers(5, 13) if SMALL else rng.integers(6, 30)); n_rules = int(rng.integers(4, 24) if SMALL else rng.integers(4, 30))
    try:
        text, ids, mix = synthetic_rule_set(n_mat, n_rules, seed=seed, **({"kinds": ("mirrored",)} if MIRRORED_ONLY else {}))
    except Exception as e:
        print(seed, "generator failed", type(e).__name__, e); continue
    try:
        rules = se.parse_string(text)
    except se.SandEngineError as e:
        print(seed, "front end:", e.kind, str(e)[:200]); bad += 1; continue
    orc = load_oracle(text)
    with tempfile.TemporaryDirectory() as td:
        lib = build_emu(rules, Path(td))
        g = synthetic_grid(48, 40, seed, mix=mix, ids=ids)
        a, b = g.copy(), g.copy()
        c = g.copy()
        lut = lib.emu_lut_eligible() == 1 and lib.emu_build_lut() >= 0
        ref = None
        if WITH_REF:
            res = oracle_lang.parse_string(text)
            assert oracle_lang.emit_glsl_materials(res) == rules.glsl_materials and oracle_lang.emit_glsl_rules(res) == rules.glsl_rules
            mg, rg = oracle_lang.emit_glsl_materials(res), oracle_lang.emit_glsl_rules(res, patched_left=True)
            ref = build_ref.load_ref(mg, rg)
            ref.create(48, 40); ref.upload_ids(g); ref.frame = 1
        frame = 1
        ok = True
        for s in range(80):
            frame += 1
            orc.step_blocks_inplace(a, frame)
            lib.emu_step_inplace(b.ctypes.data, 48, 40, frame)
            if lut: lib.emu_step_lut_inplace(c.ctypes.data, 48, 40, frame)
            if ref is not None: ref.step(1)
            if not np.array_equal(a, b) or (lut and not np.array_equal(a, c)) or (ref is not None and not np.array_equal(a, ref.download_ids())):
                print(seed, f"MISMATCH at step {s+1} n_mat {n_mat} n_rules {n_rules} lut {lut}"); ok = False; bad += 1; break
        if ref is not None:      # one shared library per rule set: do not let a campaign fill oracle/_ref/
            shutil.rmtree(build_ref.REF_OUT / build_ref._key_for(mg, rg), ignore_errors=True)
        changed = int((a != g).sum())
        print(seed, "ok" if ok else "BAD", f"n_mat {n_mat} n_rules {n_rules} lut {lut} cells changed {changed}", flush=True)
print("bad:", bad)
