"""Regenerates ref_shader_goldens.json + ref_shader_arrays.npz: outputs of the REFERENCE'S OWN compute shader
(/root/reference/shaders/compute/falling_sand.glsl + includes, compiled for the CPU by oracle/build_ref.py) for the
seeded inputs of tests/ref_cases.py.  Unlike oracle_state_sha256.json these are reference outputs: they pin the C
oracle (tests/test_ref_shader.py, no /root/reference needed) and the CUDA path (tests/test_gpu_ref_goldens.py).

Run in the build container (needs /root/reference):   python tests/golden/make_ref_shader_goldens.py [case names...]
"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import ref_cases as R  # noqa: E402
from oracle import build_ref  # noqa: E402

PROBES = ((0, 0), (1, 3), (5, 5), (8, 8), (-1, -1))   # (y, x), negative = from the end


def main():
    assert build_ref.reference_available(), "needs /root/reference"
    meta, arrays = {}, {}
    only = set(sys.argv[1:])            # optional: regenerate just these cases and merge them into the existing fixtures
    if only:
        assert only <= set(R.CASE_BY_NAME), only - set(R.CASE_BY_NAME)
        meta = json.loads(R.GOLDEN_JSON.read_text())["cases"]
        arrays = dict(np.load(R.GOLDEN_NPZ))
    for case in R.CASES:
        if only and case.name not in only:
            continue
        t = time.time()
        out = R.run_case(case, R.RefEngine, want_color=case.store_color)
        last = out["ids"][case.steps]
        m = {"ids_sha256": {str(k): R.sha_ids(v) for k, v in out["ids"].items()},
             "histogram": np.bincount(last.ravel()).tolist(),
             "init_sha256": R.sha_ids(R.grid_for(case))}
        if case.lighting:
            L = out["light"]
            m["light_sha256"] = hashlib.sha256(np.ascontiguousarray(L, np.float32).tobytes()).hexdigest()
            m["light_sum"] = [float(x) for x in L.astype(np.float64).sum(axis=(0, 1))]
            m["light_probes"] = {f"{y},{x}": [float(v) for v in L[y, x]] for (y, x) in PROBES if abs(y) < case.H and abs(x) < case.W}
            if case.store_light == "full":
                arrays[case.name + "/light"] = L
            elif case.store_light == "sub8":
                arrays[case.name + "/light_sub8"] = np.ascontiguousarray(L[::8, ::8])
        if case.store_color:
            arrays[case.name + "/color"] = out["color"]
        meta[case.name] = m
        print(f"{case.name}: {time.time() - t:.1f} s", flush=True)
    meta = {c.name: meta[c.name] for c in R.CASES}          # keep the order of ref_cases.CASES
    doc = {"_what": "outputs of the reference's own compute shader compiled for the CPU (oracle/build_ref.py); inputs: tests/ref_cases.py",
           "_shader_sha256": hashlib.sha256(build_ref.shader_source().encode()).hexdigest(),
           "cases": meta}
    R.GOLDEN_JSON.write_text(json.dumps(doc, indent=1) + "\n")
    np.savez_compressed(R.GOLDEN_NPZ, **arrays)
    print("wrote", R.GOLDEN_JSON, R.GOLDEN_NPZ, R.GOLDEN_NPZ.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
