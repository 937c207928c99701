"""Regenerates reference_gen_sha256.json from /root/reference (build container only)."""
import hashlib
import json
from pathlib import Path

REF = Path("/root/reference")
FILES = ["shaders/compute/gen/materials.glsl", "shaders/compute/gen/rules.glsl", "shaders/compute/gen/falling_sand.glsl", "data/materials.yaml"]

if __name__ == "__main__":
    out = {"_comment": "SHA-256 / byte length of the reference's checked-in generated shader parts (shaders/compute/gen/*.glsl) and of its data/materials.yaml. Produced by tests/golden/make_reference_hashes.py in the build container; the files themselves are not copied."}
    for f in FILES:
        b = (REF / f).read_bytes()
        out[f] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
    (Path(__file__).parent / "reference_gen_sha256.json").write_text(json.dumps(out, indent=2) + "\n")
