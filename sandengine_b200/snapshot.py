"""Grid snapshots (the reference only has a "world loading/unloading" todo,
README.md:85-86).  The whole simulation state is (cells, light, frame): RAND is a pure function of
(block position, frame) (falling_sand.glsl:698), so a restored run continues bit-identically.

File = numpy .npz: meta (JSON: format, width, height, row_begin, row_end, frame, lighting, rules_sha256),
cells uint32 (rows, width), optional light float32 (rows, width, 4).  One file per strip for sharded runs.
"""
from __future__ import annotations

import hashlib
import json

import numpy as np

FORMAT = "sandengine_b200.snapshot.v1"


def rules_digest(rules) -> str:
    """Identity of a compiled rule set: SHA-256 of the text the reference would have generated for it."""
    return hashlib.sha256((rules.glsl_materials + "\0" + rules.glsl_rules).encode()).hexdigest()


def save(sim, path) -> None:
    meta = {"format": FORMAT, "width": sim.size[0], "height": sim.size[1], "row_begin": sim.row_begin, "row_end": sim.row_end,
            "frame": int(sim.params.frame), "lighting": bool(sim.lighting), "rules_sha256": rules_digest(sim.rules)}
    arrays = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), "cells": sim.download_cells()}
    if sim.lighting:
        arrays["light"] = sim.download_light()
    np.savez_compressed(path, **arrays)


def load(rules, path, **sim_kwargs):
    """Returns a new Simulation holding the snapshot's state. Raises ValueError when the rule set differs."""
    from .simulation import Simulation
    with np.load(path) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("format") != FORMAT:
            raise ValueError(f"not a {FORMAT} file")
        if meta["rules_sha256"] != rules_digest(rules):
            raise ValueError("snapshot was taken with a different rule set")
        sim = Simulation(rules, (meta["width"], meta["height"]), lighting=meta["lighting"], row_begin=meta["row_begin"],
                         row_end=meta["row_end"], **sim_kwargs)
        sim.upload_cells(z["cells"])
        if meta["lighting"]:
            sim.upload_light(z["light"])
        sim.params.frame = meta["frame"]
    return sim
