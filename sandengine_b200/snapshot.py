"""Grid snapshots (the reference only has a "world loading/unloading" todo,
README.md:85-86).  The whole simulation state is (cells, light, frame): RAND is a pure function of
(block position, frame) (falling_sand.glsl:698), so a restored run continues bit-identically.

File = numpy .npz: meta (JSON: format, width, height, row_begin, row_end, halo_rows, frame, lighting, rules_sha256),
cells uint32 (owned rows, width), optional light float32 (owned rows, width, 4).  One file per strip for sharded runs: a
strip file holds the rows its rank OWNS; the ghost rows are not state (the first step after an upload fetches them from
the neighbours), so a sharded run is restored with `load_strip` on the same number of ranks -- `load` refuses a strip file
instead of building a strip without neighbours.
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
    """`sim`: a Simulation or a StripSimulation (its own strip)."""
    halo = int(getattr(getattr(sim, "plan", None), "halo_rows", 0))
    sim = getattr(sim, "sim", sim)
    meta = {"format": FORMAT, "width": sim.size[0], "height": sim.size[1], "row_begin": sim.row_begin, "row_end": sim.row_end,
            "halo_rows": halo, "frame": int(sim.params.frame), "lighting": bool(sim.lighting), "rules_sha256": rules_digest(sim.rules)}
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
        if meta["row_begin"] != 0 or meta["row_end"] not in (0, meta["height"]):
            raise ValueError(f"snapshot holds rows {meta['row_begin']}..{meta['row_end']} of a sharded run: restore it with load_strip() on every rank")
        sim = Simulation(rules, (meta["width"], meta["height"]), lighting=meta["lighting"], row_begin=meta["row_begin"],
                         row_end=meta["row_end"], **sim_kwargs)
        sim.upload_cells(z["cells"])
        if meta["lighting"]:
            sim.upload_light(z["light"])
        sim.params.frame = meta["frame"]
    return sim


def load_strip(rules, path, **strip_kwargs):
    """Restore one strip of a sharded run: call on every rank (torch.distributed initialised with as many ranks as the run
    that was saved) with that rank's file.  Returns a StripSimulation; its first step fetches the ghost rows."""
    from .distributed import StripSimulation
    with np.load(path) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("format") != FORMAT:
            raise ValueError(f"not a {FORMAT} file")
        if meta["rules_sha256"] != rules_digest(rules):
            raise ValueError("snapshot was taken with a different rule set")
        strip_kwargs.setdefault("halo_rows", meta.get("halo_rows") or 32)
        strip = StripSimulation(rules, (meta["width"], meta["height"]), lighting=meta["lighting"], **strip_kwargs)
        if (strip.row_begin, strip.row_end) != (meta["row_begin"], meta["row_end"]):
            rows = (strip.row_begin, strip.row_end)
            strip.close()
            raise ValueError(f"this rank owns rows {rows[0]}..{rows[1]}, the file holds {meta['row_begin']}..{meta['row_end']} (other world size or rank)")
        strip.upload_cells(z["cells"])
        if meta["lighting"]:
            strip.upload_light(z["light"])
        strip.params.frame = meta["frame"]
    return strip
