"""ORACLE build + ctypes binding (test infrastructure, not product code).

`load_oracle(yaml_text)` generates rules_gen.h for the rule set with oracle_lang.emit_c_rules(),
compiles oracle/sand_oracle.c against it (gcc -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp) into
oracle/_build/<sha16>/liboracle.so and returns an `Oracle` wrapper.  Nothing here is imported by the
product package; the product path never falls back to it.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

from . import oracle_lang

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"
REPO = HERE.parent
DEFAULT_YAML = REPO / "data" / "materials.yaml"


class SimModification(C.Structure):  # simulation.rs:45-56 (32 bytes)
    _fields_ = [("position", C.c_int32 * 2), ("mod_shape", C.c_int32), ("mod_size", C.c_int32),
                ("mod_matID", C.c_int32), ("_pad4", C.c_int32 * 3)]


MOD_DTYPE = np.dtype([("position", "<i4", (2,)), ("mod_shape", "<i4"), ("mod_size", "<i4"),
                      ("mod_matID", "<i4"), ("_pad4", "<i4", (3,))])
assert MOD_DTYPE.itemsize == 32 and C.sizeof(SimModification) == 32


def _compile(yaml_text: str) -> Path:
    res = oracle_lang.parse_string(yaml_text)
    gen = oracle_lang.emit_c_rules(res)
    src = (HERE / "sand_oracle.c").read_bytes()
    # x86-64-v3 (AVX2), not -march=native: the .so is built in the build container and travels to the GPU box, whose
    # host CPU may lack this one's extensions (AVX-512 / AMX here)
    flags = ["-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp"]
    key = hashlib.sha256(gen.encode() + b"\0" + src + b"\0" + " ".join(flags).encode()).hexdigest()[:16]
    out_dir = BUILD / key
    so = out_dir / "liboracle.so"
    if so.exists():
        return so
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / "rules_gen.h").write_text(gen)
    tmp = out_dir / f"liboracle.{os.getpid()}.tmp.so"
    cmd = ["gcc", *flags, "-shared", "-fPIC",
           "-std=gnu11", "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-I", str(out_dir),
           str(HERE / "sand_oracle.c"), "-o", str(tmp), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle compile failed:\n{r.stderr}")
    os.replace(tmp, so)
    return so


class Oracle:
    """Host-side stepping contract of simulation.rs:195-253 on top of the C restatement."""

    def __init__(self, yaml_text: str | None = None):
        if yaml_text is None:
            yaml_text = DEFAULT_YAML.read_text()
        self.parsed = oracle_lang.parse_string(yaml_text)
        self.lib = C.CDLL(str(_compile(yaml_text)))
        u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
        self.lib.so_step_cells.argtypes = [u32p, u32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int]
        self.lib.so_step_cells.restype = None
        self.lib.so_step_blocks_inplace.argtypes = [u32p, C.c_int, C.c_int, C.c_int]
        self.lib.so_step_blocks_inplace.restype = None
        self.lib.so_step_blocks_strip.argtypes = [u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.so_step_blocks_strip.restype = None
        self.lib.so_run_blocks.argtypes = [u32p, C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.so_run_blocks.restype = C.c_int
        self.lib.so_hash43.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32 * 4),
                                       C.POINTER(C.c_float * 4)]
        self.lib.so_hashi.argtypes = [C.c_uint32]
        self.lib.so_hashi.restype = C.c_uint32
        self.n_materials = self.lib.so_n_materials()

    # -- known-answer helpers --------------------------------------------------------------
    def hash43(self, px, py, pz):
        seed = C.c_uint32()
        lanes = (C.c_uint32 * 4)()
        r = (C.c_float * 4)()
        self.lib.so_hash43(px, py, pz, C.byref(seed), C.byref(lanes), C.byref(r))
        return seed.value, list(lanes), [np.float32(v) for v in r]

    def hashi(self, x):
        return self.lib.so_hashi(x & 0xFFFFFFFF)

    # -- one dispatch (literal per-cell form) ------------------------------------------------
    def step_cells(self, cells, frame, light=None, mods=None, want_color=False):
        """cells: (H, W) uint32; frame: value after the host increment; returns (cells', light', color')."""
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        H, W = cells.shape
        out = np.empty_like(cells)
        lin = lout = None
        if light is not None:
            lin = np.ascontiguousarray(light, dtype=np.float32)
            assert lin.shape == (H, W, 4)
            lout = np.empty_like(lin)
        col = np.empty((H, W, 4), np.float32) if want_color else None
        n_mods, mptr = 0, None
        if mods is not None and len(mods):
            mods = np.ascontiguousarray(mods, dtype=MOD_DTYPE)
            n_mods, mptr = len(mods), mods.ctypes.data
        self.lib.so_step_cells(cells, out, lin.ctypes.data if lin is not None else None,
                               lout.ctypes.data if lout is not None else None,
                               col.ctypes.data if col is not None else None, W, H, int(frame), mptr, n_mods)
        return out, lout, col

    def step_blocks_inplace(self, cells, frame):
        assert cells.dtype == np.uint32 and cells.flags.c_contiguous
        H, W = cells.shape
        self.lib.so_step_blocks_inplace(cells, W, H, int(frame))

    def step_blocks_strip(self, cells, gy0, Hg, frame):
        """Local rows [gy0, gy0 + len(cells)) of a grid of Hg rows (ghost-row schedule emulation only)."""
        assert cells.dtype == np.uint32 and cells.flags.c_contiguous
        Hl, W = cells.shape
        self.lib.so_step_blocks_strip(cells, W, Hl, int(gy0), int(Hg), int(frame))

    def run_blocks(self, cells, frame, n_steps):
        """In-place, lighting off, no modifications. Returns the final frame."""
        assert cells.dtype == np.uint32 and cells.flags.c_contiguous
        H, W = cells.shape
        return self.lib.so_run_blocks(cells, W, H, int(frame), int(n_steps))

    def run(self, cells, frame, n_steps, light=None, mods_per_step=None, blocks=False):
        """simulation.rs:195-253: frame += 1 before each dispatch; modifications are consumed by exactly
        one step.  mods_per_step: optional list (len n_steps) of MOD_DTYPE arrays.  Returns (cells, light, frame)."""
        cells = np.ascontiguousarray(cells, dtype=np.uint32).copy()
        for s in range(n_steps):
            frame += 1
            m = mods_per_step[s] if mods_per_step is not None else None
            if blocks and light is None and (m is None or len(m) == 0) and frame != 1:
                self.step_blocks_inplace(cells, frame)
            else:
                cells, light, _ = self.step_cells(cells, frame, light, m)
        return cells, light, frame


_CACHE: dict = {}


def load_oracle(yaml_text: str | None = None) -> Oracle:
    key = hashlib.sha256((yaml_text or "<default>").encode()).hexdigest()
    if key not in _CACHE:
        _CACHE[key] = Oracle(yaml_text)
    return _CACHE[key]


if __name__ == "__main__":
    o = load_oracle()
    print("oracle built:", o.lib._name, "materials:", o.n_materials)
