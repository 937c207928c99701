"""ORACLE / TEST INFRASTRUCTURE — the reference's OWN compute shader, compiled for the CPU (oracle/_ref/).

The reference's hot path is one GLSL 4.30 compute shader (`shaders/compute/falling_sand.glsl` + includes; the
flattened copy `gen/falling_sand.glsl` is what `simulation.rs:130` loads).  There is no OpenGL >= 4.3 in this image,
so the shader cannot run as GLSL.  GLSL is, however, close enough to C++ that the shader text compiles with g++ once
the language built-ins exist (`oracle/glsl_shim.hpp`) and a handful of purely syntactic rewrites are applied.  This
module does exactly that:

    read  /root/reference/shaders/compute/falling_sand.glsl (+ its #includes, where they lie; never copied into the repo)
    ->    translate()  — the rewrites listed below, nothing semantic
    ->    oracle/_ref/<key>/ref_shader.cpp  = shim + translated shader + harness (oracle/ref_harness.inc); deleted after the compile
    ->    g++ -std=c++20 -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp  ->  oracle/_ref/<key>/libref_shader.so

The harness replays `Simulation::run` (simulation.rs:195-253) around the shader's `main()`: frame += 1, first
min(len, 256) modifications into the uniform block, one invocation per cell (the over-dispatched invocations of the
reference return at once, falling_sand.glsl:124-126, and are skipped), swap data/light, zero the used `mod_size`s.

Rewrites (all token-level; the statements, expressions and their order are the reference's):
  R1  comments removed; `#version`, `#extension`, `layout(local_size...) in;` dropped
  R2  `#include "x"` resolved relative to the including file (gen/materials.glsl and gen/rules.glsl can be overridden
      with text produced by the front end, to run OTHER rule sets through the reference's shader template)
  R3  `layout(...)`, `uniform`, `writeonly|readonly|volatile|coherent|restrict` removed from declarations; the
      members of a `uniform Block { ... };` become plain globals
  R4  parameter qualifiers: `inout T x` / `out T x` -> `T& x`, `in T x` -> `T x`
  R5  array types: `T[N] name` and `T name[N]` -> `glsl_array<T, N> name`
  R6  float literals get an `f` suffix (a GLSL literal without suffix is a 32-bit float)
  R7  `struct S {` gets `bool operator==(const S&) const = default;` (GLSL compares structs member-wise)
  R8  `void main()` -> `void shader_main()`
  R9  `case X: T v = e;` -> `case X: T v; v = e;` for scalar T (GLSL lets a later case label jump over an initialised
      declaration, falling_sand.glsl:147; C++ only over an uninitialised one)

Used by tests/ (to pin oracle/sand_oracle.c and to generate tests/golden/ref_shader_*.json) and by bench.py's CPU
baseline.  `/root/reference` does not exist on the GPU box: there only the prebuilt .so under oracle/_ref/ is loaded.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
import re
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_OUT = HERE / "_ref"
REFERENCE_ROOT = Path(os.environ.get("SE_REFERENCE_ROOT", "/root/reference"))
SHADER_DIR = REFERENCE_ROOT / "shaders" / "compute"
TOP = "falling_sand.glsl"
DEFAULT_KEY = "default"


def reference_available() -> bool:
    return (SHADER_DIR / TOP).is_file()


# ---- translation ------------------------------------------------------------------------------------------------
def _strip_comments(s: str) -> str:
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def _resolve_includes(path: Path, overrides: dict, depth: int = 0) -> str:
    if depth > 8:
        raise RuntimeError("include depth")
    out = []
    for line in path.read_text().splitlines(keepends=True):
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if not m:
            out.append(line)
            continue
        rel = m.group(1)
        if rel in overrides:
            out.append("\n" + overrides[rel] + "\n")
        else:
            out.append("\n" + _resolve_includes(path.parent / rel, overrides, depth + 1) + "\n")
    return "".join(out)


def translate(glsl: str) -> str:
    s = _strip_comments(glsl)                                                                     # R1
    s = re.sub(r"^\s*#\s*(version|extension)[^\n]*$", "", s, flags=re.M)
    s = re.sub(r"layout\s*\([^)]*\)\s*in\s*;", "", s)
    s = re.sub(r"\buniform\s+\w+\s*\{([^}]*)\}\s*;", r"\1", s)                                    # R3 (block)
    s = re.sub(r"layout\s*\([^)]*\)", "", s)
    s = re.sub(r"\b(uniform|writeonly|readonly|volatile|coherent|restrict)\b", "", s)
    s = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", s)                                  # R4
    s = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", s)
    s = re.sub(r"\b([A-Za-z_]\w*)\[(\d+)\]\s+(?=[A-Za-z_])", r"glsl_array<\1, \2> ", s)           # R5: T[N] name
    s = re.sub(r"\b([A-Za-z_]\w*)\s+([A-Za-z_]\w*)\[(\d+)\]\s*(=|;)", r"glsl_array<\1, \3> \2 \4", s)   # T name[N]
    s = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", s)   # R6
    s = re.sub(r"\bstruct\s+(\w+)\s*\{", r"struct \1 { bool operator==(const \1&) const = default;", s)  # R7
    s = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", s)                               # R8
    s = re.sub(r"(\bcase\s+\w+\s*:\s*)(float|int|uint|bool)\s+(\w+)\s*=", r"\1\2 \3; \3 =", s)          # R9
    return s


def shader_source(materials_glsl: str | None = None, rules_glsl: str | None = None) -> str:
    """The reference's shader with its includes resolved; gen/materials.glsl / gen/rules.glsl optionally replaced by
    text in the reference's generated format (e.g. from oracle_lang / the C++ front end's GLSL emitter)."""
    overrides = {}
    if materials_glsl is not None:
        overrides["gen/materials.glsl"] = materials_glsl
    if rules_glsl is not None:
        overrides["gen/rules.glsl"] = rules_glsl
    return _resolve_includes(SHADER_DIR / TOP, overrides)


def _key_for(materials_glsl, rules_glsl) -> str:
    if materials_glsl is None and rules_glsl is None:
        return DEFAULT_KEY
    return hashlib.sha256(((materials_glsl or "") + "\0" + (rules_glsl or "")).encode()).hexdigest()[:16]


def build(materials_glsl: str | None = None, rules_glsl: str | None = None, force: bool = False) -> Path:
    """Translate + compile; returns the .so.  Needs /root/reference (the sources are read where they lie)."""
    key = _key_for(materials_glsl, rules_glsl)
    out_dir = REF_OUT / key
    so = out_dir / "libref_shader.so"
    if not reference_available():
        if so.exists():
            return so
        raise FileNotFoundError(f"{SHADER_DIR / TOP} is not present and {so} was not prebuilt")
    glsl = shader_source(materials_glsl, rules_glsl)
    shim = (HERE / "glsl_shim.hpp").read_text()
    harness = (HERE / "ref_harness.inc").read_text()
    stamp = hashlib.sha256((glsl + "\0" + shim + "\0" + harness + "\0" + translate.__code__.co_code.hex() + "\0x86-64-v3").encode()).hexdigest()
    stamp_file = out_dir / "stamp.json"
    if so.exists() and not force and stamp_file.exists() and json.loads(stamp_file.read_text()).get("stamp") == stamp:
        return so
    out_dir.mkdir(parents=True, exist_ok=True)
    cpp = out_dir / "ref_shader.cpp"
    cpp.write_text('#include "glsl_shim.hpp"\n#include <vector>\n#include <omp.h>\nnamespace glsl {\n'
                   "// ---- translated reference shader (generated at build time, not committed) ----\n"
                   + translate(glsl) +
                   "\n// ---- harness (oracle/ref_harness.inc) ----\n" + harness + "\n}  // namespace glsl\n")
    tmp = out_dir / f"libref_shader.{os.getpid()}.tmp.so"
    cmd = ["g++", "-std=c++20", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-w",
           "-shared", "-fPIC", "-I", str(HERE), str(cpp), "-o", str(tmp)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if os.environ.get("SE_REF_KEEP_CPP") != "1":
        cpp.unlink()          # the translated text is derived from reference source: only the binary stays in oracle/_ref/
    if r.returncode != 0:
        raise RuntimeError("reference shader did not compile as C++:\n" + r.stderr[-6000:])
    os.replace(tmp, so)
    stamp_file.write_text(json.dumps({"stamp": stamp, "glsl_sha256": hashlib.sha256(glsl.encode()).hexdigest(),
                                      "source": str(SHADER_DIR / TOP), "cmd": " ".join(cmd[:-3])}, indent=1))
    return so


# ---- binding ----------------------------------------------------------------------------------------------------
MOD_DTYPE = np.dtype([("position", "<i4", (2,)), ("mod_shape", "<i4"), ("mod_size", "<i4"),
                      ("mod_matID", "<i4"), ("_pad4", "<i4", (3,))])


class RefShader:
    """`Simulation` of the reference (simulation.rs:97-253) on top of the translated shader.  One instance per .so
    (the shader's uniforms are globals of the library)."""

    def __init__(self, so: Path):
        self.lib = L = C.CDLL(str(so))
        u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        L.ref_create.argtypes = [C.c_int, C.c_int]
        L.ref_upload_ids.argtypes = [u32p]
        L.ref_download_ids.argtypes = [u32p]
        L.ref_upload_light.argtypes = [f32p]
        L.ref_download_light.argtypes = [f32p]
        L.ref_download_color.argtypes = [f32p]
        L.ref_set_frame.argtypes = [C.c_int]
        L.ref_push_mods.argtypes = [C.c_void_p, C.c_int]
        L.ref_step.argtypes = [C.c_int]
        L.ref_hash43.argtypes = [C.c_int, C.c_int, C.c_int, f32p]
        L.ref_set_threads.argtypes = [C.c_int]
        self.size = None

    def create(self, W: int, H: int):
        if self.lib.ref_create(W, H) != 0:
            raise MemoryError("ref_create")
        self.size = (W, H)
        return self

    def upload_ids(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        assert cells.shape == (self.size[1], self.size[0])
        self.lib.ref_upload_ids(cells)

    def download_ids(self):
        out = np.empty((self.size[1], self.size[0]), np.uint32)
        self.lib.ref_download_ids(out)
        return out

    def upload_light(self, light):
        light = np.ascontiguousarray(light, dtype=np.float32)
        assert light.shape == (self.size[1], self.size[0], 4)
        self.lib.ref_upload_light(light)

    def download_light(self):
        out = np.empty((self.size[1], self.size[0], 4), np.float32)
        self.lib.ref_download_light(out)
        return out

    def download_color(self):
        out = np.empty((self.size[1], self.size[0], 4), np.float32)
        self.lib.ref_download_color(out)
        return out

    @property
    def frame(self) -> int:
        return self.lib.ref_get_frame()

    @frame.setter
    def frame(self, v: int):
        self.lib.ref_set_frame(int(v))

    def push_modifications(self, mods):
        mods = np.ascontiguousarray(mods, dtype=MOD_DTYPE)
        self.lib.ref_push_mods(mods.ctypes.data, len(mods))

    def step(self, n: int = 1):
        self.lib.ref_step(int(n))

    def hash43(self, px, py, pz):
        out = np.empty(4, np.float32)
        self.lib.ref_hash43(px, py, pz, out)
        return out

    def set_threads(self, n: int):
        self.lib.ref_set_threads(int(n))


_CACHE: dict = {}


def load_ref(materials_glsl: str | None = None, rules_glsl: str | None = None) -> RefShader:
    key = _key_for(materials_glsl, rules_glsl)
    if key not in _CACHE:
        _CACHE[key] = RefShader(build(materials_glsl, rules_glsl))
    return _CACHE[key]


def prebuilt_default() -> Path | None:
    so = REF_OUT / DEFAULT_KEY / "libref_shader.so"
    return so if so.exists() else None


if __name__ == "__main__":
    print(build(force=True))
