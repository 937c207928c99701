"""Host-side mirror of the `sandengine-lang` crate surface, on top of the C ABI.

Reference (all under /root/reference/sandengine-lang/src):
  parse_string / parse_path        parser.rs:93, lib.rs:10
  ParsingResult{rules,types,materials}   parser.rs:84-89
  create_glsl_from_parser          lib.rs:17  -> here `create_cuda_from_parser` (CUDA C + sm_100a cubin);
                                   the GLSL text is still available as a known-answer check.
Parsing, code generation and NVRTC compilation all happen in the native library
(csrc/lang/*.cpp); this module only marshals.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional

from . import _capi
from ._capi import SandEngineError  # noqa: F401  (re-export)


@dataclass
class SandMaterial:  # parser/materials.rs:10-29
    id: int
    name: str
    mattype: str
    color: tuple
    emission: tuple
    selectable: bool
    density: float


@dataclass
class SandRule:  # parser/rules.rs:25-44 (subset exposed over the ABI)
    name: str
    used: bool
    ruletype: str  # "Mirrored" | "Left" | "Right" as executed by this build
    precondition: Optional[str]


class ParsingResult:
    """Owns a native `se_rules` (parsed rule set + generated CUDA C + cubin)."""

    def __init__(self, handle: int, compiled: bool):
        self._h = C.c_void_p(handle)
        self.compiled = compiled
        L = _capi.lib()
        nr, nt, nm = C.c_int32(), C.c_int32(), C.c_int32()
        _capi.check(L.se_rules_counts(self._h, C.byref(nr), C.byref(nt), C.byref(nm)))
        self.n_types = nt.value
        self.materials: List[SandMaterial] = []
        for i in range(nm.value):
            name, tname = C.c_char_p(), C.c_char_p()
            dens, sel = C.c_float(), C.c_int32()
            col, em = (C.c_float * 4)(), (C.c_float * 4)()
            _capi.check(L.se_rules_material(self._h, i, C.byref(name), C.byref(tname), C.byref(dens), col, em, C.byref(sel)))
            self.materials.append(SandMaterial(i, name.value.decode(), tname.value.decode(), tuple(col), tuple(em), bool(sel.value), dens.value))
        self.rules: List[SandRule] = []
        for i in range(nr.value):
            name, pre = C.c_char_p(), C.c_char_p()
            used, kind = C.c_int32(), C.c_int32()
            _capi.check(L.se_rules_rule(self._h, i, C.byref(name), C.byref(used), C.byref(kind), C.byref(pre)))
            self.rules.append(SandRule(name.value.decode(), bool(used.value), ["Mirrored", "Left", "Right"][kind.value],
                                       pre.value.decode() if pre.value is not None else None))

    def _text(self, which: int) -> str:
        p, n = C.c_void_p(), C.c_size_t()
        _capi.check(_capi.lib().se_rules_text(self._h, which, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value).decode()

    @property
    def glsl_materials(self) -> str:
        """What the reference writes to shaders/compute/gen/materials.glsl (known-answer check only)."""
        return self._text(0)

    @property
    def glsl_rules(self) -> str:
        """What the reference writes to shaders/compute/gen/rules.glsl (known-answer check only)."""
        return self._text(1)

    @property
    def cuda_header(self) -> str:
        return self._text(2)

    @property
    def nvrtc_log(self) -> str:
        return self._text(3)

    @property
    def cubin(self) -> bytes:
        p, n = C.c_void_p(), C.c_size_t()
        _capi.check(_capi.lib().se_rules_cubin(self._h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def material_id(self, name: str) -> int:
        out = C.c_int32()
        _capi.check(_capi.lib().se_rules_material_id(self._h, name.encode(), C.byref(out)))
        return out.value

    @property
    def selectable_materials(self) -> List[SandMaterial]:  # sandengine-core/src/lib.rs:21-27
        return [m for m in self.materials if m.selectable]

    def __del__(self):
        try:
            if self._h:
                _capi.lib().se_rules_destroy(self._h)
                self._h = None
        except Exception:
            pass


def parse_string(text: str, compile: bool = True) -> ParsingResult:
    """parser.rs:93.  Raises SandEngineError with .kind in {MissingField, InvalidType, NotFound, NotRecognized, Yaml}.
    compile=True also generates CUDA C and builds the sm_100a cubin with NVRTC (no GPU needed)."""
    data = text.encode("utf-8")
    h = C.c_void_p()
    L = _capi.lib()
    fn = L.se_rules_compile_yaml if compile else L.se_rules_parse_only
    _capi.check(fn(data, len(data), C.byref(h)))
    return ParsingResult(h.value, compile)


def parse_path(path, compile: bool = True) -> ParsingResult:
    """sandengine-lang/src/lib.rs:10-13."""
    return parse_string(Path(path).read_text(), compile)


def create_cuda_from_parser(result: ParsingResult, out_dir) -> None:
    """Counterpart of create_glsl_from_parser (lib.rs:17): writes the generated artefacts to `out_dir`
    (rules_gen.cuh, sand_kernels.cubin) instead of shaders/compute/gen/*.glsl."""
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    (out / "rules_gen.cuh").write_text(result.cuda_header)
    if result.compiled:
        (out / "sand_kernels.cubin").write_bytes(result.cubin)
