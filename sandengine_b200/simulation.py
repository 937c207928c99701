"""Host-side mirror of `sandengine_core::simulation` on top of the C ABI.

Reference: /root/reference/sandengine-core/src/simulation.rs
  Simulation::new(display, size)   :128-192   -> Simulation(rules, size, ...)
  Simulation::run()                :195-253   -> Simulation.run()   (one step), .step(n)
  Params                           :70-92     -> Params
  SimModification, MODSHAPE_*      :41-57     -> SimModification (numpy structured dtype, 32 B)
  fields `params`, `modifications` :122,125   -> same names
The GL textures `input_data`/`input_light` become device buffers owned by the native `se_sim`; the
host reaches them through upload_/download_ (headless readback) or `device_cells()` (zero-copy).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _capi
from .lang import ParsingResult, SandMaterial

MODSHAPE_CIRCLE = 0   # simulation.rs:41
MODSHAPE_SQUARE = 1   # simulation.rs:42
MAX_MODIFICATIONS = 256  # simulation.rs:43

MOD_DTYPE = np.dtype([("position", "<i4", (2,)), ("mod_shape", "<i4"), ("mod_size", "<i4"), ("mod_matID", "<i4"),
                      ("_pad4", "<i4", (3,))])
assert MOD_DTYPE.itemsize == 32


def SimModification(position, mod_shape=MODSHAPE_CIRCLE, mod_size=0, mod_matID=0) -> np.ndarray:
    """One `SimModification` record (simulation.rs:45-56)."""
    m = np.zeros((), dtype=MOD_DTYPE)
    m["position"] = position
    m["mod_shape"] = mod_shape
    m["mod_size"] = mod_size
    m["mod_matID"] = mod_matID
    return m


@dataclass
class Params:  # simulation.rs:70-92
    moveRight: bool = True
    mousePos: Tuple[float, float] = (0.0, 0.0)
    mousePressed: bool = False
    brushSize: int = 5
    brushMaterial: Optional[SandMaterial] = None
    time: float = 0.0
    frame: int = 0


class Simulation:
    """`Simulation` of the reference, backed by CUDA kernels on packed-u32 device cell buffers."""

    def __init__(self, rules: ParsingResult, size: Tuple[int, int], lighting: bool = False, device: int = 0,
                 row_begin: int = 0, row_end: int = 0, halo_rows: int = 0, temporal_block: int = 0,
                 running_census: bool = False, lit_strip: bool = False, fused_light: bool = False, device_share: int = 1):
        if not rules.compiled:
            raise ValueError("rules must be compiled (parse_string(..., compile=True))")
        self.rules = rules
        self.size = (int(size[0]), int(size[1]))
        self.lighting = bool(lighting)
        self.row_begin = int(row_begin)
        self.row_end = int(row_end) if row_end else self.size[1]
        self.params = Params()
        self.modifications: List[np.ndarray] = []
        flags = ((_capi.SE_FLAG_LIGHTING if lighting else 0) | (_capi.SE_FLAG_RUNNING_CENSUS if running_census else 0)
                 | (_capi.SE_FLAG_LIT_STRIP_EXPERIMENTAL if lit_strip else 0) | (_capi.SE_FLAG_FUSED_LIGHT_EXPERIMENTAL if fused_light else 0))   # the last two: accepted for compatibility, see the header
        prm = _capi.se_create_params(self.size[0], self.size[1], flags, device,
                                     self.row_begin, self.row_end, int(halo_rows), int(temporal_block), int(device_share))
        h = C.c_void_p()
        _capi.check(_capi.lib().se_sim_create(rules._h, C.byref(prm), C.byref(h)))
        self._h = h

    # -- stepping ------------------------------------------------------------------------------
    @property
    def owned_shape(self) -> Tuple[int, int]:
        return (self.row_end - self.row_begin, self.size[0])

    def _sync_frame_in(self):
        _capi.check(_capi.lib().se_sim_set_frame(self._h, int(self.params.frame)))

    def _sync_frame_out(self):
        f = C.c_int32()
        _capi.check(_capi.lib().se_sim_get_frame(self._h, C.byref(f)))
        self.params.frame = f.value

    def step(self, n_steps: int = 1) -> None:
        """n_steps x `run()`; pending `modifications` are consumed by the first step only (simulation.rs:246-252)."""
        L = _capi.lib()
        self._sync_frame_in()
        if self.modifications:
            arr = np.ascontiguousarray(np.stack([np.asarray(m, dtype=MOD_DTYPE) for m in self.modifications]).reshape(-1))
            _capi.check(L.se_sim_push_modifications(self._h, arr.ctypes.data, len(arr)))
            self.modifications.clear()
        _capi.check(L.se_sim_step(self._h, int(n_steps)))
        self._sync_frame_out()

    def run(self) -> None:
        """simulation.rs:195 -- one step."""
        self.step(1)

    def push_modifications(self, mods: np.ndarray) -> None:
        """Append an array of MOD_DTYPE records to the pending list (no per-record Python objects)."""
        arr = np.ascontiguousarray(mods, dtype=MOD_DTYPE).reshape(-1)
        if len(arr):
            _capi.check(_capi.lib().se_sim_push_modifications(self._h, arr.ctypes.data, len(arr)))

    def push_brush(self) -> None:
        """sandengine-core/src/lib.rs:59-67: one CIRCLE of brushSize / brushMaterial at mousePos * size."""
        p = self.params
        mat = p.brushMaterial.id if p.brushMaterial is not None else 0
        self.modifications.append(SimModification(
            [int(p.mousePos[0] * self.size[0]), int(p.mousePos[1] * self.size[1])], MODSHAPE_CIRCLE, int(p.brushSize), mat))

    # -- state ---------------------------------------------------------------------------------
    def upload_cells(self, cells: np.ndarray) -> None:
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        if cells.shape != self.owned_shape:
            raise ValueError(f"cells must have shape {self.owned_shape}")
        _capi.check(_capi.lib().se_sim_upload_cells(self._h, cells.ctypes.data))

    def download_cells(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.owned_shape, np.uint32)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.shape == self.owned_shape
        _capi.check(_capi.lib().se_sim_download_cells(self._h, out.ctypes.data))
        return out

    def upload_light(self, light: np.ndarray) -> None:
        light = np.ascontiguousarray(light, dtype=np.float32)
        if light.shape != self.owned_shape + (4,):
            raise ValueError("light must have shape (rows, width, 4)")
        _capi.check(_capi.lib().se_sim_upload_light(self._h, light.ctypes.data))

    def download_light(self) -> np.ndarray:
        out = np.empty(self.owned_shape + (4,), np.float32)
        _capi.check(_capi.lib().se_sim_download_light(self._h, out.ctypes.data))
        return out

    def download_color(self, rgba8: bool = False) -> np.ndarray:
        """The reference's `output_color` (operations.glsl:100-108) shaded on demand from the id buffer:
        float32 (rows, width, 4), or packed RGBA8 uint32 (rows, width) for headless readback."""
        if rgba8:
            out = np.empty(self.owned_shape, np.uint32)
            _capi.check(_capi.lib().se_sim_download_color(self._h, None, out.ctypes.data))
        else:
            out = np.empty(self.owned_shape + (4,), np.float32)
            _capi.check(_capi.lib().se_sim_download_color(self._h, out.ctypes.data, None))
        return out

    def save_color_png(self, path) -> None:
        """Headless frame dump: the reference's `output_color` as an 8-bit RGBA PNG (render.py)."""
        from .render import write_png
        write_png(path, self.download_color(rgba8=True))

    def upload_cells_ptr(self, host_ptr: int) -> None:
        _capi.check(_capi.lib().se_sim_upload_cells(self._h, C.c_void_p(host_ptr)))

    def download_cells_ptr(self, host_ptr: int) -> None:
        _capi.check(_capi.lib().se_sim_download_cells(self._h, C.c_void_p(host_ptr)))

    def device_cells(self) -> Tuple[int, int]:
        p, pitch = C.c_void_p(), C.c_size_t()
        _capi.check(_capi.lib().se_sim_device_cells(self._h, C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    def census(self) -> np.ndarray:
        out = np.zeros(256, np.uint64)
        _capi.check(_capi.lib().se_sim_census(self._h, out.ctypes.data))
        return out

    def checksum(self) -> int:
        """Sharding-independent checksum of the owned rows (se_sim_checksum): the strips' values add up mod 2^64."""
        v = C.c_uint64()
        _capi.check(_capi.lib().se_sim_checksum(self._h, C.byref(v)))
        return v.value

    def census_async(self, host_ptr: int) -> None:
        """Enqueue a census whose 256 x uint64 result lands at `host_ptr` (pinned memory) after census_wait();
        it runs on a side stream concurrently with later steps."""
        _capi.check(_capi.lib().se_sim_census_async(self._h, C.c_void_p(host_ptr)))

    def census_wait(self) -> None:
        _capi.check(_capi.lib().se_sim_census_wait(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        _capi.check(_capi.lib().se_sim_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        _capi.check(_capi.lib().se_sim_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        n = C.c_uint64()
        _capi.check(_capi.lib().se_sim_launch_count(self._h, C.byref(n)))
        return n.value

    # -- strips ----------------------------------------------------------------------------------
    def ipc_export(self):
        buf = (C.c_ubyte * 128)()
        lr, gt, gb = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _capi.check(_capi.lib().se_sim_ipc_export(self._h, buf, C.byref(lr), C.byref(gt), C.byref(gb)))
        return bytes(buf), lr.value, gt.value, gb.value

    def ipc_attach(self, which: int, handles: bytes, local_rows: int, ghost_top: int, ghost_bottom: int):
        buf = (C.c_ubyte * 128).from_buffer_copy(handles.ljust(128, b"\0"))
        _capi.check(_capi.lib().se_sim_ipc_attach(self._h, which, buf, local_rows, ghost_top, ghost_bottom))

    def ipc_export_light(self) -> bytes:
        buf = (C.c_ubyte * 128)()
        _capi.check(_capi.lib().se_sim_ipc_export_light(self._h, buf))
        return bytes(buf)

    def ipc_attach_light(self, which: int, handles: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(handles.ljust(128, b"\0"))
        _capi.check(_capi.lib().se_sim_ipc_attach_light(self._h, which, buf))

    def attach_local(self, which: int, neighbour: "Simulation"):
        """Neighbour strip owned by this process (0 = above, 1 = below)."""
        _capi.check(_capi.lib().se_sim_attach_local(self._h, which, neighbour._h))

    def halo_push(self) -> None:
        _capi.check(_capi.lib().se_sim_halo_push(self._h))

    def halo_exchange_async(self) -> None:
        """Device-ordered ghost-row exchange (no host synchronisation); lock-step on every strip."""
        _capi.check(_capi.lib().se_sim_halo_exchange_async(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            _capi.lib().se_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
