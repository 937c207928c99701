"""ctypes binding of include/sandengine_b200.h (the C ABI).  Fails loudly when the library is missing:
there is no Python / CPU fallback for the simulation path."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libsandengine_b200.so"

SE_OK = 0
SE_ERR_YAML = -1
SE_ERR_MISSING_FIELD = -2
SE_ERR_INVALID_TYPE = -3
SE_ERR_NOT_FOUND = -4
SE_ERR_NOT_RECOGNIZED = -5
SE_ERR_UNSUPPORTED = -6
SE_ERR_COMPILE = -7
SE_ERR_CUDA = -8
SE_ERR_INVALID_ARG = -9
SE_ERR_INTERNAL = -10

SE_FLAG_LIGHTING = 1
SE_FLAG_RUNNING_CENSUS = 2   # experimental, see the header
SE_FLAG_LIT_STRIP_EXPERIMENTAL = 4
SE_FLAG_FUSED_LIGHT_EXPERIMENTAL = 8
SE_MODSHAPE_CIRCLE = 0
SE_MODSHAPE_SQUARE = 1
SE_MAX_MODIFICATIONS = 256

EXPORTS = [
    "se_rules_compile_yaml", "se_rules_parse_only", "se_rules_destroy", "se_rules_text", "se_rules_cubin", "se_rules_counts",
    "se_rules_material", "se_rules_material_id", "se_rules_rule",
    "se_sim_create", "se_sim_destroy", "se_sim_step", "se_sim_push_modifications", "se_sim_set_frame", "se_sim_get_frame",
    "se_sim_upload_cells", "se_sim_download_cells", "se_sim_upload_light", "se_sim_download_light", "se_sim_download_color", "se_sim_device_cells",
    "se_sim_census", "se_sim_checksum", "se_sim_census_async", "se_sim_census_wait", "se_sim_set_stream", "se_sim_synchronize", "se_sim_launch_count",
    "se_sim_ipc_export", "se_sim_ipc_attach", "se_sim_ipc_export_light", "se_sim_ipc_attach_light", "se_sim_attach_local", "se_sim_halo_push", "se_sim_halo_exchange_async", "se_last_error", "se_version",
]


class se_modification(C.Structure):  # == simulation.rs:45-56
    _fields_ = [("position", C.c_int32 * 2), ("mod_shape", C.c_int32), ("mod_size", C.c_int32), ("mod_matID", C.c_int32),
                ("_pad4", C.c_int32 * 3)]


class se_create_params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("flags", C.c_uint32), ("device", C.c_int32),
                ("row_begin", C.c_uint32), ("row_end", C.c_uint32), ("halo_rows", C.c_uint32), ("temporal_block", C.c_uint32),
                ("device_share", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m sandengine_b200.build` (nvcc, sm_100a). "
            "sandengine_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
    P = C.POINTER
    L.se_last_error.restype = cp
    L.se_version.restype = cp
    L.se_rules_compile_yaml.argtypes = [cp, sz, P(vp)]
    L.se_rules_parse_only.argtypes = [cp, sz, P(vp)]
    L.se_rules_destroy.argtypes = [vp]
    L.se_rules_text.argtypes = [vp, C.c_int, P(vp), P(sz)]
    L.se_rules_cubin.argtypes = [vp, P(vp), P(sz)]
    L.se_rules_counts.argtypes = [vp, P(C.c_int32), P(C.c_int32), P(C.c_int32)]
    L.se_rules_material.argtypes = [vp, C.c_int32, P(cp), P(cp), P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_int32)]
    L.se_rules_material_id.argtypes = [vp, cp, P(C.c_int32)]
    L.se_rules_rule.argtypes = [vp, C.c_int32, P(cp), P(C.c_int32), P(C.c_int32), P(cp)]
    L.se_sim_create.argtypes = [vp, P(se_create_params), P(vp)]
    L.se_sim_destroy.argtypes = [vp]
    L.se_sim_step.argtypes = [vp, C.c_uint32]
    L.se_sim_push_modifications.argtypes = [vp, vp, C.c_uint32]
    L.se_sim_set_frame.argtypes = [vp, C.c_int32]
    L.se_sim_get_frame.argtypes = [vp, P(C.c_int32)]
    L.se_sim_upload_cells.argtypes = [vp, vp]
    L.se_sim_download_cells.argtypes = [vp, vp]
    L.se_sim_upload_light.argtypes = [vp, vp]
    L.se_sim_download_light.argtypes = [vp, vp]
    L.se_sim_download_color.argtypes = [vp, vp, vp]
    L.se_sim_device_cells.argtypes = [vp, P(vp), P(sz)]
    L.se_sim_census.argtypes = [vp, vp]
    L.se_sim_checksum.argtypes = [vp, vp]
    L.se_sim_census_async.argtypes = [vp, vp]
    L.se_sim_census_wait.argtypes = [vp]
    L.se_sim_set_stream.argtypes = [vp, vp]
    L.se_sim_synchronize.argtypes = [vp]
    L.se_sim_launch_count.argtypes = [vp, P(C.c_uint64)]
    L.se_sim_ipc_export.argtypes = [vp, vp, P(C.c_uint64), P(C.c_uint64), P(C.c_uint64)]
    L.se_sim_ipc_attach.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.se_sim_ipc_export_light.argtypes = [vp, vp]
    L.se_sim_ipc_attach_light.argtypes = [vp, C.c_int, vp]
    L.se_sim_attach_local.argtypes = [vp, C.c_int, vp]
    L.se_sim_halo_push.argtypes = [vp]
    L.se_sim_halo_exchange_async.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("se_last_error", "se_version"):
            fn.restype = C.c_int
    _lib = L
    return L


class SandEngineError(RuntimeError):
    """Raised for every non-zero se_status.  `.status` is the code, `.kind` the ParsingErr class name."""
    KINDS = {SE_ERR_YAML: "Yaml", SE_ERR_MISSING_FIELD: "MissingField", SE_ERR_INVALID_TYPE: "InvalidType",
             SE_ERR_NOT_FOUND: "NotFound", SE_ERR_NOT_RECOGNIZED: "NotRecognized", SE_ERR_UNSUPPORTED: "Unsupported",
             SE_ERR_COMPILE: "Compile", SE_ERR_CUDA: "Cuda", SE_ERR_INVALID_ARG: "InvalidArg", SE_ERR_INTERNAL: "Internal"}

    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status = status
        self.kind = self.KINDS.get(status, "Unknown")


def check(status: int) -> None:
    if status != SE_OK:
        raise SandEngineError(status, lib().se_last_error().decode("utf-8", "replace"))
