"""The header is plain C (not just C++): examples/headless_step.c compiles with -std=c99 -pedantic -Werror against
include/sandengine_b200.h and links the shared library.  Without a GPU the program must fail loudly at se_sim_create
(exit code 3, "no CPU fallback"); with one it runs the simulation (marked gpu)."""
import ctypes as C
import os
import subprocess

import pytest

from conftest import REPO


@pytest.fixture(scope="module")
def example_binary(native_lib, tmp_path_factory):
    out = tmp_path_factory.mktemp("c_example") / "headless_step"
    libdir = REPO / "sandengine_b200"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(REPO / "include"),
                           str(REPO / "examples" / "headless_step.c"), "-L", str(libdir), "-lsandengine_b200",
                           f"-Wl,-rpath,{libdir}", "-o", str(out)])
    return out


def test_c_example_fails_loudly_without_a_device(example_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(example_binary), str(REPO / "data" / "materials.yaml"), "64", "64", "4"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.stdout, r.stderr)
    assert "rule set: " in r.stdout and "11 materials" in r.stdout          # the front end + NVRTC ran without a device
    assert "se_sim_create failed" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs(example_binary, oracle):
    import numpy as np

    import sandengine_b200 as se
    r = subprocess.run([str(example_binary), str(REPO / "data" / "materials.yaml"), "256", "256", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "frame 52:" in r.stdout
    n_sand = int(r.stdout.split("frame 52:")[1].split("sand cells")[0])
    # the same calls on the oracle: frame 1 clears, frame 2 takes the stamp, 50 plain steps
    m = np.zeros(1, se.MOD_DTYPE)
    m[0]["position"] = (128, 64); m[0]["mod_shape"] = se.MODSHAPE_CIRCLE; m[0]["mod_size"] = 20; m[0]["mod_matID"] = 3
    none = np.zeros(0, se.MOD_DTYPE)
    ref, _, frame = oracle.run(np.zeros((256, 256), np.uint32), 0, 52, mods_per_step=[none, m] + [none] * 50)
    assert frame == 52 and n_sand == int((ref == 3).sum()) > 1000


def test_abi_rejects_null_arguments(native_lib):
    """Every entry point validates its pointers: a negative status and a message, never a crash."""
    from sandengine_b200 import _capi
    L = _capi.lib()
    vp = C.c_void_p
    null = vp(None)
    out = vp()
    assert L.se_rules_compile_yaml(None, 0, C.byref(out)) == _capi.SE_ERR_INVALID_ARG
    assert L.se_rules_parse_only(b"x", 1, None) == _capi.SE_ERR_INVALID_ARG
    assert L.se_rules_destroy(null) == 0 and L.se_sim_destroy(null) == 0     # destroying nothing is fine
    for fn, args in [(L.se_rules_text, (null, 0, None, None)), (L.se_rules_cubin, (null, None, None)),
                     (L.se_rules_counts, (null, None, None, None)), (L.se_rules_material_id, (null, None, None)),
                     (L.se_rules_material, (null, 0, None, None, None, None, None, None)), (L.se_rules_rule, (null, 0, None, None, None, None)),
                     (L.se_sim_create, (null, None, None)), (L.se_sim_step, (null, 1)), (L.se_sim_push_modifications, (null, None, 0)),
                     (L.se_sim_set_frame, (null, 0)), (L.se_sim_get_frame, (null, None)), (L.se_sim_upload_cells, (null, None)),
                     (L.se_sim_download_cells, (null, None)), (L.se_sim_upload_light, (null, None)), (L.se_sim_download_light, (null, None)),
                     (L.se_sim_download_color, (null, None, None)), (L.se_sim_device_cells, (null, None, None)), (L.se_sim_census, (null, None)), (L.se_sim_checksum, (null, None)),
                     (L.se_sim_census_async, (null, None)), (L.se_sim_census_wait, (null,)), (L.se_sim_set_stream, (null, None)),
                     (L.se_sim_synchronize, (null,)), (L.se_sim_launch_count, (null, None)), (L.se_sim_ipc_export, (null, None, None, None, None)),
                     (L.se_sim_ipc_attach, (null, 0, None, 0, 0, 0)), (L.se_sim_ipc_export_light, (null, None)),
                     (L.se_sim_ipc_attach_light, (null, 0, None)), (L.se_sim_attach_local, (null, 0, null)),
                     (L.se_sim_halo_push, (null,)), (L.se_sim_halo_exchange_async, (null,))]:
        rc = fn(*args)
        assert rc == _capi.SE_ERR_INVALID_ARG, (fn.__name__, rc)
        assert L.se_last_error()
