"""Static resource budget of the default kernels (ptxas -v of the NVRTC translation unit for data/materials.yaml,
cross-compiled for sm_100a -- no GPU needed).  Guards the occupancy each kernel was designed and measured at
(DESIGN.md section 4): no local-memory spills anywhere, and registers within the launch bounds' implied budget
(65536 registers per SM / (threads per CTA x resident CTAs))."""
import re
import subprocess

import pytest

from conftest import REPO

# kernel -> (threads per CTA, resident CTAs per SM the design assumes)
BUDGET = {
    "se_step_tiles": (1024, 1),          # K1b: one persistent CTA of 1024 threads (two halves of 16 warps) per SM
    "se_step_lut_global": (512, 2),      # K1c
    "se_light": (256, 4),
    "se_step_inplace": (256, 8), "se_step_inplace_mods": (256, 6),
    "se_step_pingpong": (256, 8), "se_step_pingpong_mods": (256, 8),
    "se_shade": (256, 6), "se_fill_cells": (256, 8), "se_build_lut": (256, 8),
}


@pytest.fixture(scope="module")
def ptxas_table(native_lib):
    """Resource usage of the cubin NVRTC produced for the default rule set (what actually runs), plus the spill lines of
    the ptxas -v log of the ahead-of-time compile of the same translation unit with the same defines."""
    from sandengine_b200 import build
    build.inspect_default()
    res = subprocess.run(["cuobjdump", "-res-usage", str(REPO / "build" / "sand_kernels_default.nvrtc.cubin")], capture_output=True, text=True).stdout
    table = {}
    for m in re.finditer(r"Function (\w+):\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        table[m.group(1)] = dict(regs=int(m.group(2)), stack=int(m.group(3)), local=int(m.group(5)), spill_st=0, spill_ld=0)
    log = (REPO / "build" / "ptxas_default.log").read_text()
    for m in re.finditer(r"Compiling entry function '(\w+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
                         r".*?Used (\d+) registers", log, flags=re.S):
        assert m.group(1) in table
        table[m.group(1)].update(spill_st=int(m.group(3)), spill_ld=int(m.group(4)), regs_aot=int(m.group(5)))
    return table


def test_every_default_kernel_is_compiled_for_sm_100a(ptxas_table):
    assert set(BUDGET) <= set(ptxas_table), set(BUDGET) - set(ptxas_table)


@pytest.mark.parametrize("kernel", sorted(BUDGET))
def test_no_spills_and_registers_within_the_occupancy_budget(ptxas_table, kernel):
    t = ptxas_table[kernel]
    assert t["stack"] == 0 and t["local"] == 0 and t["spill_st"] == 0 and t["spill_ld"] == 0, (kernel, t)
    assert abs(t.get("regs_aot", t["regs"]) - t["regs"]) <= 4, (kernel, t, "the ahead-of-time inspect build differs from the NVRTC cubin")
    threads, ctas = BUDGET[kernel]
    assert t["regs"] * threads * ctas <= 65536, (kernel, t["regs"], "registers do not allow", ctas, "CTAs of", threads, "threads per SM")


def test_the_tile_kernel_uses_vector_loads_and_shared_memory():
    """SASS of the cubin NVRTC produced (what runs): 128-bit global loads/stores for the tile interior, LDS/STS for the
    sub-steps, no local memory traffic."""
    cubin = REPO / "build" / "sand_kernels_default.nvrtc.cubin"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "se_step_tiles", str(cubin)], capture_output=True, text=True).stdout
    assert "LDG.E.128" in sass and "STG.E.128" in sass
    assert re.search(r"\bLDS", sass) and re.search(r"\bSTS", sass)


def test_the_tile_kernel_keeps_local_memory_out_of_the_sub_step_loops():
    """The sub-step loops are the code between the first and the last 64-bit shared-memory access of the kernel (tile rows
    are read and written as 8-byte words only there and in the load/store phases): no LDL / STL in between the table reads."""
    cubin = REPO / "build" / "sand_kernels_default.nvrtc.cubin"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "se_step_tiles", str(cubin)], capture_output=True, text=True).stdout
    lines = [ln for ln in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
    votes = [i for i, ln in enumerate(lines) if "IDP.4A" in ln]
    assert votes, "the table index is two byte dot products"
    hot = lines[votes[0]:votes[-1] + 1]
    assert not any(re.search(r"\b(LDL|STL)\b", ln) for ln in hot)
