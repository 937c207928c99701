"""bench.py --impl reference: the JSON contract of the reference arm, on the CPU (no GPU involved in this arm)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def run_arm(extra_env=None, args=("--steps", "3", "--warmup", "1", "--size", "256")):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def check_line(line, kind):
    assert line["impl"] == "reference" and line["unit"] == "Gcell-updates/s" and line["higher_is_better"] is True
    assert line["steps"] == 3 and line["warmup"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert "16384^2" in line["metric"] and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == kind and cb["value"] == line["value"] and cb["cores"] >= 1 and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_the_reference_shader_when_prebuilt():
    from oracle import build_ref
    if build_ref.prebuilt_default() is None:
        if not build_ref.reference_available():
            pytest.skip("oracle/_ref was not prebuilt and /root/reference is absent")
        build_ref.build()
    out = run_arm()
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    check_line(line, "reference")
    assert "falling_sand.glsl" in line["cpu_baseline"]["sample"] and line["cpu_baseline"]["port"]["value"] > 0
    # the reference's shader is far slower than the hand-written port of the same algorithm: both are reported
    assert line["cpu_baseline"]["port"]["value"] > line["value"]


def test_reference_arm_falls_back_to_the_port(tmp_path):
    """Without oracle/_ref the arm times the oracle port and says so (kind "port")."""
    import shutil
    work = tmp_path / "repo"
    shutil.copytree(REPO, work, ignore=shutil.ignore_patterns(".git", "_ref", "gpurun_out", "profiles", "__pycache__", "*.npz", "build"))
    r = subprocess.run([sys.executable, str(work / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--size", "256"],
                       capture_output=True, text=True, timeout=300, env={k: v for k, v in os.environ.items() if k != "RANK"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    check_line(line, "port")
    assert "was not found" in line["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_print_nothing():
    assert run_arm({"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""
