"""include/sandengine_b200.hpp -- the C++17 host-side mirror of sandengine_lang / sandengine_core::simulation -- built
with -Wall -Wextra -Werror and exercised by tests/cpp/reference_style_tests.cpp (written like the reference's own
tests/test_sandengine-lang.rs).  Without a GPU: the parser error classes, the default rule set through NVRTC, and the loud
failure of Simulation::new_; with one (marked gpu): the survey's state KAT, brush stamp and clear-frame through the class."""
import subprocess

import pytest

from conftest import DEFAULT_YAML, REPO


@pytest.fixture(scope="module")
def cpp_tests(native_lib, tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp_mirror") / "reference_style_tests"
    libdir = REPO / "sandengine_b200"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(REPO / "include"),
                           str(REPO / "tests" / "cpp" / "reference_style_tests.cpp"), "-L", str(libdir), "-lsandengine_b200",
                           f"-Wl,-rpath,{libdir}", "-o", str(out)])
    return out


def test_cpp_mirror_parser_and_loud_failure_without_a_device(cpp_tests):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([str(cpp_tests), str(DEFAULT_YAML)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok (0 failed)" in r.stdout


@pytest.mark.gpu
@pytest.mark.first_gpu_run
def test_cpp_mirror_simulation_on_the_device(cpp_tests):
    r = subprocess.run([str(cpp_tests), str(DEFAULT_YAML), "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok (0 failed)" in r.stdout
