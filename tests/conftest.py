import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

REFERENCE = Path("/root/reference")   # present only in the build container, never on the GPU box
DEFAULT_YAML = REPO / "data" / "materials.yaml"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def default_yaml_text():
    return DEFAULT_YAML.read_text()


@pytest.fixture(scope="session")
def oracle():
    from oracle.build_oracle import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def native_lib():
    """The in-tree C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    from sandengine_b200 import build
    build.build(verbose=False)
    from sandengine_b200 import _capi
    return _capi.lib()


@pytest.fixture(scope="session")
def default_rules(native_lib):
    import sandengine_b200 as se
    return se.parse_path(DEFAULT_YAML)
