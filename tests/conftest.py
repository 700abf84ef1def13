import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build the native pieces if the tree is fresh (nvcc cross-compiles sm_100a without a GPU; ~2 min once)."""
    from onebit_b200 import build as _build
    try:
        _build.build()
        _build.build_torch_ops()
    except Exception as exc:  # surfaced by the tests that need the library
        print(f"[conftest] could not build libonebit_b200.so: {exc}")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
