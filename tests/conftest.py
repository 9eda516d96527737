import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_sessionstart(session):
    """Make sure the in-tree CUDA extension exists (nvcc cross-compiles without a GPU) so that a fresh
    checkout can run the suite; a stale library is rebuilt.  Building is skipped silently when nvcc is
    absent -- the tests that need the library then fail with its own "build me first" error."""
    try:
        from eagle_b200.build import build_native
        build_native()
    except Exception as e:  # noqa: BLE001
        print(f"[conftest] could not build libeagle_b200.so: {e}")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
