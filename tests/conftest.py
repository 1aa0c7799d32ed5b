import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def engine_lib():
    """The built C-ABI library; building is part of the product (no fallback if it is absent)."""
    from manisdp_matlab_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from manisdp_matlab_b200.build import build
        build()
    return _lib.load()
