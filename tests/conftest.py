import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The C-ABI library must exist; tests never fall back to anything else."""
    import __graft_entry__ as ge
    lib = os.path.join(ROOT, "epa-ng_b200", "libepa_b200.so")
    if not os.path.exists(lib):
        ge.build()
    return ge.load_package()
