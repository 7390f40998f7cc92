import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared libraries must exist (the driver runs __graft_entry__.build() first; this covers
    a bare `pytest` invocation in a fresh clone)."""
    so = os.path.join(ROOT, "semantic_slam_b200", "libssb.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(so) and os.path.exists(orc)):
        import __graft_entry__ as ge
        ge.build()
    yield
