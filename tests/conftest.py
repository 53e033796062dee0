import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library must exist for every test session (built by __graft_entry__.build())."""
    so = os.path.join(ROOT, "lidarcrafter_b200", "csrc", "libb200lidar.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    yield
