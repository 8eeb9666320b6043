import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_libm():
    from oracle.oracle import Oracle
    return Oracle(port=False)


@pytest.fixture(scope="session")
def oracle_port():
    from oracle.oracle import Oracle
    return Oracle(port=True)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree native pieces exist (nvcc and gcc work without a GPU)."""
    import r2p2_b200
    r2p2_b200.build()
    from oracle import oracle as orc
    orc.build()
