import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the shared objects once per session (the .so files travel with the snapshot,
    so on the GPU box this is a no-op unless something is stale)."""
    import __graft_entry__ as entry

    entry.build_library()
    entry.build_hostsim()
    from oracle import oracle

    oracle.build()
