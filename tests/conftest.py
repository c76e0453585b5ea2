import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libmtgv.so and the host harness once per session (no-op when up to date)."""
    import __graft_entry__ as g

    if os.environ.get("MTGV_SKIP_BUILD") != "1":
        g.build()
    yield
