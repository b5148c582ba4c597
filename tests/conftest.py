import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (oracle/) if they are missing; the CUDA library is built by __graft_entry__.build()."""
    from oracle import pydriver
    if not pydriver.available("oracle"):
        pydriver.build(("liboracle_driver.so",))
    if not pydriver.available("reference") and os.path.isdir("/root/reference/src/multifast"):
        pydriver.build(("ref",))
    yield
