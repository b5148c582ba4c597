import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are skipped, not failed (tests/test_cabi.py keeps the explicit check that
    the library refuses to work without one)."""
    try:
        from php_aho_corasick_b200 import native
        have = native.lib().acb200_device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (oracle/) if they are missing; the CUDA library is built by __graft_entry__.build()."""
    from oracle import pydriver
    if not pydriver.available("oracle"):
        pydriver.build(("liboracle_driver.so",))
    if not pydriver.available("reference") and os.path.isdir("/root/reference/src/multifast"):
        pydriver.build(("ref",))
    yield
