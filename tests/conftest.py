import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One c2g_context on cuda:0 for the GPU tests.  No fallback: if the library or the device is
    missing the tests fail (they are only selected with -m gpu)."""
    from critic2_b200 import capi

    c = capi.Context(0)
    yield c
    c.close()
