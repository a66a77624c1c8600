import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    from squarna_b200 import _lib
    if _lib.load().sqrn_device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    ctx = _lib.Context(0)
    yield ctx
    ctx.close()
