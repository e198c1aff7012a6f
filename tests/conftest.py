import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "code-gallery_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def dgx():
    import dgx as _dgx
    return _dgx


@pytest.fixture(scope="session")
def gpu_ctx(dgx):
    ctx = dgx.Context(0)
    yield ctx
    ctx.close()
