import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Harness + oracle are plain g++ builds (seconds); the CUDA library is built by
    __graft_entry__.build() and must already exist for the gpu tests."""
    from gudni_b200 import _build
    _build.build_host()
    _build.build_oracle()
    yield


@pytest.fixture(scope="session")
def rasterizer():
    from gudni_b200.raster import setup_rasterizer
    r = setup_rasterizer()
    yield r
    r.close()
