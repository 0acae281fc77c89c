import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _gpu_unavailable():
    """Why the -m gpu tests cannot run here (None if they can): no CUDA library built, or no sm_100 device."""
    try:
        from gudni_b200.raster import setup_rasterizer
        setup_rasterizer().close()
        return None
    except Exception as e:  # noqa: BLE001 - GudniError(NO_DEVICE) or a missing .so
        return str(e)


def pytest_collection_modifyitems(config, items):
    # plain `pytest` on a CPU box: the gpu tests are skipped instead of erroring out of the session.
    # With `-m gpu` the driver asked for them explicitly: there a missing device must fail loudly.
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or "gpu" in (config.getoption("-m") or "").replace("not gpu", ""):
        return
    why = _gpu_unavailable()
    if why is not None:
        skip = pytest.mark.skip(reason=f"needs a B200: {why}")
        for it in gpu_items:
            it.add_marker(skip)


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
