"""The RasterSpec an unmodified Gudni derives on this GPU (determineRasterSpec, OpenCL/Setup.hs:71-87, from the
OpenCL device's limits: 1,024-pixel tiles, 1,024 threads per tile, MAXTHRESHOLDS 2,853 — see
tests/test_reference_pin.py) through the CUDA path, tiles binned by the oracle (level 1) and on the GPU
(level 2).  The canonical spec of the benchmarks is smaller (G = 256); this is the one the reference itself
would run with here."""
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec
from gudni_b200.raster import setup_rasterizer

from parity import level1_parity, level2_parity

pytestmark = pytest.mark.gpu

DEVICE_SPEC = RasterSpec(1024, 1024, 1024, 2853, 2851, 127)


@pytest.fixture(scope="module")
def device_rasterizer():
    r = setup_rasterizer(spec=DEVICE_SPEC)
    yield r
    r.close()


def test_level2(device_rasterizer):
    level2_parity(device_rasterizer, scenes.mixed_bag(200, 1300, 900, 11), spec=DEVICE_SPEC)


def test_level1(device_rasterizer):
    level1_parity(device_rasterizer, scenes.mixed_bag(200, 1300, 900, 11), spec=DEVICE_SPEC)
