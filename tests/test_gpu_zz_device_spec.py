"""The RasterSpec an unmodified Gudni derives on this GPU (determineRasterSpec, OpenCL/Setup.hs:71-87, from the
OpenCL device's limits: 1,024-pixel tiles, 1,024 threads per tile, MAXTHRESHOLDS 2,853 — see
tests/test_reference_pin.py) through the CUDA path, tiles binned on the GPU (level 2) and by the oracle
(level 1): images and threshold totals identical to the oracle's.  The canonical spec of the benchmarks is
smaller (G = 256); this is the one the reference itself would run with here.  (Per-thread taps and tile lists
under this spec are compared in tests/test_kernels_emulated.py.)"""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec
from gudni_b200.raster import setup_rasterizer
from oracle import oracle

pytestmark = pytest.mark.gpu

DEVICE_SPEC = RasterSpec(1024, 1024, 1024, 2853, 2851, 127)


def test_levels_1_and_2():
    scene = scenes.mixed_bag(200, 1300, 900, 11)
    ref = oracle.render(scene, DEVICE_SPEC, taps=False)
    r = setup_rasterizer(spec=DEVICE_SPEC)
    try:
        img, stats = r.raster_scene(0, scene)
        assert np.array_equal(img, ref.image)
        assert stats.n_thresholds == ref.total_thresholds and stats.n_overflow_threads == 0
        img, stats = r.queue_raster_jobs(1, scene, ref.jobs)
        assert np.array_equal(img, ref.image)
        assert stats.n_thresholds == ref.total_thresholds
    finally:
        r.close()
