"""Pins the restated oracle to the reference itself.

`oracle/_ref/libgudni_ref.so` is the reference's kernel file (src/Graphics/Gudni/OpenCL/Kernels.cl —
generateThresholds, sortThresholds, renderThresholds and every device function they call) compiled for the
host by oracle/refbuild/build_ref.py, IEEE f32 without contraction.  These tests run the restated oracle
(oracle/kernels_oracle.cpp) and the compiled reference on the same raster jobs and require identical
per-thread threshold counts, identical shape-bit counts and identical BGRA words: the restatement is the
reference's arithmetic, statement for statement.  They need the reference tree (or a prebuilt library) and
skip without it; the committed vectors in tests/golden/ carry the pin where the tree is absent.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec
from oracle import oracle

pytestmark = pytest.mark.skipif(oracle.reference_lib() is None,
                                reason="reference kernels not built (no /root/reference, no oracle/_ref)")

SPEC_64 = RasterSpec(max_tile_size=64, threads_per_tile=64, max_tiles_per_call=64, max_thresholds=512,
                     max_strands_per_tile=510)
SPEC_32 = RasterSpec(max_tile_size=32, threads_per_tile=32, max_tiles_per_call=32, max_thresholds=256,
                     max_strands_per_tile=254)


def both(scene, spec=None):
    kw = {} if spec is None else {"spec": spec}
    jobs = oracle.build_raster_jobs(scene, **kw)
    mine = oracle.raster_jobs(scene, jobs, **kw)
    ref = oracle.raster_jobs(scene, jobs, reference=True, **kw)
    assert ref.overflow_threads == 0 and mine.overflow_threads == 0
    return mine, ref


def assert_identical(mine, ref):
    assert mine.total_thresholds == ref.total_thresholds
    for a, b in zip(mine.n_thresholds, ref.n_thresholds):
        assert np.array_equal(a, b), f"threshold counts differ at threads {np.flatnonzero(a != b)[:8]}"
    for a, b in zip(mine.shape_bits, ref.shape_bits):
        assert np.array_equal(a, b), f"shape bits differ at threads {np.flatnonzero(a != b)[:8]}"
    bad = np.argwhere(mine.image != ref.image)
    assert len(bad) == 0, f"{len(bad)} pixels differ, first at (y,x)={bad[:5].tolist()}"


def test_library_exports():
    L = oracle.reference_lib()
    for sym in ("gudni_ref_raster_job", "gudni_ref_threads", "gudni_ref_set_threads"):
        assert hasattr(L, sym)
    assert isinstance(L, ctypes.CDLL) and L.gudni_ref_threads() >= 1


def test_golden_vectors_are_what_the_reference_produces_now():
    """The committed vectors are regenerated from the compiled reference and must not have drifted."""
    from golden.make_golden import SCENES, digest, render
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_hashes.json")))
    for name in SCENES:
        assert digest(render(name, reference=True)) == golden[name], name


CATALOGUE = [scenes.tiny_square, scenes.medium_square, scenes.full_rectangle, scenes.stack_of_squares,
             scenes.open_square, scenes.concentric_squares2, scenes.concentric_squares3,
             scenes.six_point_rectangle, scenes.hour_glass, scenes.translucent_stack]


@pytest.mark.parametrize("make", CATALOGUE, ids=lambda f: f.__name__)
def test_catalogue_scenes(make):
    assert_identical(*both(make()))


def test_coordinates_at_the_far_end_of_float32():
    """Rectangles 4e12 pixels across, and a coordinate turned into NaN or +-3e38: the reference's kernels terminate on
    all of them and the restatement produces the same bits.  (A point at infinity is the one thing they do not
    terminate on — DESIGN section 7 — so it is not run.)"""
    assert_identical(*both(scenes.huge_boxes()))
    for seed in range(8):         # rectangles and circles of 1 to 1e30 pixels across the canvas
        assert_identical(*both(scenes.far_shapes(12, 150, 110, 0xFA50 + seed)))
    for value in (np.nan, 3.0e38, -3.0e38):
        scene = scenes.medium_square()
        scene.geometry = scene.geometry.copy()
        scene.geometry.view(np.float32)[2 * 2] = value
        assert_identical(*both(scene))


@pytest.mark.parametrize("theta", [0.3, 0.4, 0.5, 0.625])
@pytest.mark.parametrize("size", [100, 512])
def test_s1_square(size, theta):
    assert_identical(*both(scenes.square(size, theta)))


@pytest.mark.parametrize("seed", range(6))
def test_random_rectangles(seed):
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(33, 400)), int(rng.integers(17, 300))
    assert_identical(*both(scenes.random_rectangles(int(rng.integers(20, 250)), w, h, seed)))


@pytest.mark.parametrize("spec", [None, SPEC_64, SPEC_32], ids=["G256", "G64", "G32"])
@pytest.mark.parametrize("seed", [3, 4])
def test_fuzzy_circles(seed, spec):
    rng = np.random.default_rng(2000 + seed)
    w, h = int(rng.integers(100, 600)), int(rng.integers(100, 400))
    assert_identical(*both(scenes.fuzzy_circles(int(rng.integers(100, 500)), w, h, 4, 40, seed), spec))


def test_deep_translucent_stacks():
    """Circles crowded enough that tiles split down towards the 8-pixel floor and stacks run deep."""
    assert_identical(*both(scenes.fuzzy_circles(3000, 256, 256, 5, 50, 77)))


def test_picture_substances():
    assert_identical(*both(scenes.picture_scene()))
    assert_identical(*both(scenes.picture_scene(320, 300, scale=2.0, flowers_size=(350, 200))))


def test_thin_rectangles():
    assert_identical(*both(scenes.thin_rectangles(40)))
    assert_identical(*both(scenes.thin_rectangles(64, width=96, thickness=0.15, skew=0.7)))


def test_paragraph_reduced():
    assert_identical(*both(scenes.s2(640, 300, lines=6)))


def test_s4b_full_size():
    """BASELINE.json's literal '100k curves at 3840x2160' (S4b: 6,250 circles): 744,366 thresholds,
    8.3 M pixels, every word equal."""
    mine, ref = both(scenes.s4b())
    assert_identical(mine, ref)
    assert mine.total_thresholds == 744366


@pytest.mark.parametrize("case", range(16))
def test_mixed_bag(case):
    """Free-form curves with random control points (knobs, self-intersections), slivers, holes, two-outline
    shapes, rotated rectangles, pictures at several scales, on awkward canvases, two RasterSpecs."""
    rng = np.random.default_rng(5000 + case)
    w, h = int(rng.integers(20, 700)), int(rng.integers(20, 500))
    scene = scenes.mixed_bag(int(rng.integers(1, 300)), w, h, 7000 + case)
    assert_identical(*both(scene, SPEC_64 if case % 3 == 0 else None))


DEVICE_SPEC = RasterSpec(max_tile_size=1024, threads_per_tile=1024, max_tiles_per_call=1024, max_thresholds=2853,
                         max_strands_per_tile=2851)


def test_the_spec_the_reference_derives_on_this_gpu():
    """determineRasterSpec (OpenCL/Setup.hs:71-87) on the B200's OpenCL device (profiles/r1_opencl_reference.json:
    max work-group size 1024, max allocation 47,875,719,168 bytes): tile = threads = tiles per call = 1024,
    maxThresholds = 47875719168 div (1024^2 * 16) = 2853.  The canonical spec of the benchmarks is smaller; this
    is the one an unmodified Gudni would run with here."""
    assert 47875719168 // (1024 ** 2 * 16) == DEVICE_SPEC.max_thresholds
    assert_identical(*both(scenes.fuzzy_circles(2000, 1500, 1100, 5, 50, 3), DEVICE_SPEC))
    assert_identical(*both(scenes.mixed_bag(200, 1300, 900, 11), DEVICE_SPEC))
    assert_identical(*both(scenes.picture_scene(1100, 1030), DEVICE_SPEC))
