"""GPU parity, level 1: the oracle's tiles through the CUDA kernels via the C ABI."""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec

from parity import level1_parity, level2_parity

pytestmark = pytest.mark.gpu

SMALL = [scenes.tiny_square, scenes.medium_square, scenes.full_rectangle, scenes.stack_of_squares,
         scenes.open_square, scenes.concentric_squares2, scenes.concentric_squares3,
         scenes.six_point_rectangle, scenes.hour_glass, scenes.translucent_stack]


@pytest.mark.parametrize("make", SMALL, ids=lambda f: f.__name__)
def test_catalogue_scenes(rasterizer, make):
    level1_parity(rasterizer, make())


@pytest.mark.parametrize("theta", [0.3, 0.4, 0.5, 0.625])
@pytest.mark.parametrize("size", [100, 512])
def test_s1_square(rasterizer, size, theta):
    level1_parity(rasterizer, scenes.square(size, theta))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_shapes(rasterizer, seed):
    level1_parity(rasterizer, scenes.random_rectangles(300, 640, 480, seed))


def test_fuzzy_circles_small(rasterizer):
    level1_parity(rasterizer, scenes.fuzzy_circles(3000, 800, 600, 5, 50, 0xC0FFEE))


def test_dense_circles_split_to_small_tiles(rasterizer):
    # enough overlap that tiles split down towards 8x8 and some exceed 126 shapes
    img, stats, ref = level1_parity(rasterizer, scenes.fuzzy_circles(4000, 256, 256, 5, 50, 0xD15EA5E))
    assert min(int(j.tiles["right"].min() - j.tiles["left"].min()) for j in ref.jobs) >= 0
    assert stats.n_tiles > 64


def test_empty_scene_is_background(rasterizer):
    s = scenes.fuzzy_circles(0, 300, 200, 5, 50, 1, background=(0.25, 0.5, 0.75, 1.0))
    img, stats, ref = level1_parity(rasterizer, s)
    assert np.all(img == img[0, 0])


def test_shared_reciprocal_division_is_ieee(rasterizer):
    """composite's three divisions share one reciprocal (raster_device.cuh div3); the result must be
    the correctly rounded quotient, bit for bit, on every operand triple."""
    assert rasterizer.debug_selftest(n=1 << 30, seed=0x5EED) == 0
    assert rasterizer.debug_selftest(n=1 << 28, seed=12345) == 0


def test_picture_substances(rasterizer):
    """Texture substances (testPict): per-pixel picture lookups, subtract inside a picture substance."""
    level1_parity(rasterizer, scenes.picture_scene(640, 480, flowers_size=(700, 375)))
    level1_parity(rasterizer, scenes.picture_scene(900, 700, scale=2.0, flowers_size=(700, 375)))


def test_s2_paragraph_reduced(rasterizer):
    level1_parity(rasterizer, scenes.s2(960, 540, lines=14))


def test_level1_jobs_are_collected_across_launches(rasterizer):
    # 10 jobs / 2,325 tiles: the queued jobs are launched once 2,048 tiles are waiting and the rest at
    # frame_end; thread counts, taps and pixels must not depend on where the launches fall
    scene = scenes.fuzzy_circles(30000, 2048, 1024, 5, 50, 0x7117)
    img, stats, ref = level1_parity(rasterizer, scene)
    assert len(ref.jobs) >= 2 and stats.n_tiles > 2048
    level2_parity(rasterizer, scene, ref=ref)
