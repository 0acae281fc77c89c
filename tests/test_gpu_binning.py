"""GPU parity, level 2: tile binning on the GPU must reproduce the reference's tile tree
(Raster/TileTree.hs) bit for bit — boxes, depths, traversal order, per-tile shape order."""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec

from parity import level2_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("make", [scenes.tiny_square, scenes.full_rectangle, scenes.open_square,
                                  scenes.concentric_squares3, scenes.translucent_stack],
                         ids=lambda f: f.__name__)
def test_catalogue_scenes(rasterizer, make):
    level2_parity(rasterizer, make())


@pytest.mark.parametrize("size,theta", [(100, 0.4), (512, 0.3), (512, 0.625)])
def test_s1_square(rasterizer, size, theta):
    level2_parity(rasterizer, scenes.square(size, theta))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_shapes(rasterizer, seed):
    level2_parity(rasterizer, scenes.random_rectangles(400, 700, 500, seed))


def test_fuzzy_circles_multi_root(rasterizer):
    # 1000x700 -> 4x4 root tiles of 256, several of them outside the canvas
    level2_parity(rasterizer, scenes.fuzzy_circles(6000, 1000, 700, 5, 50, 0xBEEF))


def test_dense_splits_to_minimum_tiles(rasterizer):
    # heavy overlap: leaves go down to 8x8 and keep more than 126 shapes there
    img, stats, ref = level2_parity(rasterizer, scenes.fuzzy_circles(5000, 200, 200, 5, 50, 0xFACE))
    sizes = np.concatenate([j.tiles["right"] - j.tiles["left"] for j in ref.jobs])
    counts = np.concatenate([j.tiles["shape_count"] for j in ref.jobs])
    assert sizes.min() == 8 and counts.max() > 126


def test_strand_limit_forces_split(rasterizer):
    # a tiny strand budget makes the strand sum, not the shape count, drive the splitting
    spec = RasterSpec(max_strands_per_tile=40)
    from gudni_b200.raster import setup_rasterizer
    r = setup_rasterizer(spec=spec)
    try:
        level2_parity(r, scenes.fuzzy_circles(300, 512, 512, 5, 50, 0xABCD), spec=spec)
    finally:
        r.close()


def test_empty_scene(rasterizer):
    s = scenes.fuzzy_circles(0, 300, 200, 5, 50, 1, background=(0.25, 0.5, 0.75, 1.0))
    level2_parity(rasterizer, s)


def test_small_tile_spec(rasterizer):
    spec = RasterSpec(64, 64, 64, 256, 254, 127)
    from gudni_b200.raster import setup_rasterizer
    r = setup_rasterizer(spec=spec)
    try:
        level2_parity(r, scenes.fuzzy_circles(800, 300, 260, 5, 40, 0x1234), spec=spec)
    finally:
        r.close()


def test_picture_scene(rasterizer):
    level2_parity(rasterizer, scenes.picture_scene(640, 480, flowers_size=(700, 375)))


def test_s2_paragraph(rasterizer):
    level2_parity(rasterizer, scenes.s2(1920, 1080))


def test_s3_plots_textures_reduced(rasterizer):
    level2_parity(rasterizer, scenes.s3(1920, 1080))
