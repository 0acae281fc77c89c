"""Level 3 on the GPU: gudni_b200_raster_outlines takes the scene BEFORE serialisation (outlines +
transformer stacks), builds the geometry heap and the shape entries on the device, then bins and rasterizes.
Held to (1) the harness's restatement of Raster/Strand.hs & co. — geometry heap and entries byte for byte —
and (2) the oracle's image of the same scene, bit-exact."""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import ENTRY_DTYPE
from oracle import oracle

pytestmark = pytest.mark.gpu


def check(rasterizer, scene, image=True):
    img, stats = rasterizer.raster_outlines(0, scene)
    geometry, entries = rasterizer.debug_strands()
    assert len(entries) == scene.n_shapes
    for field in ENTRY_DTYPE.names:
        assert np.array_equal(entries[field], scene.entries[field]), field
    assert geometry.nbytes == scene.geometry.nbytes
    bad = np.flatnonzero(geometry != scene.geometry)
    assert len(bad) == 0, f"geometry heaps differ from byte {bad[:4]}"
    if image:
        ref = oracle.render(scene, taps=False)
        assert np.array_equal(img, ref.image), "level-3 image differs from the oracle"
        assert stats.n_thresholds == ref.total_thresholds
    return stats


CATALOGUE = [scenes.tiny_square, scenes.open_square, scenes.hour_glass, scenes.translucent_stack, scenes.full_rectangle]


@pytest.mark.parametrize("make", CATALOGUE, ids=lambda f: f.__name__)
def test_catalogue_scenes(rasterizer, make):
    check(rasterizer, make())


@pytest.mark.parametrize("case", range(8))
def test_mixed_bag(rasterizer, case):
    rng = np.random.default_rng(5000 + case)
    w, h = int(rng.integers(20, 700)), int(rng.integers(20, 500))
    check(rasterizer, scenes.mixed_bag(int(rng.integers(1, 300)), w, h, 7000 + case))


def test_pictures_and_glyphs(rasterizer):
    check(rasterizer, scenes.picture_scene())
    check(rasterizer, scenes.s2(640, 300, lines=6))


def test_empty_scene(rasterizer):
    from gudni_b200.scene import SceneBuilder
    scene = SceneBuilder(64, 48, (0.25, 0.5, 0.75, 1.0)).freeze()
    img, stats = rasterizer.raster_outlines(0, scene)
    assert stats.n_thresholds == 0 and np.all(img == img[0, 0])


def test_everything_culled(rasterizer):
    from gudni_b200.scene import SceneBuilder
    b = SceneBuilder(64, 48, (0.25, 0.5, 0.75, 1.0))
    b.circle(b.solid(1, 0, 0, 1), [("translate", -500.0, -500.0), ("scale", 10.0)])
    scene = b.freeze()
    assert scene.n_shapes == 0 and scene.culled == 1
    check(rasterizer, scene)


def test_s4_full_size(rasterizer):
    """100,000 placements of one outline: 32 MB of geometry built on the device, identical to the harness's;
    the image identical to the frame rasterized from the harness's geometry (level 2)."""
    scene = scenes.s4()
    stats = check(rasterizer, scene, image=False)
    img3, _ = rasterizer.raster_outlines(1, scene)
    img2, _ = rasterizer.raster_scene(2, scene)
    assert np.array_equal(img3, img2)
    assert stats.ms_strands > 0.0


def test_bad_indices_are_refused(rasterizer):
    from gudni_b200.raster import GudniError
    scene = scenes.fuzzy_circles(10, 64, 64, 2, 10, 1)
    shapes, outlines, pairs, transforms = (a.copy() for a in scene.raw)
    shapes["first_outline"][3] = 7
    scene.raw = (shapes, outlines, pairs, transforms)
    with pytest.raises(GudniError):
        rasterizer.raster_outlines(0, scene)
    img, _ = rasterizer.raster_scene(1, scenes.tiny_square())     # the context is still usable
    assert img.shape == (16, 16)


def test_hand_worked_outlines(rasterizer):
    """The GPU strand kernels against the fixtures derived by hand from Raster/Strand.hs, Deknob.hs and ReorderTable.hs
    (tests/golden/strands_handworked.py): knob split, last run first, reversed runs, a run of 17 cut at 16, tree order."""
    from golden.strands_handworked import CASES
    from test_strands_handworked import check as check_fixture, scene_of
    for name in sorted(CASES):
        scene = scene_of(CASES[name])
        rasterizer.raster_outlines(0, scene)
        geometry, entries = rasterizer.debug_strands()
        check_fixture(CASES[name], geometry, entries)
