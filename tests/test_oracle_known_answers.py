"""Pins for the CPU oracle.  The reference has no tests or golden vectors of its own for this path
(SURVEY.md §4); the oracle is pinned (1) by scenes whose answer follows by hand from the reference's
definitions and (2) by golden vectors produced by the reference's own kernel source compiled for the
host (tests/golden/, tests/test_reference_pin.py).  Hand-derived: exact area coverage of axis-aligned rectangles, `composite` (Kernels.cl:878-887),
truncating BGRA conversion (:842-844), add/subtract semantics of determineColor (:1447-1513)."""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec
from oracle import oracle

F = np.float32


def to_byte(v):
    return int(np.trunc(F(v) * F(255.0)))


def rgb(img, x, y):
    w = int(img[y, x])
    return ((w >> 16) & 0xFF, (w >> 8) & 0xFF, w & 0xFF, (w >> 24) & 0xFF)


def mix(cov, fg, bg):
    """Expected bytes for a pixel covered `cov` by opaque fg over opaque bg, the way the reference
    accumulates: sum(colour*area)/sum(area) in f32 then truncation."""
    out = []
    for f, b in zip(fg, bg):
        acc = F(f) * F(cov) + F(b) * (F(1.0) - F(cov))
        out.append(int(np.trunc(acc * F(255.0))))
    return tuple(out)


def test_tiny_square_coverage():
    r = oracle.render(scenes.tiny_square())
    img = r.image
    assert r.overflow_threads == 0
    red, blue = (1, 0, 0), (0, 0, 1)
    # covered fractions per pixel: x,y in {0: 0.9, 1: 1.0, 2: 0.1}
    frac = {0: 0.9, 1: 1.0, 2: 0.1, 3: 0.0}
    for y in range(4):
        for x in range(4):
            cov = F(frac[x]) * F(frac[y])
            got = rgb(img, x, y)
            want = mix(cov, red, blue)
            assert got[3] == 255
            assert all(abs(g - w) <= 1 for g, w in zip(got[:3], want)), (x, y, got, want)
    assert rgb(img, 1, 1)[:3] == (255, 0, 0)
    assert rgb(img, 5, 5)[:3] == (0, 0, 255)
    # exact spot values: 0.81 -> 206, 0.9 -> 229, 0.09 -> 22, 0.01 -> 2
    assert rgb(img, 0, 0)[0] == 206 and rgb(img, 1, 0)[0] == 229
    assert rgb(img, 2, 0)[0] == 22 and rgb(img, 2, 2)[0] == 2


def test_medium_square_interior_and_edges():
    img = oracle.render(scenes.medium_square()).image
    for y in range(1, 10):
        for x in range(1, 10):
            assert rgb(img, x, y)[:3] == (255, 0, 0)
    assert rgb(img, 0, 5)[0] == 229 and rgb(img, 10, 5)[0] in (25, 26)
    assert rgb(img, 12, 12)[:3] == (0, 0, 255)


def test_full_rectangle_covers_canvas():
    s = scenes.full_rectangle()
    img = oracle.render(s).image
    assert np.all(img == np.uint32(0xFFFF0000))


def test_stack_of_squares_abutting():
    img = oracle.render(scenes.stack_of_squares()).image
    for x in range(4):
        for y in range(4):
            assert rgb(img, x, y)[:3] == (255, 0, 0)
            assert rgb(img, x, y + 4)[:3] == (0, 255, 0)
    assert rgb(img, 4, 0)[:3] == (0, 0, 255) and rgb(img, 0, 8)[:3] == (0, 0, 255)


def test_open_square_subtraction_and_alpha():
    img = oracle.render(scenes.open_square(alpha=0.5)).image
    # ring: orange (1,.5,0) a=.5 over opaque blue background -> composite formula
    fg = np.array([1.0, 0.5, 0.0, 0.5], F)
    bg = np.array([0.0, 0.0, 1.0, 1.0], F)
    a = fg[3] + bg[3] * (F(1) - fg[3])
    c = (fg[:3] * fg[3] + bg[:3] * bg[3] * (F(1) - fg[3])) / a
    want = tuple(int(np.trunc(v * F(255))) for v in c)
    for (x, y) in [(0, 0), (4, 4), (0, 2), (2, 0), (4, 1)]:
        assert rgb(img, x, y)[:3] == want, (x, y)
    # hole shows the background: the subtract shape on top suppresses its own substance
    for (x, y) in [(1, 1), (2, 2), (3, 3)]:
        assert rgb(img, x, y)[:3] == (0, 0, 255), (x, y)
    assert rgb(img, 5, 5)[:3] == (0, 0, 255)


def test_concentric_squares_abut_without_seams():
    img = oracle.render(scenes.concentric_squares3()).image
    red, green, blue, black = (255, 0, 0), (0, 255, 0), (0, 0, 255), (0, 0, 0)
    for y in range(10):
        for x in range(10):
            ring = min(x, y, 9 - x, 9 - y)
            want = red if ring < 2 else green if ring < 4 else blue
            assert rgb(img, x, y)[:3] == want, (x, y)
    assert rgb(img, 10, 10)[:3] == black
    img2 = oracle.render(scenes.concentric_squares2()).image
    for y in range(5):
        for x in range(5):
            ring = min(x, y, 4 - x, 4 - y)
            want = red if ring < 1 else blue if ring < 2 else black
            assert rgb(img2, x, y)[:3] == want, (x, y)


def test_six_point_rectangle_and_hourglass():
    img = oracle.render(scenes.six_point_rectangle()).image
    assert rgb(img, 0, 0)[:3] == (255, 0, 0) and rgb(img, 1, 0)[:3] == (255, 0, 0)
    assert rgb(img, 2, 0)[:3] == (0, 0, 255) and rgb(img, 0, 1)[:3] == (0, 0, 255)
    img = oracle.render(scenes.hour_glass()).image
    # the bow-tie (0,0)-(8,8)-(8,0)-(0,8): left and right triangles filled, top/bottom empty
    assert rgb(img, 0, 4)[:3] == (255, 0, 0) and rgb(img, 7, 4)[:3] == (255, 0, 0)
    assert rgb(img, 4, 0)[:3] == (0, 0, 255) and rgb(img, 4, 7)[:3] == (0, 0, 255)


def test_translucent_stack_matches_composite_chain():
    s = scenes.translucent_stack(size=32, layers=5)
    img = oracle.render(s).image
    colors = [scenes.RED, scenes.GREEN, scenes.BLUE, scenes.YELLOW, scenes.ORANGE]

    def over(fg, bg):
        a = F(fg[3]) + F(bg[3]) * (F(1) - F(fg[3]))
        c = [(F(fg[i]) * F(fg[3]) + F(bg[i]) * F(bg[3]) * (F(1) - F(fg[3]))) / a for i in range(3)]
        return (c[0], c[1], c[2], a)

    # centre pixel is under all five layers; layer 0 is top-most
    for depth, (x, y) in enumerate([(0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (5, 5)]):
        base = (F(0), F(0), F(0), F(0))
        present = [i for i in range(5) if i + 1 <= x]   # layer i covers [i+1, size-i-1)
        for i in present:                               # top-most first = lowest index first
            base = over(base, colors[i] + (0.5,))
        base = over(base, (1.0, 1.0, 1.0, 1.0))
        want = tuple(int(np.trunc(F(v) * F(255))) for v in base[:3])
        assert rgb(img, x, y)[:3] == want, (x, y, depth)


def test_area_sums_to_one_everywhere():
    """Structural invariant: opaque full-canvas shape under anything gives alpha-consistent pixels;
    here: every pixel of a random scene is written exactly once with alpha 255."""
    s = scenes.random_rectangles(40, 96, 80, seed=7)
    r = oracle.render(s)
    assert r.overflow_threads == 0
    assert np.all((r.image >> 24) == 0xFF)


@pytest.mark.parametrize("spec", [RasterSpec(), RasterSpec(64, 64, 64, 256, 254, 127),
                                  RasterSpec(32, 32, 32, 512, 510, 127)])
def test_image_independent_of_raster_spec(spec):
    """Tile size / threads per tile change the slab decomposition, not the picture, for geometry on
    coordinates where the slab-relative arithmetic is exact."""
    s = scenes.concentric_squares3(size=40)
    base = oracle.render(s).image
    assert np.array_equal(oracle.render(s, spec).image, base)


def test_picture_substance_texel_lookup():
    """readColor for a picture (Kernels.cl:1420-1441): texel = trunc(absPos / scale - translate),
    colour = RGBA8 / 255, outside the picture transparent; pixel = trunc(colour * 255)."""
    from gudni_b200.scene import SceneBuilder
    rng = np.random.default_rng(3)
    pict = rng.integers(0, 256, size=(6, 8, 4), dtype=np.uint8)
    pict[..., 3] = 255
    b = SceneBuilder(32, 32, (0.0, 0.0, 0.0, 1.0))
    p = b.picture(pict)
    s = b.picture_substance(p, translate=(2.0, 3.0), scale=2.0)
    b.shape(s, [scenes._straight_outline([(0, 0), (30, 0), (30, 30), (0, 30)])], is_picture=True)
    img = oracle.render(b.freeze()).image
    for (x, y) in [(4, 6), (5, 7), (10, 9), (19, 17), (3, 6), (4, 5), (20, 6), (4, 18)]:
        tx = int(np.trunc(F(x) / F(2.0) - F(2.0)))
        ty = int(np.trunc(F(y) / F(2.0) - F(3.0)))
        if 0 <= tx < 8 and 0 <= ty < 6:
            want = tuple(int(np.trunc((F(pict[ty, tx, c]) / F(255.0)) * F(255.0))) for c in range(3))
        else:
            want = (0, 0, 0)   # transparent picture over the opaque black background
        assert rgb(img, x, y)[:3] == want, (x, y, tx, ty)


def test_oracle_matches_reference_golden_vectors():
    """The restated oracle against vectors the REFERENCE'S OWN kernels produced
    (tests/golden/reference_hashes.json, written by tests/golden/make_golden.py from
    oracle/_ref/libgudni_ref.so = Kernels.cl compiled for the host): image, threshold total,
    per-thread threshold counts and shape-bit counts, all bit-exact."""
    import json
    import os
    from golden.make_golden import SCENES, digest, render
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_hashes.json")
    golden = json.load(open(path))
    assert set(golden) == set(SCENES)
    for name in SCENES:
        assert digest(render(name, reference=False)) == golden[name], name


def test_reference_golden_images_verbatim():
    """The smallest golden scenes are stored as images too; the oracle reproduces them word for word."""
    import os
    from golden.make_golden import VERBATIM, render
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_images.npz"))
    for name in VERBATIM:
        assert np.array_equal(render(name, reference=False).image, stored[name]), name


def _opencl_hashes():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "opencl_b200_hashes.json")))


def test_oracle_matches_vectors_from_a_real_opencl_run():
    """tests/golden/opencl_b200_hashes.json was written on the GPU box by oracle/refbuild/ocl_run.py: the
    reference's Kernels.cl, verbatim, compiled and run by NVIDIA's OpenCL 3.0 runtime on the B200
    ("strict": FP_CONTRACT OFF + correctly rounded divide, i.e. IEEE like the oracle).  Image, threshold
    total, per-thread counts and shape bits of the oracle hash to the same values — small scenes here,
    S2/S4b/S4 at full size below."""
    from golden.make_golden import SCENES, digest, render
    strict = _opencl_hashes()["strict"]
    for name in SCENES:
        assert digest(render(name, reference=False)) == strict[name], name
    extra = {"fuzzy_circles_2000": scenes.fuzzy_circles(2000, 640, 480, 5, 50, 0x5EED),
             "random_rectangles_300": scenes.random_rectangles(300, 640, 480, 5)}
    for name, scene in extra.items():
        assert digest(oracle.render(scene)) == strict[name], name


@pytest.mark.parametrize("name", ["s4b", "s4"])
def test_oracle_matches_real_opencl_run_at_full_size(name):
    """BASELINE.json's 3840x2160 configurations: S4b (100k curves) and S4 (100k circles, 14.1 M thresholds)."""
    from golden.make_golden import digest
    strict = _opencl_hashes()["strict"]
    assert digest(oracle.render(getattr(scenes, name)())) == strict[name]


def test_reference_build_options_stay_within_the_north_star_tolerance():
    """Under the reference's own options (-cl-fast-relaxed-math) the same OpenCL run differs from the IEEE
    one by at most 1/255 on any channel; recorded by ocl_run.py in profiles/r1_opencl_reference.json."""
    import json
    import os
    prof = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "r1_opencl_reference.json")))
    for name, r in prof["results"]["reference"].items():
        assert r["max_channel_diff"] <= 1 and r["shape_bits_equal"], name
        if name in ("s2", "s3", "s4b", "s4"):
            assert r["exact_rate"] >= 0.999, name
    for name, r in prof["results"]["strict"].items():
        assert r["digest_equals_oracle"] and r["pixels_differing"] == 0, name
