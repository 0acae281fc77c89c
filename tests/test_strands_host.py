"""Level 3 (outlines -> geometry heap + shape entries), CPU half: the per-shape logic the CUDA kernels run
(gudni_b200/csrc/strand_build.cuh compiled for the host by tests/native/strand_check.cpp) against the
harness's restatement of the reference's Haskell (Raster/Strand.hs, Deknob.hs, ReorderTable.hs, the geometry
half of Serialize.hs:onShape — gudni_b200/csrc/host/strand.hpp, scene.cpp), byte for byte: geometry heap,
shape entries (tag, geoStart, strand count, bounding box), culling.  Parity unpinned for this row: the
Haskell itself cannot run here; the known answers below are the ones SURVEY.md App. A derives by hand."""
import ctypes

import numpy as np
import pytest

from gudni_b200 import _build, scenes
from gudni_b200.formats import ENTRY_DTYPE


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(_build.build_strand_check())
    c = ctypes
    L.strand_check_build.restype = c.c_int64
    L.strand_check_build.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                     c.c_void_p, c.c_size_t, c.POINTER(c.c_size_t), c.c_void_p, c.POINTER(c.c_int64)]
    L.strand_check_table_row.argtypes = [c.c_int, c.c_void_p]
    return L


def build(lib, scene):
    shapes, outlines, pairs, transforms = (np.ascontiguousarray(a) for a in scene.raw)
    nbytes, nstr = ctypes.c_size_t(), ctypes.c_int64()
    ptr = lambda a: a.ctypes.data if a.size else None  # noqa: E731
    kept = lib.strand_check_build(ptr(shapes), len(shapes), ptr(outlines), ptr(pairs), ptr(transforms), scene.width,
                                  scene.height, None, 0, ctypes.byref(nbytes), None, ctypes.byref(nstr))
    geometry = np.zeros(nbytes.value, np.uint8)
    entries = np.zeros(kept, ENTRY_DTYPE)
    got = lib.strand_check_build(ptr(shapes), len(shapes), ptr(outlines), ptr(pairs), ptr(transforms), scene.width,
                                 scene.height, geometry.ctypes.data, geometry.nbytes, ctypes.byref(nbytes),
                                 entries.ctypes.data, ctypes.byref(nstr))
    assert got == kept
    return geometry, entries, nstr.value


def assert_same_as_harness(lib, scene):
    geometry, entries, n_strands = build(lib, scene)
    assert len(entries) == scene.n_shapes, (len(entries), scene.n_shapes)
    assert len(scene.raw[0]) - len(entries) == scene.culled
    for field in ENTRY_DTYPE.names:
        assert np.array_equal(entries[field], scene.entries[field]), field
    assert geometry.nbytes == scene.geometry.nbytes
    bad = np.flatnonzero(geometry != scene.geometry)
    assert len(bad) == 0, f"geometry heaps differ from byte {bad[:4]}"
    assert n_strands == int(scene.entries["num_strands"].sum())


def test_reorder_table_known_answers(lib):
    """n=4 -> [8,0,1,4,5,2,3,6,7]; n=8 -> [16,0,1,8,9,4,5,12,13,2,3,6,7,10,11,14,15]."""
    def row(n):
        out = np.zeros(2 * n + 1, np.uint8)
        lib.strand_check_table_row(n, out.ctypes.data)
        return out.tolist()
    assert row(1) == [2, 0, 1]
    assert row(4) == [8, 0, 1, 4, 5, 2, 3, 6, 7]
    assert row(8) == [16, 0, 1, 8, 9, 4, 5, 12, 13, 2, 3, 6, 7, 10, 11, 14, 15]
    for n in range(1, 17):
        assert sorted(row(n)) == list(range(2 * n + 1))     # a permutation of the strand's points


CATALOGUE = [scenes.tiny_square, scenes.medium_square, scenes.full_rectangle, scenes.stack_of_squares,
             scenes.open_square, scenes.concentric_squares2, scenes.concentric_squares3,
             scenes.six_point_rectangle, scenes.hour_glass, scenes.translucent_stack]


@pytest.mark.parametrize("make", CATALOGUE, ids=lambda f: f.__name__)
def test_catalogue_scenes(lib, make):
    assert_same_as_harness(lib, make())


@pytest.mark.parametrize("theta", [0.3, 0.4, 0.5, 0.625])
def test_rotated_square(lib, theta):
    assert_same_as_harness(lib, scenes.square(512, theta))


def test_circles_share_one_outline(lib):
    scene = scenes.fuzzy_circles(3000, 640, 480, 5, 50, 0x5EED)
    assert len(scene.raw[1]) == 1 and len(scene.raw[2]) == 17      # one unit circle: 16 arcs + the closing sliver
    assert_same_as_harness(lib, scene)


@pytest.mark.parametrize("case", range(12))
def test_mixed_bag(lib, case):
    """Knobs to split, strands cut at 16 Béziers, right-to-left strands, last-run-first ordering,
    several outlines per shape, rotations, shapes culled off every side of the canvas."""
    rng = np.random.default_rng(5000 + case)
    w, h = int(rng.integers(20, 700)), int(rng.integers(20, 500))
    assert_same_as_harness(lib, scenes.mixed_bag(int(rng.integers(1, 300)), w, h, 7000 + case))


def test_pictures_glyphs_and_rectangles(lib):
    assert_same_as_harness(lib, scenes.picture_scene())
    assert_same_as_harness(lib, scenes.s2(640, 300, lines=6))
    assert_same_as_harness(lib, scenes.random_rectangles(300, 640, 480, 5))


def test_s4_full_size(lib):
    """S4: 100,000 placements of one 16-pair outline -> 32 MB of geometry, 1.6 M Béziers."""
    assert_same_as_harness(lib, scenes.s4())


def parse_heap(geometry, entries):
    """Walk the geometry heap the way buildThresholdArray does (Kernels.cl:1557-1581): per shape
    `num_strands` strands from 16*geo_start, each a u16 size in 8-byte units followed by size-1 points."""
    out = []
    for e in entries:
        at = 16 * int(e["geo_start"])
        strands = []
        for _ in range(int(e["num_strands"])):
            size = int(np.frombuffer(geometry[at:at + 2].tobytes(), "<u2")[0])
            assert geometry[at + 2:at + 8].tobytes() == b"\0" * 6
            pts = np.frombuffer(geometry[at + 8:at + 8 * size].tobytes(), "<f4").reshape(-1, 2)
            strands.append(pts)
            at += 8 * size
        out.append(strands)
    return out


def test_rectangle_by_hand(lib):
    """An axis-aligned w x h rectangle at (x,y), derived by hand from Raster/Strand.hs: the four sides are
    four runs (vertical sides never join, Strand.hs:77-81); the run still open at the end — the left side —
    comes first (:98-102); the bottom side runs right to left and is reversed (:137-143); a one-Bézier strand
    is stored as [right end, left end, control] (ReorderTable.hs:96-104), controls at the midpoints
    (Figure/Outline.hs:98-104)."""
    from gudni_b200.scene import SceneBuilder
    x, y, w, h = 3.25, 2.5, 7.0, 4.0
    b = SceneBuilder(32, 32)
    b.rectangle(b.solid(1, 0, 0, 1), w, h, [("translate", x, y)])
    scene = b.freeze()
    geometry, entries, n_strands = build(lib, scene)
    assert n_strands == 4 and geometry.nbytes == 4 * 32
    strands = parse_heap(geometry, entries)[0]
    tl, tr, br, bl = (x, y), (x + w, y), (x + w, y + h), (x, y + h)
    mid = lambda p, q: (0.5 * p[0] + 0.5 * q[0], 0.5 * p[1] + 0.5 * q[1])  # noqa: E731
    expected = [
        [tl, bl, mid(bl, tl)],       # left side, walked bottom-left -> top-left: [end, start, control]
        [tr, tl, mid(tl, tr)],       # top side, left to right
        [br, tr, mid(tr, br)],       # right side, top to bottom
        [br, bl, mid(br, bl)],       # bottom side, walked right to left, stored left to right
    ]
    for got, want in zip(strands, expected):
        assert np.array_equal(got, np.asarray(want, np.float32)), (got, want)
    e = entries[0]
    assert (e["left"], e["top"], e["right"], e["bottom"]) == (np.float32(x), np.float32(y), np.float32(x + w), np.float32(y + h))


@pytest.mark.parametrize("case", range(6))
def test_structural_invariants(lib, case):
    """What Kernels.cl relies on: size field = 2n+2 with 1 <= n <= 16, every strand runs
    left to right and is x-monotone (so the tree search by x is valid), the tree order is a search tree: the
    node at heap slot i splits the x-range of its subtree; strands tile the shape's slice of the heap."""
    rng = np.random.default_rng(5000 + case)
    scene = scenes.mixed_bag(int(rng.integers(20, 200)), int(rng.integers(50, 500)), int(rng.integers(50, 400)), 9000 + case)
    geometry, entries, _ = build(lib, scene)
    curves = 0
    for strands in parse_heap(geometry, entries):
        for pts in strands:
            n = (len(pts) - 1) // 2
            assert len(pts) == 2 * n + 1 and 1 <= n <= 16
            curves += n
            right, left = pts[0], pts[1]
            assert left[0] <= right[0]
            tree = pts[3::2][:n - 1] if n > 1 else np.zeros((0, 2), np.float32)    # on-curve points, heap order

            def check(i, lo, hi):
                if i >= len(tree):
                    return
                assert lo <= tree[i][0] <= hi, (i, lo, tree[i][0], hi)
                check(2 * i + 1, lo, tree[i][0])
                check(2 * i + 2, tree[i][0], hi)
            check(0, left[0], right[0])
    assert curves == scene.curves


def test_degenerate_outlines(lib):
    """Fewer than two curve pairs: no strands (Strand.hs:175-177), but the points still count towards the box."""
    from gudni_b200.scene import SceneBuilder
    b = SceneBuilder(64, 48)
    square = np.array([[10, 10, 20, 10], [30, 10, 30, 20], [30, 30, 20, 30], [10, 30, 10, 20]], np.float32)
    b.shape(b.solid(1, 0, 0, 0.5), [square, np.array([[50, 40, 51, 41]], np.float32)])
    b.shape(b.solid(0, 1, 0, 1), [np.array([[5, 5, 6, 6]], np.float32)])
    b.shape(b.solid(0, 0, 1, 1), [np.array([[500, 5, 600, 6]], np.float32)])      # off canvas: culled
    scene = b.freeze()
    assert scene.n_shapes == 2 and scene.culled == 1
    assert_same_as_harness(lib, scene)
