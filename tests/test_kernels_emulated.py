"""The CUDA kernels themselves in the CPU suite.

tests/native/raster_emu.cpp compiles gudni_b200/csrc/raster_kernels.cu (+ raster_device.cuh, raster_warp.cuh —
the text nvcc compiles, unchanged) with g++ against a host stand-in for the CUDA device language
(tests/native/emu/cuda_runtime.h) and runs it under a cooperative SIMT emulator: one fiber per CUDA thread, warp
shuffles / ballots / __syncthreads as rendezvous points, one CTA at a time, deadlock detection for collectives
that cannot complete.  These tests hold the emulated kernels — generate, sweep with its colour cache and
pending list, the HBM-queue replay of spilled threads, tile ordering, strand bounds — to the oracle: image,
per-thread threshold counts and shape-bit counts, bit-exact.  It is not the product path (that needs a B200 and
has no fallback) and not a performance model; it is how a kernel change is checked when no GPU is at hand."""
import ctypes

import numpy as np
import pytest

from gudni_b200 import _build, scenes
from gudni_b200.scene import SceneBuilder
from gudni_b200.formats import CANONICAL_SPEC, CSpec, RasterSpec, SHAPE_DTYPE
from oracle import oracle


@pytest.fixture(scope="module")
def emu():
    L = ctypes.CDLL(_build.build_raster_emu())
    c = ctypes
    L.raster_emu_frame.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                   c.POINTER(CSpec), c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p, c.c_int, c.c_int64,
                                   c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
    return L


def emulate_frame(L, scene, jobs, spec=CANONICAL_SPEC):
    """Level 1 of the ABI the way the shim lays a frame out: all jobs end to end."""
    tiles, shapes, base = [], [], []
    shape_base = column_base = 0
    for job in jobs:
        t = job.tiles.copy()
        t["shape_start"] += shape_base
        base.append((column_base + job.tiles["column_allocation"]).astype(np.int32))
        tiles.append(t)
        shapes.append(job.shapes)
        shape_base += len(job.shapes)
        column_base += job.columns
    tiles = np.ascontiguousarray(np.concatenate(tiles))
    shapes = np.ascontiguousarray(np.concatenate(shapes)) if shape_base else np.zeros(1, SHAPE_DTYPE)
    base = np.ascontiguousarray(np.concatenate(base))
    out = np.zeros((scene.height, scene.width), np.uint32)
    counts = np.zeros(column_base, np.int32)
    bits = np.zeros(column_base, np.int32)
    stats = np.zeros(8, np.int64)
    g = np.ascontiguousarray(scene.geometry)
    s = np.ascontiguousarray(scene.substances, np.float32)
    p = np.ascontiguousarray(scene.picture_bytes)
    u = np.ascontiguousarray(scene.picture_uses)
    bg = np.ascontiguousarray(scene.background, np.float32)
    ptr = lambda a: a.ctypes.data if a.size else None  # noqa: E731
    cs = spec.to_c()
    rc = L.raster_emu_frame(ptr(g), g.nbytes, ptr(s), ptr(p), ptr(u), bg.ctypes.data, scene.width, scene.height,
                            ctypes.byref(cs), shapes.ctypes.data, shape_base, tiles.ctypes.data, base.ctypes.data,
                            len(tiles), column_base, out.ctypes.data, counts.ctypes.data, bits.ctypes.data, stats.ctypes.data)
    return rc, out, counts, bits, stats


def run(L, scene, spec=CANONICAL_SPEC):
    ref = oracle.render(scene, spec, taps=True)
    assert ref.overflow_threads == 0
    rc, out, counts, bits, stats = emulate_frame(L, scene, ref.jobs, spec)
    assert rc == 0
    assert np.array_equal(counts, np.concatenate(ref.n_thresholds)), "per-thread threshold counts"
    assert np.array_equal(bits, np.concatenate(ref.shape_bits)), "per-thread shape bits"
    assert stats[0] == ref.total_thresholds and stats[2] == 0
    bad = np.argwhere(out != ref.image)
    assert len(bad) == 0, f"{len(bad)} pixels differ, first at (y,x)={bad[:5].tolist()}"
    return stats


CATALOGUE = [scenes.tiny_square, scenes.medium_square, scenes.full_rectangle, scenes.stack_of_squares,
             scenes.open_square, scenes.concentric_squares2, scenes.concentric_squares3,
             scenes.six_point_rectangle, scenes.hour_glass, scenes.translucent_stack]


@pytest.mark.parametrize("make", CATALOGUE, ids=lambda f: f.__name__)
def test_catalogue_scenes(emu, make):
    run(emu, make())


@pytest.mark.parametrize("value", [np.inf, -np.inf])
@pytest.mark.parametrize("coord", [0, 1], ids=["x", "y"])
def test_infinite_coordinate_is_refused_not_rasterized(emu, value, coord):
    """A point at +-infinity makes the bisection of Kernels.cl:1226-1258 loop for ever (the reference's own kernels
    hang on it).  strand_bounds_kernel flags it, tile_order_kernel empties the shape lists of the launch so that the
    raster kernels never walk a strand, and gudni_b200_frame_end turns the flag into GUDNI_ERR_ARGUMENT.  The jobs come from the finite scene: the oracle
    would not return on the poisoned one."""
    scene = scenes.medium_square()
    jobs = oracle.build_raster_jobs(scene)
    scene.geometry = scene.geometry.copy()
    points = scene.geometry.view(np.float32)
    points[2 * 2 + coord] = value            # record 0 is the strand header; unit 2 is an on-curve point
    rc, out, counts, bits, stats = emulate_frame(emu, scene, jobs)
    assert rc == 0
    assert stats[4] == 1 and stats[0] == 0 and stats[1] == 0
    assert (counts[counts >= 0] == 0).all()            # background only: no thread generated a threshold


@pytest.mark.parametrize("what", ["geo_start", "size_word", "odd_size"])
def test_records_that_leave_the_geometry_heap_are_refused(emu, what):
    """geo_start, the strand count and the strands' size words come from the caller; strand_bounds_kernel walks every
    strand header before the raster kernels do, flags a record that leaves the heap (or a size that is not whole
    16-byte records) and the frame is defused like one with an infinite coordinate — no out-of-bounds read."""
    scene = scenes.medium_square()
    jobs = oracle.build_raster_jobs(scene)
    if what == "geo_start":
        for job in jobs:
            job.shapes = job.shapes.copy()
            job.shapes["geo_start"] = len(scene.geometry) // 16 + 1000
    else:
        scene.geometry = scene.geometry.copy()
        words = scene.geometry.view(np.uint16)
        words[0] = 0x7FF0 if what == "size_word" else words[0] + 1
    rc, out, counts, bits, stats = emulate_frame(emu, scene, jobs)
    assert rc == 0
    assert (stats[4] & 2) and stats[0] == 0
    assert (counts[counts >= 0] == 0).all()


def test_nan_and_huge_coordinates_still_rasterize(emu):
    """NaN and finite values up to 3e38 terminate in the reference, so they are not refused (bits equal the oracle's)."""
    for value in (np.nan, 3.0e38, -3.0e38):
        scene = scenes.medium_square()
        scene.geometry = scene.geometry.copy()
        scene.geometry.view(np.float32)[2 * 2] = value
        stats = run(emu, scene)
        assert stats[4] == 0


def test_circles_rectangles_pictures(emu):
    run(emu, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
    run(emu, scenes.random_rectangles(120, 200, 160, 7))
    run(emu, scenes.picture_scene(320, 300, flowers_size=(350, 200)))
    run(emu, scenes.mixed_bag(120, 300, 200, 7003))


def test_small_tile_spec(emu):
    spec = RasterSpec(64, 64, 64, 256, 254, 127)
    run(emu, scenes.fuzzy_circles(200, 150, 130, 5, 40, 0x1234), spec)


def test_queue_leaves_the_shared_memory_window(emu):
    stats = run(emu, scenes.thin_rectangles(60, width=64, spacing=3.0, thickness=1.3))
    assert stats[1] == 0                      # > 64 thresholds per column, still on chip


def identical_shapes(n, width=48, height=40):
    """n copies of one rotated rectangle, each its own shape: every column's queue holds runs of thresholds whose
    sort keys are equal in all three components, so the order among them is the order they were built in."""
    from gudni_b200.scene import SceneBuilder
    b = SceneBuilder(width, height, (1.0, 1.0, 1.0, 1.0), name=f"identical-{n}")
    for i in range(n):
        b.rectangle(b.solid(0.1 + 0.8 * (i % 7) / 7.0, 0.5, 0.9 - 0.8 * (i % 5) / 5.0, 0.35), 20.3, 25.7,
                    [("translate", 9.2, 6.1), ("rotate", 0.03)])
    return b.freeze()


@pytest.mark.parametrize("n", [12, 50, 100])
def test_equal_keys_keep_the_order_they_were_built_in(emu, n):
    """The rank sort (queues up to 64) and the bitonic network (longer ones) against the reference's bubble sort on
    queues full of ties: 2 n thresholds per column, n of them equal to each other at the top edge."""
    run(emu, identical_shapes(n))


def test_runs_past_the_slice_kernels_scratch_take_the_wide_pass(emu):
    """A run of more than 12 thresholds that start together does not fit the slice kernel's shared-memory scratch: the
    thread is flagged, raster_slice_wide_kernel slices it again with room for 72, and only a longer run still goes to
    the lane-private replay."""
    assert run(emu, identical_shapes(12))[1] == 0
    assert run(emu, identical_shapes(13))[1] == 0
    assert run(emu, identical_shapes(50))[1] == 0
    assert run(emu, identical_shapes(72))[1] == 0
    assert run(emu, identical_shapes(73))[1] > 0
    emu.raster_emu_set_batches(3)
    try:
        assert run(emu, identical_shapes(40, width=300, height=40))[1] == 0
    finally:
        emu.raster_emu_set_batches(1)


def test_long_queues_take_the_bitonic_network(emu):
    stats = run(emu, scenes.thin_rectangles(100, width=64, spacing=2.0, thickness=0.9))
    assert stats[1] == 0                      # 128 < thresholds per column <= 256, still on chip


def test_replay_of_spilled_threads(emu):
    """Queues past the on-chip capacity of 256, and tiles with more shapes than stack bits at the 8-pixel
    floor: both go through the lane-private HBM-queue replay kernel."""
    stats = run(emu, scenes.thin_rectangles(150, width=256, height=256, spacing=1.5, thickness=0.7, one_shape=True))
    assert stats[1] > 0
    stats = run(emu, scenes.fuzzy_circles(1200, 96, 96, 5, 50, 77))
    assert stats[1] > 0


# ---- levels 2 and 3: the binning and strand-building kernels under the emulator as well ----------------------
from gudni_b200.formats import ENTRY_DTYPE, TILE_DTYPE  # noqa: E402


@pytest.fixture(scope="module")
def emu_scene(emu):
    c = ctypes
    vp, i32, i64, sz = c.c_void_p, c.c_int, c.c_int64, c.c_size_t
    emu.raster_emu_scene.argtypes = [vp, sz, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, c.POINTER(CSpec), vp,
                                     vp, sz, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp]
    return emu


def run_scene(L, scene, level, spec=CANONICAL_SPEC):
    ref = oracle.render(scene, spec, taps=True)
    ref_tiles, ref_shapes = oracle.tiles_in_tree_order(ref.jobs)
    ptr = lambda a: a.ctypes.data if a is not None and a.size else None  # noqa: E731
    g = np.ascontiguousarray(scene.geometry)
    e = np.ascontiguousarray(scene.entries)
    s = np.ascontiguousarray(scene.substances, np.float32)
    p = np.ascontiguousarray(scene.picture_bytes)
    u = np.ascontiguousarray(scene.picture_uses)
    bg = np.ascontiguousarray(scene.background, np.float32)
    raw = [np.ascontiguousarray(a) for a in scene.raw] if level == 3 else [None] * 4
    out = np.zeros((scene.height, scene.width), np.uint32)
    geometry_out = np.zeros(g.nbytes + 64, np.uint8)
    entries_out = np.zeros(len(e) + 4, ENTRY_DTYPE)
    tiles_out = np.zeros(len(ref_tiles) + 16, TILE_DTYPE)
    shapes_out = np.zeros(len(ref_shapes) + 16, SHAPE_DTYPE)
    columns = sum(j.columns for j in ref.jobs)
    counts = np.zeros(columns, np.int32)
    bits = np.zeros(columns, np.int32)
    sizes = np.zeros(5, np.int64)
    stats = np.zeros(8, np.int64)
    cs = spec.to_c()
    rc = L.raster_emu_scene(ptr(g) if level == 2 else None, g.nbytes if level == 2 else 0, ptr(e) if level == 2 else None,
                            len(e) if level == 2 else 0, ptr(raw[0]), len(raw[0]) if level == 3 else 0, ptr(raw[1]), ptr(raw[2]),
                            ptr(raw[3]), ptr(s), ptr(p), ptr(u), bg.ctypes.data, scene.width, scene.height, ctypes.byref(cs),
                            out.ctypes.data, geometry_out.ctypes.data, geometry_out.nbytes, entries_out.ctypes.data,
                            len(entries_out), tiles_out.ctypes.data, len(tiles_out), shapes_out.ctypes.data, len(shapes_out),
                            counts.ctypes.data, bits.ctypes.data, columns, sizes.ctypes.data, stats.ctypes.data)
    assert rc == 0
    if level == 3:      # strands: heap and entries as the harness serialises them
        assert sizes[0] == len(e) and sizes[1] == g.nbytes
        assert entries_out[:len(e)].tobytes() == e.tobytes()
        assert geometry_out[:g.nbytes].tobytes() == g.tobytes()
    # binning: tiles in tree order, per-tile shape lists, column numbering
    assert sizes[2] == len(ref_tiles) and sizes[3] == len(ref_shapes) and sizes[4] == columns
    tiles = tiles_out[:len(ref_tiles)]
    for field in ("left", "top", "right", "bottom", "h_depth", "v_depth", "shape_start", "shape_count"):
        assert np.array_equal(tiles[field], ref_tiles[field]), field
    assert shapes_out[:len(ref_shapes)].tobytes() == ref_shapes.tobytes()
    # raster
    assert np.array_equal(counts, np.concatenate(list(reversed(ref.n_thresholds))))
    assert np.array_equal(bits, np.concatenate(list(reversed(ref.shape_bits))))
    assert stats[0] == ref.total_thresholds
    assert np.array_equal(out, ref.image)


@pytest.mark.parametrize("level", [2, 3])
def test_binning_and_strand_kernels(emu_scene, level):
    run_scene(emu_scene, scenes.tiny_square(), level)
    run_scene(emu_scene, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6), level)
    run_scene(emu_scene, scenes.mixed_bag(100, 300, 200, 7003), level)
    run_scene(emu_scene, scenes.fuzzy_circles(400, 150, 130, 5, 40, 0x1234), level, RasterSpec(64, 64, 64, 256, 254, 127))


def test_binning_splits_down_to_the_floor(emu_scene):
    run_scene(emu_scene, scenes.fuzzy_circles(1200, 96, 96, 5, 50, 77), 2)


def test_threshold_store_exhausted(emu):
    """A hand-over store that is too small (the shim sizes it from the previous frame's demand, so a first
    frame can meet this): the generate kernel gives whole warps to the replay kernel — slower, same pixels."""
    emu.raster_emu_set_store_entries.argtypes = [ctypes.c_size_t]
    emu.raster_emu_set_store_entries(600)
    try:
        stats = run(emu, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
        assert stats[1] > 0
    finally:
        emu.raster_emu_set_store_entries(0)


def test_degenerate_outlines(emu_scene):
    """Outlines of fewer than two curve pairs produce no strands (Strand.hs:175-177) but still count towards
    the shape's bounding box (onShape boxes the outlines first); a shape made only of such outlines is an entry
    without geometry."""
    b = SceneBuilder(64, 48, (0.1, 0.2, 0.3, 1.0))
    red = b.solid(1, 0, 0, 0.5)
    square = np.array([[10, 10, 20, 10], [30, 10, 30, 20], [30, 30, 20, 30], [10, 30, 10, 20]], np.float32)
    b.shape(red, [square, np.array([[50, 40, 51, 41]], np.float32)])        # a real outline + a single pair
    b.shape(b.solid(0, 1, 0, 1), [np.array([[5, 5, 6, 6]], np.float32)])    # only a single pair
    scene = b.freeze()
    assert scene.n_shapes == 2 and scene.entries["num_strands"][1] == 0
    assert scene.entries["right"][0] == np.float32(51.0)
    run_scene(emu_scene, scene, 3)
    run_scene(emu_scene, scene, 2)


@pytest.mark.parametrize("level", [2, 3])
def test_boxes_beyond_int32_are_binned(emu_scene, level):
    """Boxes that, divided by the root tile size, do not fit an int32 (binning.cu forEachRoot clamps in float)."""
    run_scene(emu_scene, scenes.huge_boxes(), level)
    for seed in range(3):         # sizes log-uniform up to 1e30 pixels
        run_scene(emu_scene, scenes.far_shapes(12, 150, 110, 0xFA50 + seed), level)


@pytest.mark.parametrize("field", ["left", "top", "right", "bottom"])
def test_boxes_that_break_the_contract_are_binned_as_the_tile_tree_would(emu_scene, field):
    """Level 2 asks for culled, consistent boxes.  Given others — a side turned into NaN, +-inf or +-1e30, so that the
    box is inverted or off the canvas — addShapeToTree (TileTree.hs:120-139) still does something definite: it tests
    the box against cuts only, so the first / last row or column of root tiles takes what lies beyond it.  The binning
    kernels do the same (tiles, shape lists, counts and pixels equal the oracle's)."""
    for value in (np.nan, np.inf, -np.inf, 1e30, -1e30):
        scene = scenes.fuzzy_circles(30, 300, 280, 5, 40, 11)      # 2 x 2 root tiles of 256 pixels
        scene.entries = scene.entries.copy()
        scene.entries[field][3] = value
        scene.entries[field][17] = value
        run_scene(emu_scene, scene, 2)


@pytest.mark.parametrize("level", [2, 3])
def test_infinite_coordinate_is_refused_at_levels_2_and_3(emu_scene, level):
    """Level 2: the poisoned heap arrives with finite boxes.  Level 3: the outline itself holds the infinite point; the
    strand kernels (transform, box, knob splitting, reordering) get through it, the shapes whose box still meets the
    canvas reach the heap, and the flag goes up there.  Either way no thread generates a threshold."""
    L = emu_scene
    scene = scenes.fuzzy_circles(20, 120, 100, 5, 30, 3)
    ptr = lambda a: a.ctypes.data if a.size else None  # noqa: E731
    s = np.ascontiguousarray(scene.substances, np.float32)
    bg = np.ascontiguousarray(scene.background, np.float32)
    cs = CANONICAL_SPEC.to_c()
    for value in (np.inf, -np.inf):
        for coord in (0, 1):
            canvas = np.zeros((scene.height, scene.width), np.uint32)
            sizes, stats = np.zeros(5, np.int64), np.zeros(8, np.int64)
            if level == 2:
                g = np.ascontiguousarray(scene.geometry).copy()
                g.view(np.float32)[2 * 2 + coord] = value
                e = np.ascontiguousarray(scene.entries)
                rc = L.raster_emu_scene(ptr(g), g.nbytes, ptr(e), len(e), None, 0, None, None, None, ptr(s), None, None,
                                        bg.ctypes.data, scene.width, scene.height, ctypes.byref(cs), canvas.ctypes.data,
                                        None, 0, None, 0, None, 0, None, 0, None, None, 0, sizes.ctypes.data, stats.ctypes.data)
            else:
                raw = [np.ascontiguousarray(a).copy() for a in scene.raw]
                raw[2].view(np.float32).reshape(-1)[coord] = value
                rc = L.raster_emu_scene(None, 0, None, 0, ptr(raw[0]), len(raw[0]), ptr(raw[1]), ptr(raw[2]), ptr(raw[3]), ptr(s),
                                        None, None, bg.ctypes.data, scene.width, scene.height, ctypes.byref(cs),
                                        canvas.ctypes.data, None, 0, None, 0, None, 0, None, 0, None, None, 0,
                                        sizes.ctypes.data, stats.ctypes.data)
            assert rc == 0
            assert stats[4] == 1 and stats[0] == 0 and stats[1] == 0, (level, value, coord, stats)


@pytest.mark.parametrize("level", [2, 3])
def test_strips_reassemble_the_frame(emu_scene, level):
    """gudni_b200_frame_strip on the kernels' side: the binning kernels skip root tiles outside the strip, the
    raster kernels' threads outside it are inactive.  Three strips of whole root-tile rows (64-pixel tiles)
    rendered separately into one canvas give the frame, all entries passed to every strip."""
    L = emu_scene
    spec = RasterSpec(64, 64, 64, 256, 254, 127)
    scene = scenes.fuzzy_circles(300, 200, 230, 5, 40, 0x57A1)
    ref = oracle.render(scene, spec, taps=False)
    ptr = lambda a: a.ctypes.data if a.size else None  # noqa: E731
    g = np.ascontiguousarray(scene.geometry)
    e = np.ascontiguousarray(scene.entries)
    s = np.ascontiguousarray(scene.substances, np.float32)
    bg = np.ascontiguousarray(scene.background, np.float32)
    raw = [np.ascontiguousarray(a) for a in scene.raw]
    canvas = np.full((scene.height, scene.width), 0xDEADBEEF, np.uint32)
    sizes, stats = np.zeros(5, np.int64), np.zeros(8, np.int64)
    cs = spec.to_c()
    total = 0
    L.raster_emu_set_strip.argtypes = [ctypes.c_int, ctypes.c_int]
    try:
        for rows in ((0, 64), (64, 192), (192, 230)):
            L.raster_emu_set_strip(*rows)
            if level == 2:
                rc = L.raster_emu_scene(ptr(g), g.nbytes, ptr(e), len(e), None, 0, None, None, None, ptr(s), None, None,
                                        bg.ctypes.data, scene.width, scene.height, ctypes.byref(cs), canvas.ctypes.data,
                                        None, 0, None, 0, None, 0, None, 0, None, None, 0, sizes.ctypes.data, stats.ctypes.data)
            else:     # every strip builds the strands of the whole scene, then bins its own rows
                rc = L.raster_emu_scene(None, 0, None, 0, ptr(raw[0]), len(raw[0]), ptr(raw[1]), ptr(raw[2]), ptr(raw[3]), ptr(s),
                                        None, None, bg.ctypes.data, scene.width, scene.height, ctypes.byref(cs),
                                        canvas.ctypes.data, None, 0, None, 0, None, 0, None, 0, None, None, 0,
                                        sizes.ctypes.data, stats.ctypes.data)
            assert rc == 0
            total += int(stats[0])
            # rows of the other strips are untouched so far or already final
            assert np.array_equal(canvas[rows[0]:rows[1]], ref.image[rows[0]:rows[1]])
    finally:
        L.raster_emu_set_strip(0, 0)
    assert np.array_equal(canvas, ref.image)
    assert total == ref.total_thresholds


def test_shared_reciprocal_division(emu):
    """div3 (raster_device.cuh) — one reciprocal, a Newton step and FMA corrections for the three quotients of
    `composite` — runs here from a model of rcp.approx.  It returns the correctly rounded quotient whichever
    way the approximation errs (0, +-1, +-2, +-8 ulps), on the operand generator of the device's own self-test;
    and a frame rendered with a reciprocal that is off by an ulp either way is the same frame."""
    emu.raster_emu_selftest_div3.restype = ctypes.c_uint64
    emu.raster_emu_selftest_div3.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    emu.raster_emu_set_rcp_error.argtypes = [ctypes.c_int]
    try:
        for ulps in (0, 1, -1, 2, -2, 8, -8):
            emu.raster_emu_set_rcp_error(ulps)
            assert emu.raster_emu_selftest_div3(400_000, 1 + abs(ulps)) == 0, ulps
        for ulps in (1, -1):
            emu.raster_emu_set_rcp_error(ulps)
            run(emu, scenes.fuzzy_circles(120, 160, 120, 4, 40, 9))
    finally:
        emu.raster_emu_set_rcp_error(0)


def test_kernel_variant_small_tables():
    """A compile-time variant checked the same way (a 64-line stack cache in the resolve kernel, small chunks of
    strands and (strand, column) pairs in the generate kernel: every eviction and chunk-boundary path runs on small
    scenes) — what tools/variants.sh builds for the GPU can be held to the oracle here first."""
    L = ctypes.CDLL(_build.build_raster_emu(extra=("-DGUDNI_COLOR_LINES=64", "-DGUDNI_STRAND_TABLE=8", "-DGUDNI_GEN_ITEMS=256"),
                                            suffix="_small"))
    c = ctypes
    L.raster_emu_frame.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                   c.POINTER(CSpec), c.c_void_p, c.c_int64, c.c_void_p, c.c_void_p, c.c_int, c.c_int64,
                                   c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
    run(L, scenes.translucent_stack())
    run(L, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
    run(L, scenes.picture_scene(320, 300, flowers_size=(350, 200)))


@pytest.mark.parametrize("shift", [1, 2])
def test_units_narrower_than_a_warp(emu, shift):
    """Launches of few tiles are dealt out in units of 16 or 8 column-threads (forEachUnit): same bits."""
    emu.raster_emu_set_lane_shift(shift)
    try:
        run(emu, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
        run(emu, scenes.picture_scene(320, 300, flowers_size=(350, 200)))
        run(emu, scenes.thin_rectangles(60, width=64, spacing=3.0, thickness=1.3))
    finally:
        emu.raster_emu_set_lane_shift(0)


@pytest.mark.parametrize("batches", [2, 3, 8])
def test_batches_of_a_launch(emu, batches):
    """A launch's tiles dealt into interleaved batches with their own work cursors and their own region of the
    stack table (rasterTiles): same bits; more batches than tiles included."""
    emu.raster_emu_set_batches(batches)
    try:
        run(emu, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
        run(emu, scenes.picture_scene(320, 300, flowers_size=(350, 200)))
        run(emu, scenes.fuzzy_circles(400, 600, 300, 4, 40, 7), RasterSpec(64, 64, 256, 1024, 1022, 127))
    finally:
        emu.raster_emu_set_batches(1)


def test_device_derived_spec(emu, emu_scene):
    """The RasterSpec the reference would derive from the B200's OpenCL device (1,024-pixel tiles, 1,024 threads
    per tile, MAXTHRESHOLDS 2,853 — tests/test_reference_pin.py): 32 warps per tile, 1,024-pixel root tiles."""
    spec = RasterSpec(1024, 1024, 1024, 2853, 2851, 127)
    run(emu, scenes.fuzzy_circles(150, 300, 200, 4, 40, 6), spec)
    run_scene(emu_scene, scenes.mixed_bag(80, 260, 180, 7003), 3, spec)
    run_scene(emu_scene, scenes.fuzzy_circles(60, 1100, 600, 10, 80, 8), 2, spec)


def test_threshold_overflow_is_counted_like_the_oracle(emu):
    """MAXTHRESHOLDS = 16 and a stack of thin slanted rectangles over the left half of the canvas: those columns
    overflow (undefined behaviour in the reference, App. B #9).  The kernels report exactly the threads the
    oracle flags and leave every other column's pixels exact."""
    spec = RasterSpec(max_thresholds=16)
    b = SceneBuilder(32, 96, (1.0, 1.0, 1.0, 1.0))
    for i in range(40):
        b.shape(b.solid(0.1, 0.2, 0.8, 0.5), [scenes._straight_outline([(0.0, 2.0 * i + 1.0), (15.5, 2.0 * i + 1.6),
                                                                        (15.5, 2.0 * i + 2.4), (0.0, 2.0 * i + 1.8)])])
    b.rectangle(b.solid(0.9, 0.1, 0.1, 0.7), 10.0, 60.0, [("translate", 19.3, 11.7)])
    scene = b.freeze()
    ref = oracle.render(scene, spec, taps=True)
    assert 0 < ref.overflow_threads < 32
    job = ref.jobs[0]
    assert len(ref.jobs) == 1
    out = np.zeros((scene.height, scene.width), np.uint32)
    stats = np.zeros(8, np.int64)
    g = np.ascontiguousarray(scene.geometry)
    s = np.ascontiguousarray(scene.substances, np.float32)
    bg = np.ascontiguousarray(scene.background, np.float32)
    base = np.ascontiguousarray(job.tiles["column_allocation"].astype(np.int32))
    tiles, shapes = np.ascontiguousarray(job.tiles), np.ascontiguousarray(job.shapes)
    cs = spec.to_c()
    emu.raster_emu_frame(g.ctypes.data, g.nbytes, s.ctypes.data, None, None, bg.ctypes.data, scene.width, scene.height,
                         ctypes.byref(cs), shapes.ctypes.data, len(shapes), tiles.ctypes.data, base.ctypes.data, len(tiles),
                         job.columns, out.ctypes.data, None, None, stats.ctypes.data)
    assert stats[2] == ref.overflow_threads
    written = ref.image != 0                       # the oracle leaves an overflowed thread's pixels untouched
    assert written.any() and np.array_equal(out[written], ref.image[written])


@pytest.mark.parametrize("mode", [1, 2], ids=["descending", "shuffled"])
def test_thread_order_between_rendezvous_points_does_not_matter(emu, emu_scene, mode):
    """A poor man's race check.  The emulator runs each thread until it reaches a shuffle, ballot, __syncwarp or
    __syncthreads; in which order the threads get there is arbitrary on hardware.  With the lanes run in
    descending order, or in a fresh pseudo-random order on every scheduling pass, every kernel still produces
    the same bits — shared-memory hand-offs (colour cache lines, pending list, queue windows, CTA scans) are all
    behind the synchronisation they need."""
    emu.raster_emu_set_schedule.argtypes = [ctypes.c_int]
    emu.raster_emu_set_schedule(mode)
    try:
        run(emu, scenes.translucent_stack())
        run(emu, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
        run(emu, scenes.thin_rectangles(150, width=256, height=256, spacing=1.5, thickness=0.7, one_shape=True))
        run_scene(emu_scene, scenes.mixed_bag(100, 300, 200, 7003), 3)
        run_scene(emu_scene, scenes.fuzzy_circles(400, 150, 130, 5, 40, 0x1234), 2, RasterSpec(64, 64, 64, 256, 254, 127))
    finally:
        emu.raster_emu_set_schedule(0)
