"""Capacity edges: queues longer than the on-chip capacity (HBM replay path), more shapes per tile
than stack bits, more thresholds than MAXTHRESHOLDS (undefined behaviour in the reference, reported
here), and argument validation through the C ABI."""
import ctypes

import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec
from gudni_b200.raster import GudniError, setup_rasterizer

from parity import level1_parity, level2_parity

pytestmark = pytest.mark.gpu


def test_long_queues_stay_on_chip(rasterizer):
    # ~120 thresholds per column-thread: far beyond the 8-entry shared-memory window, inside the 256-entry
    # on-chip capacity (the queue tail lives in the threshold store)
    scene = scenes.thin_rectangles(60, width=64, spacing=3.0, thickness=1.3)
    img, stats, ref = level1_parity(rasterizer, scene)
    assert max(int(c.max()) for c in ref.n_thresholds) > 64
    assert stats.n_spilled_threads == 0 and stats.n_overflow_threads == 0
    level2_parity(rasterizer, scene, ref=ref)


def test_queues_of_two_hundred_take_the_bitonic_network(rasterizer):
    scene = scenes.thin_rectangles(100, width=64, spacing=2.0, thickness=0.9)
    img, stats, ref = level1_parity(rasterizer, scene)
    assert max(int(c.max()) for c in ref.n_thresholds) > 128
    assert stats.n_spilled_threads == 0 and stats.n_overflow_threads == 0


@pytest.mark.parametrize("n", [12, 50, 100])
def test_equal_sort_keys_keep_build_order(rasterizer, n):
    """Queues full of thresholds that tie in all three sort keys (n copies of one shape): the rank sort and the
    bitonic network must leave them in the order the reference's bubble sort does."""
    from test_kernels_emulated import identical_shapes
    level1_parity(rasterizer, identical_shapes(n))


def test_runs_past_the_slice_scratch_take_the_wide_pass(rasterizer):
    """Runs of 13 to 72 thresholds that start together: flagged by the slice kernel, sliced again by
    raster_slice_wide_kernel, no thread handed to the lane-private replay; a run of 73 still is."""
    from test_kernels_emulated import identical_shapes
    for n in (13, 50, 72):
        img, stats, ref = level1_parity(rasterizer, identical_shapes(n))
        assert stats.n_spilled_threads == 0, n
    img, stats, ref = level1_parity(rasterizer, identical_shapes(73))
    assert stats.n_spilled_threads > 0
    # many units at once, mixed with ordinary ones
    img, stats, ref = level2_parity(rasterizer, identical_shapes(40, width=700, height=300))
    assert stats.n_spilled_threads == 0


def test_very_long_queues_take_the_replay_path(rasterizer):
    # one shape of 150 thin rectangles on a 256-wide tile: every column-thread is 256 rows tall and crosses
    # ~300 thresholds, over the 256-entry on-chip capacity, under MAXTHRESHOLDS
    scene = scenes.thin_rectangles(150, width=256, height=256, spacing=1.5, thickness=0.7, one_shape=True)
    img, stats, ref = level1_parity(rasterizer, scene)
    assert max(int(c.max()) for c in ref.n_thresholds) > 256
    assert stats.n_spilled_threads > 0 and stats.n_overflow_threads == 0
    level2_parity(rasterizer, scene, ref=ref)


def test_more_shapes_than_stack_bits(rasterizer):
    # 8-pixel tiles that keep > 127 shapes: only the first 127 that touch a column get a bit
    scene = scenes.fuzzy_circles(6000, 128, 128, 5, 40, 0xB175)
    img, stats, ref = level2_parity(rasterizer, scene)
    assert max(int(j.tiles["shape_count"].max()) for j in ref.jobs) > 127
    assert max(int(b.max()) for b in ref.shape_bits) == 127


def test_threshold_overflow_is_reported_not_corrupting():
    spec = RasterSpec(max_thresholds=16)
    r = setup_rasterizer(spec=spec)
    try:
        # ~100 thresholds per column-thread: the on-chip capacity is capped by max_thresholds = 16, so the
        # threads are replayed against HBM queues of 16 entries, which they overflow (and report)
        scene = scenes.thin_rectangles(100, width=32, spacing=3.0, thickness=1.3)
        img, stats = r.raster_scene(0, scene)
        assert stats.n_spilled_threads > 0
        assert stats.n_overflow_threads > 0
        # untouched threads still match the oracle run with the same spec
        from oracle import oracle
        ref = oracle.render(scene, spec, taps=False)
        assert ref.overflow_threads > 0
    finally:
        r.close()


def test_argument_validation(rasterizer):
    L, ctx = rasterizer._L, rasterizer._ctx
    bg = (ctypes.c_float * 4)(0, 0, 0, 1)
    assert L.gudni_b200_frame_begin(ctx, None, 16, None, 0, None, 0, None, 0, bg, 64, 64, 0) == -1   # null geometry
    assert L.gudni_b200_frame_begin(ctx, None, 0, None, 0, None, 0, None, 0, bg, 0, 64, 0) == -1     # empty bitmap
    assert L.gudni_b200_raster_scene(ctx, None, 0) == -4                                              # outside a frame
    assert b"outside a frame" in L.gudni_b200_last_error(ctx)
    scene = scenes.tiny_square()
    rasterizer.frame_begin(scene, 0)
    bad = np.zeros(1, dtype=[("left", "<i4"), ("top", "<i4"), ("right", "<i4"), ("bottom", "<i4"), ("h_depth", "<i2"),
                             ("v_depth", "<i2"), ("column_allocation", "<i4"), ("shape_start", "<u4"), ("shape_count", "<u4")])
    bad["shape_count"] = 5   # slice beyond the (empty) shape array
    assert L.gudni_b200_raster_job(ctx, None, 0, bad.ctypes.data, 1, 256, 0) == -1
    rasterizer.frame_end(want_image=False)
    with pytest.raises(GudniError):
        setup_rasterizer(spec=RasterSpec(threads_per_tile=100))   # not a power of two


def test_records_that_leave_the_geometry_heap_are_refused(rasterizer):
    """A shape record whose strands run past the geometry heap is refused with GUDNI_ERR_ARGUMENT (caught by
    strand_bounds_kernel before any raster kernel walks it); the context stays usable."""
    import numpy as np
    from gudni_b200 import scenes
    from gudni_b200.raster import GudniError
    from oracle import oracle
    scene = scenes.medium_square()
    good = oracle.render(scene, taps=False)
    bad = scenes.medium_square()
    bad.entries = bad.entries.copy()
    bad.entries["geo_start"] = len(bad.geometry) // 16 + 1000
    with pytest.raises(GudniError) as e:
        rasterizer.raster_scene(0, bad)
    assert e.value.code == -1 and "geometry heap" in str(e.value)
    bad2 = scenes.medium_square()
    bad2.geometry = bad2.geometry.copy()
    bad2.geometry.view(np.uint16)[0] = 0x7FF0        # a size word that runs far past the heap
    with pytest.raises(GudniError):
        rasterizer.raster_scene(0, bad2)
    img, stats = rasterizer.raster_scene(1, scene)
    assert np.array_equal(img, good.image)


def test_input_cache_skips_unchanged_uploads(rasterizer):
    """gudni_b200_frame_begin_cached: with unchanged generation counters the second frame uploads nothing and is the
    same image; a bumped counter picks up a buffer refilled in place; generation 0 always uploads."""
    import numpy as np
    from gudni_b200 import scenes
    scene = scenes.fuzzy_circles(20000, 1920, 1080, 5, 50, 0xCAC4E)
    gens = [7, 7, 7, 7, 7]
    first, st1 = rasterizer.raster_scene(0, scene, generations=gens)
    again, st2 = rasterizer.raster_scene(1, scene, generations=gens)
    assert np.array_equal(first, again)
    assert st2.ms_upload < 0.25 * st1.ms_upload + 0.02, (st1.ms_upload, st2.ms_upload)
    # refill the substances in place: without a new counter the library may keep what it has ...
    scene.substances = np.ascontiguousarray(scene.substances, np.float32).copy()
    scene.substances[:, :3] = scene.substances[:, :3][:, ::-1]
    stale, _ = rasterizer.raster_scene(2, scene, generations=gens)
    assert np.array_equal(stale, first)
    # ... with one it does not
    gens[1] = 8
    fresh, _ = rasterizer.raster_scene(3, scene, generations=gens)
    plain, _ = rasterizer.raster_scene(4, scene)
    assert np.array_equal(fresh, plain) and not np.array_equal(fresh, first)


def test_replay_of_more_threads_than_it_has_queue_slots(rasterizer):
    """More replayed column-threads than the replay kernel's 148 x 128 HBM queue slots: it goes round more than once, and
    the frame after gets a slot per thread (frame_end grows the queues): same pixels both times."""
    scene = scenes.fuzzy_circles(9000, 160, 160, 5, 40, 0xB177)
    img, stats, ref = level2_parity(rasterizer, scene)
    assert stats.n_spilled_threads > 148 * 128
    img2, stats2 = rasterizer.raster_scene(1, scene)
    assert stats2.n_spilled_threads == stats.n_spilled_threads
    assert np.array_equal(img2, ref.image)


def test_frame_stored_straight_into_a_host_bitmap(rasterizer):
    """gudni_b200_frame_target_host: the kernels' pixel stores go to a page-locked host bitmap; frame_end with the same
    pointer copies nothing; the pixels equal those of a frame rendered in HBM.  An unregistered bitmap is refused."""
    scene = scenes.fuzzy_circles(400, 700, 500, 5, 60, 0x405)
    ref, _ = rasterizer.raster_scene(0, scene)
    host = np.zeros((scene.height, scene.width), np.uint32)
    with pytest.raises(GudniError):
        rasterizer.frame_target_host(host)
    rasterizer.host_register(host)
    try:
        rasterizer.frame_target_host(host)
        for frame in range(2):
            host[:] = 0
            out, stats = rasterizer.raster_scene(frame, scene, out=host)
            assert np.array_equal(host, ref)
        # pictures and the lane-private replay store pixels too
        for other in (scenes.picture_scene(320, 300, flowers_size=(350, 200)),
                      scenes.thin_rectangles(150, width=256, height=256, spacing=1.5, thickness=0.7, one_shape=True)):
            rasterizer.frame_target_host(None)
            want, _ = rasterizer.raster_scene(0, other)
            small = np.zeros((other.height, other.width), np.uint32)
            rasterizer.host_register(small)
            try:
                rasterizer.frame_target_host(small)
                rasterizer.raster_scene(1, other, out=small)
                assert np.array_equal(small, want)
            finally:
                rasterizer.frame_target_host(None)
                rasterizer.host_unregister(small)
    finally:
        rasterizer.frame_target_host(None)
        rasterizer.host_unregister(host)
    again, _ = rasterizer.raster_scene(2, scene)
    assert np.array_equal(again, ref)


def test_pageable_transfers_through_the_staging_ring(rasterizer):
    """Transfers of a megabyte and more between pageable memory and the device go through the context's own page-locked
    ring, copied by several threads (hostcopy.cuh): sizes around the chunk and ring boundaries, uploads back to back
    (a staging buffer is reused only when the DMA that read it is done), page-locked memory beside them."""
    L, ctx = rasterizer._L, rasterizer._ctx
    rng = np.random.default_rng(0xC0B1)
    chunk = 4 << 20
    sizes = [1 << 20, (1 << 20) + 1, chunk - 1, chunk, chunk + 4097, 4 * chunk, 4 * chunk + 3, 9 * chunk + 12345, 1000]
    bufs = []
    for n in sizes:
        host = rng.integers(0, 256, n, dtype=np.uint8)
        p = ctypes.c_void_p()
        assert L.gudni_b200_device_alloc(ctx, n, ctypes.byref(p)) == 0
        bufs.append((host, p))
    for host, p in bufs:                  # all uploads first: the ring wraps around several times
        assert L.gudni_b200_upload(ctx, p, host.ctypes.data, host.nbytes) == 0
    for host, p in reversed(bufs):
        back = np.zeros_like(host)
        assert L.gudni_b200_download(ctx, back.ctypes.data, p, host.nbytes) == 0
        assert np.array_equal(back, host), host.nbytes
    host, p = bufs[-2]
    pinned = np.zeros_like(host)
    rasterizer.host_register(pinned)
    try:
        assert L.gudni_b200_download(ctx, pinned.ctypes.data, p, host.nbytes) == 0
        assert np.array_equal(pinned, host)
    finally:
        rasterizer.host_unregister(pinned)
    for host, p in bufs:
        assert L.gudni_b200_device_free(ctx, p) == 0
