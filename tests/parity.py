"""Shared helpers for the parity tests: run a scene through the oracle and through the C ABI and
compare at the three levels BASELINE.md §4 names — tile assignments (bit-exact), per-thread
threshold / shape-bit counts (bit-exact), BGRA8 pixels (<= 1/255 per channel, >= 99.9 % exact)."""
import numpy as np

from gudni_b200.formats import CANONICAL_SPEC
from oracle import oracle


def channel_diff(a, b):
    d = np.zeros(a.shape, dtype=np.int32)
    for shift in (0, 8, 16, 24):
        d = np.maximum(d, np.abs(((a >> shift) & 0xFF).astype(np.int32) - ((b >> shift) & 0xFF).astype(np.int32)))
    return d


def assert_images_match(gpu, ref, exact=False):
    assert gpu.shape == ref.shape
    if np.array_equal(gpu, ref):   # the usual case; spares the int32 temporaries on a 16384^2 canvas
        return
    d = channel_diff(gpu, ref)
    if exact:
        bad = np.argwhere(d > 0)
        assert len(bad) == 0, f"{len(bad)} pixels differ, first at (y,x)={bad[:5].tolist()}"
    else:
        assert d.max() <= 1, f"max channel error {d.max()} at {np.argwhere(d > 1)[:5].tolist()}"
        exact_rate = float((d == 0).mean())
        assert exact_rate >= 0.999, f"exact-pixel rate {exact_rate:.5f} < 0.999"


def level1_parity(rasterizer, scene, spec=CANONICAL_SPEC, exact=True):
    """Tiles binned by the oracle's tile tree, rasterized by the CUDA path (level 1 of the ABI)."""
    ref = oracle.render(scene, spec, taps=True)
    assert ref.overflow_threads == 0
    rasterizer.debug_enable(True)
    img, stats = rasterizer.queue_raster_jobs(0, scene, ref.jobs)
    n_thr, bits = rasterizer.debug_thread_counts()
    rasterizer.debug_enable(False)
    ref_thr = np.concatenate(ref.n_thresholds)
    ref_bits = np.concatenate(ref.shape_bits)
    assert n_thr.shape == ref_thr.shape
    assert np.array_equal(n_thr, ref_thr), f"threshold counts differ at threads {np.flatnonzero(n_thr != ref_thr)[:8]}"
    assert np.array_equal(bits, ref_bits), f"shape bits differ at threads {np.flatnonzero(bits != ref_bits)[:8]}"
    assert stats.n_thresholds == ref.total_thresholds
    assert stats.n_overflow_threads == 0
    assert_images_match(img, ref.image, exact=exact)
    return img, stats, ref


def level2_parity(rasterizer, scene, spec=CANONICAL_SPEC, exact=True, ref=None):
    """Tile binning AND rasterization on the GPU (level 2): tiles, per-tile shape lists, thread
    counts and pixels against the oracle's tile tree + kernels."""
    if ref is None:
        ref = oracle.render(scene, spec, taps=True)
    assert ref.overflow_threads == 0
    ref_tiles, ref_shapes = oracle.tiles_in_tree_order(ref.jobs)
    rasterizer.debug_enable(True)
    img, stats = rasterizer.raster_scene(0, scene)
    tiles, shapes = rasterizer.debug_binned()
    n_thr, bits = rasterizer.debug_thread_counts()
    rasterizer.debug_enable(False)
    # tile assignments: boxes, depths, order, column allocations, shape slices — bit-exact
    assert len(tiles) == len(ref_tiles), (len(tiles), len(ref_tiles))
    for field in ("left", "top", "right", "bottom", "h_depth", "v_depth", "shape_start", "shape_count"):
        assert np.array_equal(tiles[field], ref_tiles[field]), field
    # per-job column allocation as addTileToRasterJob assigns it
    per_job = np.concatenate([j.tiles["column_allocation"] for j in reversed(ref.jobs)])
    assert np.array_equal(tiles["column_allocation"], per_job)
    assert shapes.tobytes() == ref_shapes.tobytes(), "per-tile shape lists differ"
    # per-thread counts: reorder the oracle's per-job arrays into tree order
    ref_thr = np.concatenate(list(reversed(ref.n_thresholds)))
    ref_bits = np.concatenate(list(reversed(ref.shape_bits)))
    assert np.array_equal(n_thr, ref_thr)
    assert np.array_equal(bits, ref_bits)
    assert stats.n_tiles == len(ref_tiles) and stats.n_shape_refs == len(ref_shapes)
    assert stats.n_thresholds == ref.total_thresholds
    assert_images_match(img, ref.image, exact=exact)
    return img, stats, ref
