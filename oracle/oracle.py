"""ctypes driver of the CPU oracle (libgudni_oracle.so).  TEST INFRASTRUCTURE, NOT PRODUCT.

Pinned to the reference's own kernel file compiled for the host (oracle/_ref/libgudni_ref.so, built
by oracle/refbuild/build_ref.py; `reference=True` below drives it): see kernels_oracle.hpp and
tests/test_reference_pin.py.  The tile tree (Haskell) stays a restatement.

`render(scene, spec)` plays the role of drawFrame's hot path (Application.hs:239-241):
buildRasterJobs -> queueRasterJobs -> per job generate / sort / render.
"""
import ctypes
import os
import subprocess

import numpy as np

from gudni_b200.formats import SHAPE_DTYPE, TILE_DTYPE, CSpec, RasterSpec, CANONICAL_SPEC

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgudni_oracle.so")
_lib = None


_REF_PATH = os.path.join(_HERE, "_ref", "libgudni_ref.so")
_ref_lib = None


def build_reference(force=False):
    """oracle/_ref/libgudni_ref.so — the reference's own Kernels.cl compiled for the host
    (oracle/refbuild/build_ref.py).  Returns None when neither /root/reference nor a prebuilt
    library is present."""
    from oracle.refbuild import build_ref
    return build_ref.build(force=force)


def reference_lib():
    """The compiled reference kernels, or None if they cannot be had (no reference tree and no
    prebuilt library)."""
    global _ref_lib
    if _ref_lib is None:
        path = build_reference()
        if path is None:
            return None
        c = ctypes
        L = ctypes.CDLL(path)
        L.gudni_ref_threads.restype = c.c_int
        L.gudni_ref_set_threads.argtypes = [c.c_int]
        L.gudni_ref_raster_job.restype = c.c_int64
        L.gudni_ref_raster_job.argtypes = [c.c_void_p] * 5 + [c.c_int, c.c_int, c.POINTER(CSpec), c.c_void_p,
                                                              c.c_void_p, c.c_int, c.c_int, c.c_void_p,
                                                              c.c_void_p, c.c_void_p, c.POINTER(c.c_int64)]
        _ref_lib = L
    return _ref_lib


def build(force=False):
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        c = ctypes
        L = ctypes.CDLL(_LIB_PATH)
        L.gudni_oracle_threads.restype = c.c_int
        L.gudni_oracle_set_threads.argtypes = [c.c_int]
        L.gudni_oracle_build_jobs.restype = c.c_void_p
        L.gudni_oracle_build_jobs.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.POINTER(CSpec)]
        L.gudni_oracle_jobs_count.argtypes = [c.c_void_p]
        L.gudni_oracle_job_info.argtypes = [c.c_void_p, c.c_int] + [c.POINTER(c.c_int)] * 3
        L.gudni_oracle_job_shapes.restype = c.c_void_p
        L.gudni_oracle_job_shapes.argtypes = [c.c_void_p, c.c_int]
        L.gudni_oracle_job_tiles.restype = c.c_void_p
        L.gudni_oracle_job_tiles.argtypes = [c.c_void_p, c.c_int]
        L.gudni_oracle_jobs_free.argtypes = [c.c_void_p]
        L.gudni_oracle_raster_job.restype = c.c_int64
        L.gudni_oracle_raster_job.argtypes = [c.c_void_p] * 5 + [c.c_int, c.c_int, c.POINTER(CSpec), c.c_void_p,
                                                                 c.c_void_p, c.c_int, c.c_int, c.c_void_p,
                                                                 c.c_void_p, c.c_void_p, c.POINTER(c.c_int64)]
        _lib = L
    return _lib


class Job:
    """RasterJob (Raster/Job.hs:68-73)."""

    def __init__(self, shapes, tiles, columns):
        self.shapes, self.tiles, self.columns = shapes, tiles, int(columns)


def build_raster_jobs(scene, spec: RasterSpec = CANONICAL_SPEC):
    """buildRasterJobs (OpenCL/CallKernels.hs:244-255) over addShapeToTree (Raster/TileTree.hs:113).
    Jobs come back in the order the reference submits them: last-created first."""
    L = lib()
    cs = spec.to_c()
    entries = np.ascontiguousarray(scene.entries)
    h = L.gudni_oracle_build_jobs(entries.ctypes.data, len(entries), scene.width, scene.height, ctypes.byref(cs))
    jobs = []
    try:
        for j in range(L.gudni_oracle_jobs_count(h)):
            ns, nt, col = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            L.gudni_oracle_job_info(h, j, ctypes.byref(ns), ctypes.byref(nt), ctypes.byref(col))
            sp = L.gudni_oracle_job_shapes(h, j)
            tp = L.gudni_oracle_job_tiles(h, j)
            shapes = (np.frombuffer(ctypes.string_at(sp, ns.value * 16), dtype=SHAPE_DTYPE).copy()
                      if ns.value else np.zeros(0, SHAPE_DTYPE))
            tiles = np.frombuffer(ctypes.string_at(tp, nt.value * 32), dtype=TILE_DTYPE).copy()
            jobs.append(Job(shapes, tiles, col.value))
    finally:
        L.gudni_oracle_jobs_free(h)
    return jobs


def tiles_in_tree_order(jobs):
    """Concatenate the jobs' tiles/shapes in tile-tree traversal order (= creation order =
    reverse of submission order) with shape_start rebased to the concatenated shape array and
    column_allocation rebased to one frame-wide thread numbering."""
    tiles, shapes = [], []
    shape_base = col_base = 0
    for job in reversed(jobs):
        t = job.tiles.copy()
        t["shape_start"] += shape_base
        t["column_allocation"] += col_base
        tiles.append(t)
        shapes.append(job.shapes)
        shape_base += len(job.shapes)
        col_base += job.columns
    return np.concatenate(tiles), (np.concatenate(shapes) if shapes else np.zeros(0, SHAPE_DTYPE))


class RenderResult:
    def __init__(self, image, n_thresholds, shape_bits, total_thresholds, overflow_threads, jobs):
        self.image = image                    # (H, W) uint32 BGRA words
        self.n_thresholds = n_thresholds      # per job: int32[columns], -1 = inactive thread
        self.shape_bits = shape_bits
        self.total_thresholds = total_thresholds
        self.overflow_threads = overflow_threads
        self.jobs = jobs


def raster_jobs(scene, jobs, spec: RasterSpec = CANONICAL_SPEC, taps=True, threads=None, reference=False):
    """queueRasterJobs (OpenCL/CallKernels.hs:218-242): every job through the three kernels —
    the restated ones, or with `reference=True` the reference's own (oracle/_ref)."""
    if reference:
        L = reference_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/libgudni_ref.so is not built and /root/reference is absent")
        raster_job = L.gudni_ref_raster_job
        if threads:
            L.gudni_ref_set_threads(int(threads))
    else:
        L = lib()
        raster_job = L.gudni_oracle_raster_job
        if threads:
            L.gudni_oracle_set_threads(int(threads))
    cs = spec.to_c()
    out = np.zeros((scene.height, scene.width), dtype=np.uint32)
    geometry = np.ascontiguousarray(scene.geometry)
    substances = np.ascontiguousarray(scene.substances, dtype=np.float32)
    pict = np.ascontiguousarray(scene.picture_bytes)
    uses = np.ascontiguousarray(scene.picture_uses)
    bg = np.ascontiguousarray(scene.background, dtype=np.float32)
    counts, bits = [], []
    total = 0
    overflow = 0
    for job in jobs:
        nt = np.full(job.columns, -1, np.int32) if taps else None
        sb = np.full(job.columns, -1, np.int32) if taps else None
        tt = ctypes.c_int64(0)
        shapes = np.ascontiguousarray(job.shapes)
        tiles = np.ascontiguousarray(job.tiles)
        overflow += raster_job(
            geometry.ctypes.data, substances.ctypes.data, pict.ctypes.data, uses.ctypes.data, bg.ctypes.data,
            scene.width, scene.height, ctypes.byref(cs), shapes.ctypes.data, tiles.ctypes.data, len(tiles),
            job.columns, out.ctypes.data, nt.ctypes.data if taps else None, sb.ctypes.data if taps else None,
            ctypes.byref(tt))
        total += tt.value
        counts.append(nt)
        bits.append(sb)
    return RenderResult(out, counts, bits, total, overflow, jobs)


def render(scene, spec: RasterSpec = CANONICAL_SPEC, taps=True, threads=None, reference=False):
    return raster_jobs(scene, build_raster_jobs(scene, spec), spec, taps, threads, reference)


def host_threads():
    return lib().gudni_oracle_threads()


def bgra_to_rgb(image):
    """(H,W) uint32 BGRA words -> (H,W,3) uint8 RGB."""
    return np.stack([(image >> 16) & 0xFF, (image >> 8) & 0xFF, image & 0xFF], axis=-1).astype(np.uint8)
