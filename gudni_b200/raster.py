"""Host-side mirror of the reference's Rasterizer interface, over the C ABI of libgudni_b200.so.

The names follow the reference so the call sites read like `drawFrame` (Application.hs:215-247):

    rasterizer = setup_rasterizer()                       # setupOpenCL        OpenCL/Setup.hs:102
    rasterizer.queue_raster_jobs(frame, scene, jobs)      # queueRasterJobs    OpenCL/CallKernels.hs:218
    rasterizer.raster_scene(frame, scene)                 # buildRasterJobs + queueRasterJobs with the
                                                          # tile binning behind the shim (level 2)

There is no CPU path here: if the CUDA library or a B200-class device is missing, setup raises.
"""
import ctypes
import os

import numpy as np

from . import _build
from .formats import CSpec, CStats, RasterSpec, CANONICAL_SPEC, SHAPE_DTYPE, TILE_DTYPE, ENTRY_DTYPE

_lib = None

STATUS = {0: "GUDNI_OK", -1: "GUDNI_ERR_ARGUMENT", -2: "GUDNI_ERR_NO_DEVICE", -3: "GUDNI_ERR_CUDA",
          -4: "GUDNI_ERR_STATE", -5: "GUDNI_ERR_OOM"}

# every symbol include/gudni_b200.h declares; tests check the library exports all of them
ABI_SYMBOLS = [
    "gudni_b200_init", "gudni_b200_frame_begin", "gudni_b200_frame_strip", "gudni_b200_raster_job",
    "gudni_b200_raster_scene", "gudni_b200_frame_end", "gudni_b200_frame_device_ptr", "gudni_b200_frame_target", "gudni_b200_frame_target_host",
    "gudni_b200_ipc_export_frame", "gudni_b200_ipc_open", "gudni_b200_ipc_close", "gudni_b200_device_alloc",
    "gudni_b200_device_free", "gudni_b200_host_register", "gudni_b200_host_unregister", "gudni_b200_upload", "gudni_b200_download", "gudni_b200_frame_begin_device",
    "gudni_b200_raster_scene_device", "gudni_b200_raster_outlines", "gudni_b200_raster_outlines_device",
    "gudni_b200_debug_strands", "gudni_b200_sync", "gudni_b200_last_frame_ms", "gudni_b200_launch_count",
    "gudni_b200_set_stream", "gudni_b200_debug_selftest", "gudni_b200_debug_enable", "gudni_b200_debug_thread_counts", "gudni_b200_debug_binned",
    "gudni_b200_last_error", "gudni_b200_destroy",
    "gudni_b200_frame_begin_cached", "gudni_b200_raster_scene_cached",
    "gudni_b200_multi_init", "gudni_b200_multi_destroy", "gudni_b200_multi_last_error", "gudni_b200_multi_set_presenting",
    "gudni_b200_multi_canvas", "gudni_b200_multi_frame", "gudni_b200_partition_rows", "gudni_b200_rebalance_rows",
]


class GudniError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{STATUS.get(code, code)}: {message}")
        self.code = code


def load_library():
    """dlopen libgudni_b200.so (built in-tree by `__graft_entry__.build()`).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_CUDA
    if not os.path.exists(path):
        raise GudniError(-2, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
    c = ctypes
    L = ctypes.CDLL(path)
    vp, i32, i64, sz = c.c_void_p, c.c_int, c.c_int64, c.c_size_t
    L.gudni_b200_init.argtypes = [i32, c.POINTER(CSpec), c.POINTER(CSpec), c.POINTER(vp)]
    L.gudni_b200_frame_begin.argtypes = [vp, vp, sz, vp, i32, vp, sz, vp, i32, vp, i32, i32, i32]
    L.gudni_b200_frame_begin_device.argtypes = [vp, vp, sz, vp, i32, vp, sz, vp, i32, vp, i32, i32, i32]
    L.gudni_b200_frame_begin_cached.argtypes = [vp, vp, sz, vp, i32, vp, sz, vp, i32, vp, i32, i32, i32, vp]
    L.gudni_b200_raster_scene_cached.argtypes = [vp, vp, i32, c.c_uint64]
    L.gudni_b200_frame_strip.argtypes = [vp, i32, i32]
    L.gudni_b200_raster_job.argtypes = [vp, vp, i32, vp, i32, i32, i32]
    L.gudni_b200_raster_scene.argtypes = [vp, vp, i32]
    L.gudni_b200_raster_scene_device.argtypes = [vp, vp, i32]
    L.gudni_b200_raster_outlines.argtypes = [vp, vp, i32, vp, i32, vp, i64, vp, i32]
    L.gudni_b200_raster_outlines_device.argtypes = [vp, vp, i32, vp, i32, vp, i64, vp, i32]
    L.gudni_b200_debug_strands.argtypes = [vp, vp, sz, c.POINTER(sz), vp, i64, c.POINTER(i64)]
    L.gudni_b200_frame_end.argtypes = [vp, vp, c.POINTER(CStats)]
    L.gudni_b200_frame_device_ptr.argtypes = [vp, c.POINTER(vp), c.POINTER(sz)]
    L.gudni_b200_frame_target.argtypes = [vp, vp, i32]
    L.gudni_b200_frame_target_host.argtypes = [vp, vp]
    L.gudni_b200_set_stream.argtypes = [vp, vp]
    L.gudni_b200_ipc_export_frame.argtypes = [vp, vp]
    L.gudni_b200_ipc_open.argtypes = [vp, vp, c.POINTER(vp)]
    L.gudni_b200_ipc_close.argtypes = [vp, vp]
    L.gudni_b200_device_alloc.argtypes = [vp, sz, c.POINTER(vp)]
    L.gudni_b200_device_free.argtypes = [vp, vp]
    L.gudni_b200_upload.argtypes = [vp, vp, vp, sz]
    L.gudni_b200_download.argtypes = [vp, vp, vp, sz]
    L.gudni_b200_host_register.argtypes = [vp, vp, sz]
    L.gudni_b200_host_unregister.argtypes = [vp, vp]
    L.gudni_b200_sync.argtypes = [vp]
    L.gudni_b200_last_frame_ms.argtypes = [vp, c.POINTER(c.c_float)]
    L.gudni_b200_launch_count.argtypes = [vp, c.POINTER(i64)]
    L.gudni_b200_debug_selftest.argtypes = [vp, c.c_uint64, c.c_uint64, c.POINTER(c.c_uint64)]
    L.gudni_b200_debug_enable.argtypes = [vp, i32]
    L.gudni_b200_debug_thread_counts.argtypes = [vp, vp, vp, i64, c.POINTER(i64)]
    L.gudni_b200_debug_binned.argtypes = [vp, vp, i64, c.POINTER(i64), vp, i64, c.POINTER(i64)]
    L.gudni_b200_last_error.argtypes = [vp]
    L.gudni_b200_last_error.restype = c.c_char_p
    L.gudni_b200_destroy.argtypes = [vp]
    L.gudni_b200_destroy.restype = None
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class FrameStats:
    def __init__(self, c: CStats):
        for name, _ in CStats._fields_:
            setattr(self, name, getattr(c, name))

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in CStats._fields_}


class DeviceScene:
    """Frame inputs resident in HBM (bench.py's inputs-in-HBM leg)."""

    def __init__(self, rasterizer, scene, entries=None):
        self.r = rasterizer
        self.scene = scene
        self._bufs = []
        entries = scene.entries if entries is None else entries
        self.n_entries = len(entries)
        self.geometry = self._put(scene.geometry)
        self.substances = self._put(np.ascontiguousarray(scene.substances, np.float32))
        self.pictures = self._put(scene.picture_bytes)
        self.picture_uses = self._put(scene.picture_uses)
        self.entries = self._put(entries)

    def _put(self, arr):
        arr = np.ascontiguousarray(arr)
        p = ctypes.c_void_p()
        self.r._check(self.r._L.gudni_b200_device_alloc(self.r._ctx, arr.nbytes, ctypes.byref(p)))
        if arr.nbytes:
            self.r._check(self.r._L.gudni_b200_upload(self.r._ctx, p, arr.ctypes.data, arr.nbytes))
        self._bufs.append(p)
        return p

    def put_outlines(self):
        """Upload the scene's pre-serialisation form (level 3 inputs); returns the four device pointers."""
        self.outline_arrays = tuple(np.ascontiguousarray(a) for a in self.scene.raw)
        self.outline_ptrs = tuple(self._put(a) for a in self.outline_arrays)
        return self.outline_ptrs

    def put_entries(self, entries):
        """Upload one more shape-entry array (e.g. the entries of one chunk of a strip)."""
        return self._put(entries), len(entries)

    def free(self):
        for p in self._bufs:
            self.r._L.gudni_b200_device_free(self.r._ctx, p)
        self._bufs = []


class Rasterizer:
    """Rasterizer (OpenCL/Rasterizer.hs:52-60) holding a gudni_ctx* instead of an OpenCLState."""

    def __init__(self, device=-1, spec: RasterSpec = None):
        self._L = load_library()
        self._ctx = ctypes.c_void_p()
        got = CSpec()
        want = spec.to_c() if spec is not None else None
        rc = self._L.gudni_b200_init(device, ctypes.byref(want) if want is not None else None, ctypes.byref(got),
                                     ctypes.byref(self._ctx))
        if rc != 0:
            raise GudniError(rc, "gudni_b200_init failed (no usable sm_100 device, or a bad spec); "
                                 "there is no CPU fallback")
        self.spec = RasterSpec.from_c(got)   # rasterSpec, read by the tile binning on the Haskell side

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise GudniError(rc, (self._L.gudni_b200_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx:
            self._L.gudni_b200_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's call sequence -----------------------------------------------------------
    def frame_begin(self, scene, frame_count=0, generations=None):
        """`generations`: None, or (geometry, substances, pictures, picture_uses, entries) counters for the input
        cache (gudni_b200_frame_begin_cached): an input whose counter is nonzero and unchanged is not uploaded again."""
        g = np.ascontiguousarray(scene.geometry)
        s = np.ascontiguousarray(scene.substances, np.float32)
        p = np.ascontiguousarray(scene.picture_bytes)
        u = np.ascontiguousarray(scene.picture_uses)
        bg = np.ascontiguousarray(scene.background, np.float32)
        self._keep = (g, s, p, u, bg)
        gen = (ctypes.c_uint64 * 5)(*generations) if generations is not None else None
        self._check(self._L.gudni_b200_frame_begin_cached(self._ctx, _ptr(g), g.nbytes, _ptr(s), len(s), _ptr(p), p.nbytes,
                                                          _ptr(u), len(u), bg.ctypes.data, scene.width, scene.height,
                                                          frame_count, gen))
        self._dims = (scene.height, scene.width)
        self._rows = (0, scene.height)

    def frame_begin_device(self, dscene: DeviceScene, frame_count=0):
        sc = dscene.scene
        bg = np.ascontiguousarray(sc.background, np.float32)
        self._check(self._L.gudni_b200_frame_begin_device(
            self._ctx, dscene.geometry, sc.geometry.nbytes, dscene.substances, len(sc.substances), dscene.pictures,
            sc.picture_bytes.nbytes, dscene.picture_uses, len(sc.picture_uses), bg.ctypes.data, sc.width, sc.height,
            frame_count))
        self._dims = (sc.height, sc.width)
        self._rows = (0, sc.height)

    def frame_strip(self, row_begin, row_end):
        self._check(self._L.gudni_b200_frame_strip(self._ctx, row_begin, row_end))
        self._rows = (row_begin, row_end)

    def raster_job(self, job, job_index=0):
        """`raster` (OpenCL/CallKernels.hs:182-206) for one RasterJob."""
        shapes = np.ascontiguousarray(job.shapes, SHAPE_DTYPE)
        tiles = np.ascontiguousarray(job.tiles, TILE_DTYPE)
        self._check(self._L.gudni_b200_raster_job(self._ctx, _ptr(shapes), len(shapes), _ptr(tiles), len(tiles),
                                                  job.columns, job_index))

    def raster_entries(self, entries, generation=0):
        e = np.ascontiguousarray(entries)
        self._check(self._L.gudni_b200_raster_scene_cached(self._ctx, _ptr(e), len(e), generation))

    def raster_entries_device(self, dev_entries, n):
        self._check(self._L.gudni_b200_raster_scene_device(self._ctx, dev_entries, n))

    def frame_end(self, out=None, want_image=True):
        rows = self._rows[1] - self._rows[0]
        if want_image and out is None:
            out = np.empty((rows, self._dims[1]), dtype=np.uint32)
        st = CStats()
        self._check(self._L.gudni_b200_frame_end(self._ctx, out.ctypes.data if out is not None else None,
                                                 ctypes.byref(st)))
        return out, FrameStats(st)

    def queue_raster_jobs(self, frame_count, scene, jobs, out=None):
        """queueRasterJobs (OpenCL/CallKernels.hs:218-242): frame constants, then every job."""
        self.frame_begin(scene, frame_count)
        for index, job in enumerate(jobs):
            self.raster_job(job, index)
        return self.frame_end(out)

    def raster_scene(self, frame_count, scene, out=None, rows=None, generations=None):
        """buildRasterJobs + queueRasterJobs with the tile binning done on the GPU (level 2)."""
        self.frame_begin(scene, frame_count, generations)
        entries = scene.entries
        if rows is not None:
            self.frame_strip(*rows)
            entries = scene.subset_rows(*rows)
        self.raster_entries(entries, generations[4] if generations is not None else 0)
        return self.frame_end(out)

    def frame_begin_outlines(self, scene, frame_count=0):
        """frame_begin for level 3: everything but the geometry heap, which the library builds itself."""
        s = np.ascontiguousarray(scene.substances, np.float32)
        p = np.ascontiguousarray(scene.picture_bytes)
        u = np.ascontiguousarray(scene.picture_uses)
        bg = np.ascontiguousarray(scene.background, np.float32)
        self._keep = (s, p, u, bg)
        self._check(self._L.gudni_b200_frame_begin(self._ctx, None, 0, _ptr(s), len(s), _ptr(p), p.nbytes, _ptr(u), len(u),
                                                   bg.ctypes.data, scene.width, scene.height, frame_count))
        self._dims = (scene.height, scene.width)
        self._rows = (0, scene.height)

    def raster_outlines(self, frame_count, scene, out=None):
        """Level 3: onShape's geometry work (transform, box, cull, strands: Raster/Serialize.hs:148-177,
        Raster/Strand.hs:153-178) + tile binning + rasterization, all behind the shim.  `scene.raw` is the
        scene before serialisation: (shapes, outlines, curve pairs, transforms)."""
        self.frame_begin_outlines(scene, frame_count)
        shapes, outlines, pairs, transforms = (np.ascontiguousarray(a) for a in scene.raw)
        self._check(self._L.gudni_b200_raster_outlines(self._ctx, _ptr(shapes), len(shapes), _ptr(outlines), len(outlines),
                                                       _ptr(pairs), len(pairs), _ptr(transforms), len(transforms)))
        return self.frame_end(out)

    def frame_begin_device_outlines(self, dscene: DeviceScene, frame_count=0):
        sc = dscene.scene
        bg = np.ascontiguousarray(sc.background, np.float32)
        self._check(self._L.gudni_b200_frame_begin_device(
            self._ctx, None, 0, dscene.substances, len(sc.substances), dscene.pictures, sc.picture_bytes.nbytes,
            dscene.picture_uses, len(sc.picture_uses), bg.ctypes.data, sc.width, sc.height, frame_count))
        self._dims = (sc.height, sc.width)
        self._rows = (0, sc.height)

    def raster_outlines_device(self, dscene: DeviceScene):
        sh, ol, pr, tr = dscene.outline_ptrs
        a = dscene.outline_arrays
        self._check(self._L.gudni_b200_raster_outlines_device(self._ctx, sh, len(a[0]), ol, len(a[1]), pr, len(a[2]), tr,
                                                              len(a[3])))

    def debug_strands(self):
        """Geometry heap and shape entries level 3 built for the last frame."""
        nb, ne = ctypes.c_size_t(), ctypes.c_int64()
        self._check(self._L.gudni_b200_debug_strands(self._ctx, None, 0, ctypes.byref(nb), None, 0, ctypes.byref(ne)))
        geometry = np.empty(nb.value, np.uint8)
        entries = np.empty(ne.value, ENTRY_DTYPE)
        self._check(self._L.gudni_b200_debug_strands(self._ctx, _ptr(geometry), geometry.nbytes, ctypes.byref(nb),
                                                     _ptr(entries), len(entries), ctypes.byref(ne)))
        return geometry, entries

    # -- device-side access ----------------------------------------------------------------------
    def frame_device_ptr(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._L.gudni_b200_frame_device_ptr(self._ctx, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def frame_target(self, dev_ptr, row_origin=0):
        self._check(self._L.gudni_b200_frame_target(self._ctx, ctypes.c_void_p(dev_ptr) if dev_ptr else None,
                                                    row_origin))

    def frame_target_host(self, array):
        """The kernels store the frame straight into `array` (uint32, height x width, page-locked with host_register);
        frame_end(out=array) then copies nothing.  None restores the context's own frame buffer."""
        self._check(self._L.gudni_b200_frame_target_host(self._ctx, array.ctypes.data if array is not None else None))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        self._check(self._L.gudni_b200_set_stream(self._ctx, ctypes.c_void_p(cuda_stream) if cuda_stream else None))

    def host_register(self, array):
        """Page-lock a numpy array that will be passed again and again (see gudni_b200_host_register)."""
        if array.nbytes:
            self._check(self._L.gudni_b200_host_register(self._ctx, array.ctypes.data, array.nbytes))

    def host_unregister(self, array):
        if array.nbytes:
            self._check(self._L.gudni_b200_host_unregister(self._ctx, array.ctypes.data))

    def sync(self):
        self._check(self._L.gudni_b200_sync(self._ctx))

    def last_frame_ms(self):
        ms = ctypes.c_float()
        self._check(self._L.gudni_b200_last_frame_ms(self._ctx, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        n = ctypes.c_int64()
        self._check(self._L.gudni_b200_launch_count(self._ctx, ctypes.byref(n)))
        return n.value

    # -- parity taps -----------------------------------------------------------------------------
    def debug_enable(self, on=True):
        self._check(self._L.gudni_b200_debug_enable(self._ctx, int(on)))

    def debug_selftest(self, n=1 << 28, seed=1):
        bad = ctypes.c_uint64()
        self._check(self._L.gudni_b200_debug_selftest(self._ctx, n, seed, ctypes.byref(bad)))
        return bad.value

    def debug_thread_counts(self):
        n = ctypes.c_int64()
        self._check(self._L.gudni_b200_debug_thread_counts(self._ctx, None, None, 0, ctypes.byref(n)))
        a = np.empty(n.value, np.int32)
        b = np.empty(n.value, np.int32)
        self._check(self._L.gudni_b200_debug_thread_counts(self._ctx, _ptr(a), _ptr(b), n.value, ctypes.byref(n)))
        return a, b

    def debug_binned(self):
        nt, ns = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._L.gudni_b200_debug_binned(self._ctx, None, 0, ctypes.byref(nt), None, 0, ctypes.byref(ns)))
        tiles = np.empty(nt.value, TILE_DTYPE)
        shapes = np.empty(ns.value, SHAPE_DTYPE)
        self._check(self._L.gudni_b200_debug_binned(self._ctx, _ptr(tiles), nt.value, ctypes.byref(nt), _ptr(shapes),
                                                    ns.value, ctypes.byref(ns)))
        return tiles, shapes


def setup_rasterizer(device=-1, spec: RasterSpec = None) -> Rasterizer:
    """setupOpenCL (OpenCL/Setup.hs:102): pick the device, fix the RasterSpec, build the context."""
    return Rasterizer(device, spec)
