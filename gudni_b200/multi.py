"""Several GPUs of one box in ONE process, over the C ABI (gudni_b200_multi_*): what a Haskell caller would bind
to render a 16K x 16K canvas on the whole box.  The reference has no counterpart (one OpenCLState = one device,
OpenCL/Setup.hs:118-120).  The one-process-per-GPU harness the benchmark contract asks for lives in strips.py."""
import ctypes

import numpy as np

from .formats import CSpec, CStats, RasterSpec
from .raster import GudniError, load_library

MAX_DEVICES = 16


class CMultiStats(ctypes.Structure):
    """gudni_multi_stats."""
    _fields_ = [("n_devices", ctypes.c_int32), ("device", ctypes.c_int32 * MAX_DEVICES),
                ("row_begin", ctypes.c_int32 * MAX_DEVICES), ("row_end", ctypes.c_int32 * MAX_DEVICES),
                ("ms_device", ctypes.c_float * MAX_DEVICES), ("ms_frame", ctypes.c_float),
                ("ms_gather_exposed", ctypes.c_float), ("total", CStats)]


def _bind(L):
    c = ctypes
    vp, i32, sz = c.c_void_p, c.c_int, c.c_size_t
    L.gudni_b200_multi_init.argtypes = [i32, c.POINTER(i32), c.POINTER(CSpec), c.POINTER(CSpec), c.POINTER(vp)]
    L.gudni_b200_multi_destroy.argtypes = [vp]
    L.gudni_b200_multi_destroy.restype = None
    L.gudni_b200_multi_last_error.argtypes = [vp]
    L.gudni_b200_multi_last_error.restype = c.c_char_p
    L.gudni_b200_multi_set_presenting.argtypes = [vp, i32]
    L.gudni_b200_multi_canvas.argtypes = [vp, c.POINTER(vp), c.POINTER(i32)]
    L.gudni_b200_multi_frame.argtypes = [vp, vp, sz, vp, i32, vp, sz, vp, i32, vp, i32, i32, i32, vp, i32, vp, vp,
                                         c.POINTER(CMultiStats)]
    L.gudni_b200_partition_rows.argtypes = [vp, i32, i32, i32, i32, i32, c.POINTER(i32)]
    L.gudni_b200_rebalance_rows.argtypes = [c.POINTER(i32), c.POINTER(c.c_double), i32, i32, i32, c.POINTER(i32)]
    return L


def partition_rows(entries, width, height, tile_rows, n_devices):
    """gudni_b200_partition_rows: first cut of the canvas, from the shapes' boxes (no GPU needed)."""
    L = _bind(load_library())
    e = np.ascontiguousarray(entries)
    out = (ctypes.c_int * (2 * n_devices))()
    rc = L.gudni_b200_partition_rows(e.ctypes.data if len(e) else None, len(e), width, height, tile_rows, n_devices, out)
    if rc:
        raise GudniError(rc, "partition_rows")
    return [(out[2 * d], out[2 * d + 1]) for d in range(n_devices)]


def rebalance_rows(rows, ms, height, tile_rows):
    """gudni_b200_rebalance_rows: the feedback cut from last frame's strips and per-device times."""
    L = _bind(load_library())
    n = len(rows)
    rin = (ctypes.c_int * (2 * n))(*[v for r in rows for v in r])
    t = (ctypes.c_double * n)(*ms)
    out = (ctypes.c_int * (2 * n))()
    rc = L.gudni_b200_rebalance_rows(rin, t, n, height, tile_rows, out)
    if rc:
        raise GudniError(rc, "rebalance_rows")
    return [(out[2 * d], out[2 * d + 1]) for d in range(n)]


class MultiStats:
    def __init__(self, c: CMultiStats):
        n = c.n_devices
        self.n_devices = n
        self.devices = list(c.device[:n])
        self.rows = [(c.row_begin[d], c.row_end[d]) for d in range(n)]
        self.ms_device = list(c.ms_device[:n])
        self.ms_frame = c.ms_frame
        self.ms_gather_exposed = c.ms_gather_exposed
        self.total = {name: getattr(c.total, name) for name, _ in CStats._fields_}


class MultiRasterizer:
    """setup_rasterizer for several devices: gudni_b200_multi_init.  `devices` may repeat a device."""

    def __init__(self, devices, spec: RasterSpec = None):
        self._L = _bind(load_library())
        self._h = ctypes.c_void_p()
        got = CSpec()
        want = spec.to_c() if spec is not None else None
        arr = (ctypes.c_int * len(devices))(*devices)
        rc = self._L.gudni_b200_multi_init(len(devices), arr, ctypes.byref(want) if want is not None else None, ctypes.byref(got),
                                           ctypes.byref(self._h))
        if rc:
            raise GudniError(rc, "gudni_b200_multi_init failed (no usable sm_100 device, or a bad spec)")
        self.spec = RasterSpec.from_c(got)
        self.n_devices = len(devices)

    def close(self):
        if self._h:
            self._L.gudni_b200_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_presenting(self, device_index):
        rc = self._L.gudni_b200_multi_set_presenting(self._h, device_index)
        if rc:
            raise GudniError(rc, "set_presenting")

    def canvas(self):
        p, d = ctypes.c_void_p(), ctypes.c_int()
        rc = self._L.gudni_b200_multi_canvas(self._h, ctypes.byref(p), ctypes.byref(d))
        if rc:
            raise GudniError(rc, (self._L.gudni_b200_multi_last_error(self._h) or b"").decode())
        return p.value, d.value

    def frame(self, frame_count, scene, out=None, want_image=True, generations=None):
        """queueRasterJobs for the whole canvas: returns (bitmap or None, MultiStats)."""
        g = np.ascontiguousarray(scene.geometry)
        s = np.ascontiguousarray(scene.substances, np.float32)
        p = np.ascontiguousarray(scene.picture_bytes)
        u = np.ascontiguousarray(scene.picture_uses)
        e = np.ascontiguousarray(scene.entries)
        bg = np.ascontiguousarray(scene.background, np.float32)
        if want_image and out is None:
            out = np.empty((scene.height, scene.width), dtype=np.uint32)
        gen = (ctypes.c_uint64 * 5)(*generations) if generations is not None else None
        st = CMultiStats()
        ptr = lambda a: a.ctypes.data if a is not None and a.size else None
        rc = self._L.gudni_b200_multi_frame(self._h, ptr(g), g.nbytes, ptr(s), len(s), ptr(p), p.nbytes, ptr(u), len(u),
                                            bg.ctypes.data, scene.width, scene.height, frame_count, ptr(e), len(e), gen,
                                            out.ctypes.data if out is not None else None, ctypes.byref(st))
        if rc:
            raise GudniError(rc, (self._L.gudni_b200_multi_last_error(self._h) or b"").decode())
        return out, MultiStats(st)
