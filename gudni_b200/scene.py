"""Scene construction for tests and bench.py — HARNESS, not product.

In the reference the Haskell front end (Graphics.Gudni.Layout / Figure / Raster.Serialize) turns a
`Scene` into the byte buffers that cross the drop-in boundary.  GHC is not in
this image, so csrc/host/ restates that producer and this module drives it over ctypes.
"""
import ctypes
import os

import numpy as np

from . import _build
from .formats import (ENTRY_DTYPE, PICTURE_USE_DTYPE, OUTLINE_SHAPE_DTYPE, OUTLINE_DTYPE, CURVE_PAIR_DTYPE,
                      TRANSFORM_DTYPE)

_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_HOST
    if not os.path.exists(path):
        _build.build_host()
    lib = ctypes.CDLL(path)
    c = ctypes
    fp = c.POINTER(c.c_float)
    lib.gs_scene_new.restype = c.c_void_p
    lib.gs_scene_new.argtypes = [c.c_int, c.c_int, fp]
    lib.gs_scene_free.argtypes = [c.c_void_p]
    lib.gs_add_solid.argtypes = [c.c_void_p] + [c.c_float] * 4
    lib.gs_add_picture.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int]
    lib.gs_add_picture_substance.argtypes = [c.c_void_p, c.c_int, c.c_float, c.c_float, c.c_float]
    lib.gs_add_shape.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, fp, c.POINTER(c.c_int), c.c_int]
    lib.gs_add_rectangle.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_float, c.c_float, fp, c.c_int]
    lib.gs_add_circle.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, fp, c.c_int]
    lib.gs_unit_circle.argtypes = [fp, c.c_int]
    lib.gs_arc.argtypes = [c.c_float, fp, c.c_int]
    lib.gs_add_fuzzy_circles.argtypes = [c.c_void_p, c.c_int] + [c.c_float] * 4 + [c.c_uint64]
    for name in ("gs_geometry", "gs_picture_bytes"):
        getattr(lib, name).restype = c.c_void_p
        getattr(lib, name).argtypes = [c.c_void_p, c.POINTER(c.c_size_t)]
    for name in ("gs_entries", "gs_substances", "gs_picture_uses", "gs_raw_shapes", "gs_raw_outlines", "gs_raw_pairs",
                 "gs_raw_transforms"):
        getattr(lib, name).restype = c.c_void_p
        getattr(lib, name).argtypes = [c.c_void_p, c.POINTER(c.c_int)]
    lib.gs_info.argtypes = [c.c_void_p, c.POINTER(c.c_int), c.POINTER(c.c_int), fp,
                            c.POINTER(c.c_int64), c.POINTER(c.c_int64)]
    _lib = lib
    return lib


def _stack(transforms):
    """transforms: outermost first, as the client would write them:
    ("translate", x, y) | ("scale", s) | ("rotate", turns)."""
    kinds = {"translate": 0, "scale": 1, "rotate": 2}
    rows = []
    for t in transforms:
        rows.append([kinds[t[0]], t[1], t[2] if len(t) > 2 else 0.0])
    arr = np.asarray(rows, dtype=np.float32).reshape(-1, 3)
    return arr


class FrozenScene:
    """The byte buffers that cross the boundary, as numpy arrays."""

    def __init__(self, width, height, background, geometry, entries, substances, picture_bytes,
                 picture_uses, culled=0, curves=0, name="", raw=None):
        self.width, self.height = int(width), int(height)
        self.background = np.asarray(background, dtype=np.float32)
        self.geometry = geometry
        self.entries = entries
        self.substances = substances
        self.picture_bytes = picture_bytes
        self.picture_uses = picture_uses
        self.culled, self.curves = int(culled), int(curves)
        self.name = name
        # the scene before serialisation — (shapes, outlines, pairs, transforms) as level 3 takes them
        self.raw = raw

    @property
    def n_shapes(self):
        return len(self.entries)

    def subset_rows(self, row_begin, row_end):
        """Entries whose box touches canvas rows [row_begin, row_end) — what a strip rank bins."""
        e = self.entries
        keep = (e["top"] < np.float32(row_end)) & (e["bottom"] > np.float32(row_begin))
        return e[keep]


class SceneBuilder:
    """Mutable scene; mirrors what buildOverScene (Raster/Serialize.hs:266-273) accumulates."""

    def __init__(self, width, height, background=(1.0, 1.0, 1.0, 1.0), name=""):
        self._lib = _load()
        bg = (ctypes.c_float * 4)(*background)
        self._h = ctypes.c_void_p(self._lib.gs_scene_new(int(width), int(height), bg))
        self.name = name

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gs_scene_free(self._h)
            self._h = None

    def solid(self, r, g, b, a=1.0):
        return self._lib.gs_add_solid(self._h, r, g, b, a)

    def picture(self, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        return self._lib.gs_add_picture(self._h, rgba.ctypes.data_as(ctypes.c_void_p), w, h)

    def picture_substance(self, picture, translate=(0.0, 0.0), scale=1.0):
        return self._lib.gs_add_picture_substance(self._h, picture, translate[0], translate[1], scale)

    def rectangle(self, substance, w, h, transforms=(), subtract=False):
        st = _stack(transforms)
        self._lib.gs_add_rectangle(self._h, substance, int(subtract), w, h,
                                   st.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(st))

    def circle(self, substance, transforms=(), subtract=False, is_picture=False):
        st = _stack(transforms)
        self._lib.gs_add_circle(self._h, substance, int(is_picture), int(subtract),
                                st.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(st))

    def shape(self, substance, outlines, subtract=False, is_picture=False):
        """outlines: list of (n,4) float32 arrays of curve pairs (on.x, on.y, off.x, off.y),
        already transformed."""
        sizes = (ctypes.c_int * len(outlines))(*[len(o) for o in outlines])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(o, np.float32).reshape(-1, 4) for o in outlines]))
        self._lib.gs_add_shape(self._h, substance, int(is_picture), int(subtract),
                               flat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), sizes, len(outlines))

    def fuzzy_circles(self, n, range_w, range_h, min_rad, max_rad, seed):
        self._lib.gs_add_fuzzy_circles(self._h, n, range_w, range_h, min_rad, max_rad, seed)

    def freeze(self) -> FrozenScene:
        lib, h = self._lib, self._h
        nb = ctypes.c_size_t()
        n = ctypes.c_int()

        def copy(ptr, count, dtype):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            nbytes = count * np.dtype(dtype).itemsize
            return np.frombuffer(ctypes.string_at(ptr, nbytes), dtype=dtype).copy()

        p = lib.gs_geometry(h, ctypes.byref(nb))
        geometry = copy(p, nb.value, np.uint8)
        p = lib.gs_entries(h, ctypes.byref(n))
        entries = copy(p, n.value, ENTRY_DTYPE)
        p = lib.gs_substances(h, ctypes.byref(n))
        substances = copy(p, n.value * 4, np.float32).reshape(-1, 4)
        p = lib.gs_picture_bytes(h, ctypes.byref(nb))
        picture_bytes = copy(p, nb.value, np.uint8)
        p = lib.gs_picture_uses(h, ctypes.byref(n))
        picture_uses = copy(p, n.value, PICTURE_USE_DTYPE)
        raw = []
        for getter, dtype in (("gs_raw_shapes", OUTLINE_SHAPE_DTYPE), ("gs_raw_outlines", OUTLINE_DTYPE),
                              ("gs_raw_pairs", CURVE_PAIR_DTYPE), ("gs_raw_transforms", TRANSFORM_DTYPE)):
            p = getattr(lib, getter)(h, ctypes.byref(n))
            raw.append(copy(p, n.value, dtype))
        w, hh = ctypes.c_int(), ctypes.c_int()
        bg = (ctypes.c_float * 4)()
        culled, curves = ctypes.c_int64(), ctypes.c_int64()
        lib.gs_info(h, ctypes.byref(w), ctypes.byref(hh), bg, ctypes.byref(culled), ctypes.byref(curves))
        return FrozenScene(w.value, hh.value, list(bg), geometry, entries, substances, picture_bytes,
                           picture_uses, culled.value, curves.value, self.name, tuple(raw))


def unit_circle_pairs():
    lib = _load()
    buf = (ctypes.c_float * (4 * 64))()
    n = lib.gs_unit_circle(buf, 64)
    return np.frombuffer(buf, dtype=np.float32, count=4 * n).reshape(n, 4).copy()
