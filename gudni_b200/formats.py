"""Wire formats of the hot path as numpy dtypes / ctypes structs (little endian).

Layouts follow include/gudni_b200.h, which cites the Haskell `StorableM` instance and the
`Kernels.cl` struct each one mirrors.
"""
import ctypes
from dataclasses import dataclass

import numpy as np

# Shape GeoReference — Raster/Types.hs:159-168, Kernels.cl:318-321 (16 B)
SHAPE_DTYPE = np.dtype([("tag", "<u8"), ("geo_start", "<u4"), ("num_strands", "<u4")])
# Tile (Slice, Int) — Raster/Types.hs:176-198, Kernels.cl:335-341 (32 B)
TILE_DTYPE = np.dtype([("left", "<i4"), ("top", "<i4"), ("right", "<i4"), ("bottom", "<i4"),
                       ("h_depth", "<i2"), ("v_depth", "<i2"), ("column_allocation", "<i4"),
                       ("shape_start", "<u4"), ("shape_count", "<u4")])
# Shape ShapeEntry (un-binned), include/gudni_b200.h gudni_shape_entry (32 B)
ENTRY_DTYPE = np.dtype([("tag", "<u8"), ("geo_start", "<u4"), ("num_strands", "<u4"),
                        ("left", "<f4"), ("top", "<f4"), ("right", "<f4"), ("bottom", "<f4")])
# PictureUsage PictureMemoryReference — Figure/Picture.hs:183-199, Kernels.cl:349-354 (24 B)
PICTURE_USE_DTYPE = np.dtype([("translate_x", "<f4"), ("translate_y", "<f4"), ("width", "<i4"),
                              ("height", "<i4"), ("mem_offset", "<u4"), ("scale", "<f4")])
# level 3 inputs, include/gudni_b200.h gudni_outline_shape / gudni_outline / gudni_curve_pair / gudni_transform
OUTLINE_SHAPE_DTYPE = np.dtype([("tag", "<u8"), ("first_outline", "<u4"), ("n_outlines", "<u4"),
                                ("first_transform", "<u4"), ("n_transforms", "<u4"), ("reserved", "<u4", (2,))])
OUTLINE_DTYPE = np.dtype([("first_pair", "<u4"), ("n_pairs", "<u4")])
CURVE_PAIR_DTYPE = np.dtype([("on_x", "<f4"), ("on_y", "<f4"), ("off_x", "<f4"), ("off_y", "<f4")])
TRANSFORM_DTYPE = np.dtype([("kind", "<u4"), ("a", "<f4"), ("b", "<f4"), ("reserved", "<u4")])
assert OUTLINE_SHAPE_DTYPE.itemsize == 32 and OUTLINE_DTYPE.itemsize == 8
assert CURVE_PAIR_DTYPE.itemsize == 16 and TRANSFORM_DTYPE.itemsize == 16
assert SHAPE_DTYPE.itemsize == 16 and TILE_DTYPE.itemsize == 32
assert ENTRY_DTYPE.itemsize == 32 and PICTURE_USE_DTYPE.itemsize == 24

TAG_SUBSTANCE_SOLID = 0x8000000000000000
TAG_SUBSTANCE_PICTURE = 0x4000000000000000
TAG_COMPOUND_ADD = 0x2000000000000000
TAG_COMPOUND_SUBTRACT = 0x3000000000000000
TAG_SUBSTANCEID_MASK = 0x0FFFFFFFFFFFFFFF


class CSpec(ctypes.Structure):
    """gudni_spec — RasterSpec, OpenCL/Rasterizer.hs:37-50."""
    _fields_ = [("max_tile_size", ctypes.c_int32), ("threads_per_tile", ctypes.c_int32),
                ("max_tiles_per_call", ctypes.c_int32), ("max_thresholds", ctypes.c_int32),
                ("max_strands_per_tile", ctypes.c_int32), ("max_shapes", ctypes.c_int32)]


class CStats(ctypes.Structure):
    """gudni_stats."""
    _fields_ = [("n_tiles", ctypes.c_int64), ("n_shape_refs", ctypes.c_int64),
                ("n_thresholds", ctypes.c_int64), ("n_spilled_threads", ctypes.c_int64),
                ("n_overflow_threads", ctypes.c_int64), ("algorithmic_bytes", ctypes.c_int64),
                ("ms_upload", ctypes.c_float), ("ms_bin", ctypes.c_float),
                ("ms_raster", ctypes.c_float), ("ms_download", ctypes.c_float),
                ("ms_strands", ctypes.c_float), ("reserved", ctypes.c_float)]


@dataclass(frozen=True)
class RasterSpec:
    """RasterSpec (OpenCL/Rasterizer.hs:37-50).  The reference derives it from the OpenCL device
    (OpenCL/Setup.hs:71-87); the canonical values are fixed in BASELINE.md §3."""
    max_tile_size: int = 256
    threads_per_tile: int = 256
    max_tiles_per_call: int = 256
    max_thresholds: int = 1024
    max_strands_per_tile: int = 1022
    max_shapes: int = 127

    def to_c(self) -> CSpec:
        return CSpec(self.max_tile_size, self.threads_per_tile, self.max_tiles_per_call,
                     self.max_thresholds, self.max_strands_per_tile, self.max_shapes)

    @staticmethod
    def from_c(c: CSpec) -> "RasterSpec":
        return RasterSpec(c.max_tile_size, c.threads_per_tile, c.max_tiles_per_call,
                          c.max_thresholds, c.max_strands_per_tile, c.max_shapes)


CANONICAL_SPEC = RasterSpec()
