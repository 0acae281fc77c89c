"""Helpers shared by the hand-worked tile-tree tests: fixture -> entries / expected tile records."""
import numpy as np

from gudni_b200.formats import ENTRY_DTYPE, TILE_DTYPE, RasterSpec
from golden.tiletree_handworked import CASES   # noqa: F401  (re-exported)


def spec_of(case):
    return RasterSpec(**case["spec"])


def entries_of(case, geo_stride=8):
    """Shape entries in insertion order.  tag carries the shape number (so a shape list can be read back), geo_start a
    distinct 16-byte offset per shape."""
    e = np.zeros(len(case["shapes"]), ENTRY_DTYPE)
    for i, (l, t, r, b, strands) in enumerate(case["shapes"]):
        e[i] = (0x8000000000000000 | i, geo_stride * i, strands, l, t, r, b)
    return e


def expected_tiles(case):
    """(tiles in traversal order with shape slices rebased to one frame-wide list, shape numbers of that list)."""
    tiles = np.zeros(len(case["leaves"]), TILE_DTYPE)
    numbers = []
    for k, (l, t, r, b, hd, vd, shapes) in enumerate(case["leaves"]):
        tiles[k] = (l, t, r, b, hd, vd, 0, len(numbers), len(shapes))
        numbers += shapes
    return tiles, np.asarray(numbers, np.int64)


class BoxesOnly:
    """What oracle.build_raster_jobs reads of a scene."""

    def __init__(self, case):
        self.width, self.height = case["canvas"]
        self.entries = entries_of(case)
