"""Hand-worked strand fixtures: what `outlineToStrands` must put into the geometry heap for four small outlines,
derived on paper from the Haskell.  Nothing here is computed by a strand builder: the expected points are literals
and every step cites /root/reference/src/Graphics/Gudni/Raster/Strand.hs (ST), Deknob.hs (DK), ReorderTable.hs (RT),
Figure/Outline.hs (OL).  tests/test_strands_handworked.py holds the harness's restatement (csrc/host/strand.hpp) and the
kernels' per-shape logic (csrc/strand_build.cuh) to them on the CPU, tests/test_gpu_strands.py the GPU kernels.

An outline is a list of curve pairs (on-curve point, control point).  The rules, once:
  S1  pairsToBeziers (ST:112-127): Bézier i = (on_i, off_i, on_{i+1}), the last one closing to on_0.
  S2  replaceKnobs (DK:84-108): a Bézier whose control lies strictly left of BOTH ends, or strictly right of both, is
      split in two at the point findSplit returns (DK:61-79: bisection on the parameter from t = 1/2; it stops at once
      when neither half-control is beyond the on-curve point); every other Bézier is kept.
  S3  splitIntoStrands (ST:74-102): consecutive Béziers join a run iff both go the same way horizontally and neither
      is vertical (compare start.x end.x equal and not EQ); the fold leaves the LAST run in `acc`, and the result is
      `acc cons strands`: the last run comes FIRST, the others in order.  Runs never wrap around the outline's start.
  S4  splitTooLarge (ST:104-109, called with sectionSize div 2 = 16, ST:178 and mAXsECTIONsIZE = 32): a run of more
      than 16 Béziers is cut into pieces of 16 from its front.
  S5  reverseIfBackwards (ST:137-143): a run whose first point lies right of its last point is reversed (order of the
      Béziers and the ends of each).
  S6  beziersToPoints (ST:130-134): start and control of every Bézier, then the last end: 2n+1 points.
  S7  reorder (ST:147-150, RT:92-110): point i of the strand in memory is point row[i] of S6's list, where
      row = [2n, 0, 1] ++ map (+2) (makeTreeRow (2n-2)): right end, left end, first control, then the remaining
      (on-curve, control) pairs in breadth-first order of the balanced search tree over the n-1 interior on-curve
      points (RT:48-87).
  S8  in the heap a strand is a 8-byte header (u16 size = 2n+2 in 8-byte units, rest zero) followed by the points
      (ST:180-200); the shape's box covers on-curve AND control points (OL:124,133).
"""
import numpy as np


def _mid(p, q):
    return (0.5 * p[0] + 0.5 * q[0], 0.5 * p[1] + 0.5 * q[1])


def _straight(points):
    """curve pairs of a polygon: the control of a straight segment is its midpoint (Figure/Outline.hs:95-104)"""
    return [(p, _mid(p, points[(i + 1) % len(points)])) for i, p in enumerate(points)]


CASES = {}

# ---- H1: a knob ---------------------------------------------------------------------------------------------------------
# pairs: on (10,10) off (12,11); on (10,12) off (10,11).
# S1: B0 = (10,10) (12,11) (10,12);  B1 = (10,12) (10,11) (10,10).
# S2: B0's control x = 12 is right of both ends (10, 10): a right-bulging knob.  findSplit isRightOf 1/2 (DK:96,62-79):
#     search 0 1 1/2: mid0 = between 1/2 v0 control = (11, 10.5); mid1 = between 1/2 control v1 = (11, 11.5);
#     onCurve = between 1/2 mid0 mid1 = (11, 11).  top - bottom = 1 > iota; mid1 isRightOf onCurve: 11 > 11 false;
#     mid0 isRightOf onCurve: false; otherwise -> Bez mid0 onCurve mid1.  fixKnob returns
#     b0 = (10,10) (11,10.5) (11,11) and b1 = (11,11) (11,11.5) (10,12).
#     B1: control x = 10 is neither left nor right of its ends: kept as b2 = (10,12) (10,11) (10,10).
# S3: b0 goes right (LT), b1 goes left (GT): not connectable; b2 is vertical (EQ): not connectable.
#     fold: acc [b0] -> acc [b1], strands [[b0]] -> acc [b2], strands [[b0],[b1]].  Result [[b2],[b0],[b1]].
# S5: [b2] 10 > 10 false: kept.  [b0] kept.  [b1] starts at x 11, ends at x 10: reversed -> (10,12) (11,11.5) (11,11).
# S6/S7 (n = 1: row [2,0,1]): [end, start, control] each.
CASES["H1_knob"] = dict(
    pairs=[((10, 10), (12, 11)), ((10, 12), (10, 11))],
    strands=[
        [(10, 10), (10, 12), (10, 11)],
        [(11, 11), (10, 10), (11, 10.5)],
        [(11, 11), (10, 12), (11, 11.5)],
    ],
    box=(10, 10, 12, 12),
)

# ---- H1L: the mirror image, a left-bulging knob ------------------------------------------------------------------------
# pairs: on (10,10) off (8,11); on (10,12) off (10,11).  B0's control x = 8 is left of both ends: findSplit isLeftOf:
# mid0 = (9,10.5), mid1 = (9,11.5), onCurve = (9,11); neither 9 < 9: split there.  b0 = (10,10) (9,10.5) (9,11) goes
# left, b1 = (9,11) (9,11.5) (10,12) goes right, b2 vertical.  Result [[b2],[b0],[b1]]; b0 reversed (10 > 9):
# (9,11) (9,10.5) (10,10).
CASES["H1L_left_knob"] = dict(
    pairs=[((10, 10), (8, 11)), ((10, 12), (10, 11))],
    strands=[
        [(10, 10), (10, 12), (10, 11)],
        [(10, 10), (9, 11), (9, 10.5)],
        [(10, 12), (9, 11), (9, 11.5)],
    ],
    box=(8, 10, 10, 12),
)

# ---- H2: a triangle walked clockwise from its right corner: a two-Bézier run that has to be reversed ----------------
# points (14,14) (12,10) (10,14), straight sides.  S1: b0 = (14,14) (13,12) (12,10) goes left, b1 = (12,10) (11,12)
# (10,14) goes left, b2 = (10,14) (12,14) (14,14) goes right.  S3: b0, b1 connectable (GT, GT); b2 not: acc [b2],
# strands [[b0,b1]].  Result [[b2],[b0,b1]].
# [b2]: [end (14,14), start (10,14), control (12,14)].
# [b0,b1]: first point x 14 > last point x 10: reversed to (10,14) (11,12) (12,10), (12,10) (13,12) (14,14).
# S6: p0..p4 = (10,14) (11,12) (12,10) (13,12) (14,14).  S7, n = 2: row = [4,0,1] ++ map (+2) (makeTreeRow 2);
# makeTreeRow 2: internal [0], tree [0,1] -> [2,3].  Memory: p4 p0 p1 p2 p3.
CASES["H2_reversed_run"] = dict(
    pairs=_straight([(14, 14), (12, 10), (10, 14)]),
    strands=[
        [(14, 14), (10, 14), (12, 14)],
        [(14, 14), (10, 14), (11, 12), (12, 10), (13, 12)],
    ],
    box=(10, 10, 14, 14),
)

# ---- H3: a run of 17 Béziers is cut at 16 --------------------------------------------------------------------------------
# 18 points P_j = (3 + j, 5 + 2 (j mod 2)), j = 0..17: a zig-zag walking right, closed by one long side back to P_0.
# S1: b_j = (P_j, mid, P_{j+1}) for j = 0..16 all go right; b_17 = (P_17, mid, P_0) = (20,7) (11.5,6) (3,5) goes left.
# S3: acc grows to [b_0..b_16]; b_17 is not connectable: acc [b_17], strands [[b_0..b_16]].  Result [[b_17],[b_0..b_16]].
# S4: [b_17] stays; the run of 17 > 16 is cut: [b_0..b_15], [b_16].  Three strands.
# [b_17] reversed (20 > 3): (3,5) (11.5,6) (20,7): memory [(20,7), (3,5), (11.5,6)].
# [b_0..b_15]: S6 gives 33 points p_k: p_2j = P_j, p_2j+1 = mid(P_j, P_j+1).  S7, n = 16: row = [32,0,1] ++ map (+2)
#   (makeTreeRow 30): internal [0..14]; buildITree is the perfect tree (perfectTreePartition 15 = 7, 7 -> 3, 3 -> 1, 1 -> 0;
#   RT:44-68) whose breadth-first order is 7, 3,11, 1,5,9,13, 0,2,4,6,8,10,12,14; doubled and paired, plus 2:
_ROW16 = [32, 0, 1, 16, 17, 8, 9, 24, 25, 4, 5, 12, 13, 20, 21, 28, 29,
          2, 3, 6, 7, 10, 11, 14, 15, 18, 19, 22, 23, 26, 27, 30, 31]
# [b_16] = (19,5) (19.5,6) (20,7): memory [(20,7), (19,5), (19.5,6)].
_P = [(3 + j, 5 + 2 * (j % 2)) for j in range(18)]
_p = []
for _j in range(16):
    _p += [_P[_j], _mid(_P[_j], _P[_j + 1])]
_p.append(_P[16])
CASES["H3_run_of_17_cut_at_16"] = dict(
    pairs=_straight(_P),
    strands=[
        [(20, 7), (3, 5), (11.5, 6)],
        [_p[k] for k in _ROW16],
        [(20, 7), (19, 5), (19.5, 6)],
    ],
    box=(3, 5, 20, 7),
)


def pairs_array(case):
    return np.asarray([[on[0], on[1], off[0], off[1]] for on, off in case["pairs"]], np.float32)
