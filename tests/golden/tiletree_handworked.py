"""Hand-worked tile-tree fixtures: what `addShapeToTree` and `buildRasterJobs` must produce on six small inputs,
derived on paper from the Haskell, rule by rule.  NOTHING here is computed by a tile-tree algorithm — the expected
leaves are literals (ranges written out with `range`) and every step of the derivation cites the line of
/root/reference/src/Graphics/Gudni/Raster/TileTree.hs (TT), Raster/Job.hs (JOB) or OpenCL/CallKernels.hs (CK) it
applies.  tests/test_tiletree_handworked.py holds the oracle's tile tree (oracle/tiletree_oracle.cpp) to them on the
CPU and tests/test_gpu_tiletree_handworked.py the GPU binning (gudni_b200/csrc/binning.cu).

Notation.  A shape is (left, top, right, bottom, strands); shapes are numbered in insertion order.  A leaf is
(left, top, right, bottom, hDepth, vDepth, [shape numbers, NEWEST FIRST]) — `insertShapeTile` conses (TT:158), and
`addTileToRasterJob` appends that list as it stands (JOB:137-142).  Leaves are listed in `traverseTileTree` order:
top before bottom, left before right (TT:196-204).

The rules, once:
  R1  buildTileTree (TT:81-109): the tree covers the square of side 2^ceil(log2(max(w,h))) (TT:85-93); goV cuts at
      top + 2^(depth-1) and hands both halves to goH AT THE SAME depth (TT:96-100); goH cuts at left + 2^(depth-1) and
      hands both halves to goV at depth-1 (TT:104-108); at depth == tileDepth the node is a leaf with
      hDepth = vDepth = depth (TT:101,109).  So a canvas no larger than one tile is a single VLeaf.
  R2  insertShapeV on a VTree (TT:137-144): into `top` iff box.top < cut, into `bottom` iff box.bottom > cut — strict,
      tested independently, so a box may go into both, or (degenerate, top == bottom == cut) into neither.
      insertShapeH on an HTree (TT:118-127): `left` iff box.left < cut, `right` iff box.right > cut.
  R3  a leaf takes the shape if checkTileSpace holds or it is at the size floor (TT:128-132,145-149):
      checkTileSpace = shapeCount < mAXsHAPE - 1 = 126  AND  strandCount + shape.strands < maxStrandsPerTile
      (TT:167-172; both strict); floor: HLeaf when width <= 8, VLeaf when height <= 8 (mINtILEsIZE, Constants.hs:65).
  R4  otherwise the leaf splits and THEN the shape is inserted into the result (TT:134,151).  vSplit of a VLeaf
      (TT:184-190): cut = top + height div 2, children are HLeaves with vDepth-1 (hDepth unchanged), old shapes
      re-inserted OLDEST first (`foldl insertShapeV vTree $ reverse tileShapes`), through R2/R3 — so a child may split
      again while being refilled.  hSplit of an HLeaf (TT:175-181): cut = left + width div 2, children VLeaves with
      hDepth-1.
  R5  buildRasterJobs (CK:244-255) calls `accumulateRasterJobs threadsPerTile tilesPerCall` where the callee's
      parameters are `maxTilesPerJob threadsPerTile` (JOB:151-156): the two are swapped.  A job therefore closes
      after `threads_per_tile` tiles (JOB:163-171) and a tile's column allocation is the job's running sum of
      `max_tiles_per_call` (JOB:144-145,178).  The job list is `bsCurrentJob : bsJobs`, last created first (CK:255).
"""

CANONICAL = dict(max_tile_size=256, threads_per_tile=256, max_tiles_per_call=256, max_thresholds=1024,
                 max_strands_per_tile=1022, max_shapes=127)


def _down(hi, lo):
    """hi, hi-1, ..., lo"""
    return list(range(hi, lo - 1, -1))


CASES = {}

# ---- T1: four root tiles; the strict comparisons at a cut -----------------------------------------------------
# Canvas 512x512, tile 256.  R1: canvasDepth 9 > tileDepth 8: VTree cut 256 -> two HTrees cut 256 (same depth 9) ->
# four VLeaves at depth 8.  Traversal: (0,0) (256,0) (0,256) (256,256).
#   #0 (10,10,100,100): top 10 < 256 -> top, bottom 100 > 256 no; in the top HTree left 10 < 256 -> left, right
#      100 > 256 no.  Leaf A only.
#   #1 (200,200,300,300): top and bottom, left and right: all four leaves.
#   #2 (256,0,400,256) sits ON both cuts: top 0 < 256 -> top; bottom 256 > 256 is false -> not bottom; left 256 < 256
#      false -> not left; right 400 > 256 -> right.  Leaf B only.
#   #3 (256,10,256,20), zero width on the vertical cut: top only; left 256 < 256 false, right 256 > 256 false:
#      NEITHER branch — the shape is in no leaf (R2).
#   #4 (10,256,20,256), zero height on the horizontal cut: top 256 < 256 false, bottom 256 > 256 false: no leaf.
#   #5 (0,0,512,512): all four.
CASES["T1_quadrants_and_cuts"] = dict(
    spec=CANONICAL, canvas=(512, 512),
    shapes=[(10, 10, 100, 100, 4), (200, 200, 300, 300, 4), (256, 0, 400, 256, 4), (256, 10, 256, 20, 4),
            (10, 256, 20, 256, 4), (0, 0, 512, 512, 4)],
    leaves=[(0, 0, 256, 256, 8, 8, [5, 1, 0]),
            (256, 0, 512, 256, 8, 8, [5, 2, 1]),
            (0, 256, 256, 512, 8, 8, [5, 1]),
            (256, 256, 512, 512, 8, 8, [5, 1])],
)

# ---- T2: the 127th shape splits a tile, first across (vSplit), later a half along (hSplit) ---------------------
# Canvas 256x256: one VLeaf (0,0,256,256) depth (8,8) (R1).
#   #0..#62   (10,10,20,20)    63 shapes in the upper half
#   #63..#125 (10,200,20,210)  63 shapes in the lower half        -> count 126, strands 504: all taken by R3.
#   #126 (100,100,150,150): count 126 < 126 fails, height 256 > 8 -> vSplit (R4): cut 128, HLeaves
#      T (0,0,256,128) and B (0,128,256,256), depth (8,7).  Refill oldest first: #0..#62 top 10 < 128, bottom 20 > 128
#      no -> T; #63..#125 -> B.  Then #126: 100 < 128 -> T (count 63, room), 150 > 128 -> B.
#      T = [126, 62..0] (64), B = [126, 125..63] (64).
#   #127..#188 (200,10,210,20): 62 shapes, top only -> T, count 126.
#   #189 (120,50,136,60): top only; T is full, width 256 > 8 -> hSplit (R4): cut 128, VLeaves TL (0,0,128,128) and
#      TR (128,0,256,128), depth (7,7).  Refill T oldest first = #0..#62, #126, #127..#188:
#      #0..#62 left 10 < 128 -> TL; #126 left 100 < 128 -> TL and right 150 > 128 -> TR; #127..#188 left 200 -> TR only.
#      Then #189: 120 < 128 -> TL, 136 > 128 -> TR.
#      TL = [189, 126, 62..0] (65), TR = [189, 188..127, 126] (64).  B untouched.
CASES["T2_shape_count_vsplit_then_hsplit"] = dict(
    spec=CANONICAL, canvas=(256, 256),
    shapes=[(10, 10, 20, 20, 4)] * 63 + [(10, 200, 20, 210, 4)] * 63 + [(100, 100, 150, 150, 4)] +
           [(200, 10, 210, 20, 4)] * 62 + [(120, 50, 136, 60, 4)],
    leaves=[(0, 0, 128, 128, 7, 7, [189, 126] + _down(62, 0)),
            (128, 0, 256, 128, 7, 7, [189] + _down(188, 127) + [126]),
            (0, 128, 256, 256, 8, 7, [126] + _down(125, 63))],
)

# ---- T3: the strand cap drives the split; a sum EQUAL to the cap already splits ---------------------------------
# maxStrandsPerTile = 12, every shape 4 strands.  Canvas 256x256: one VLeaf.
#   #0 (10,10,20,20): 0+4 < 12.  #1 (10,200,20,210): 4+4 < 12.  Tile holds 8 strands.
#   #2 (100,100,150,150): 8+4 = 12 < 12 fails (R3, strict) -> vSplit at 128: T (0,0,256,128), B (0,128,256,256), (8,7).
#      Refill: #0 -> T, #1 -> B; #2 -> both (4+4 < 12).  T = [2,0] 8 strands, B = [2,1] 8 strands.
#   #3 (30,30,40,40): top only.  T: 8+4 = 12 fails -> hSplit at 128: TL (0,0,128,128), TR (128,0,256,128), (7,7).
#      Refill: #0 -> TL; #2 left 100 < 128 -> TL, right 150 > 128 -> TR.  TL = [2,0] 8 strands, TR = [2].
#      Then #3 -> left only: TL is a VLeaf with 8+4 = 12: fails, height 128 > 8 -> vSplit at 64:
#      TLt (0,0,128,64), TLb (0,64,128,128), depth (7,6).  Refill: #0 (10..20) top only -> TLt; #2 top 100 < 64 no,
#      bottom 150 > 64 -> TLb.  Then #3 (30..40): top only -> TLt (4+4 < 12).
#      TLt = [3,0], TLb = [2], TR = [2], B = [2,1].
CASES["T3_strand_cap"] = dict(
    spec=dict(CANONICAL, max_strands_per_tile=12), canvas=(256, 256),
    shapes=[(10, 10, 20, 20, 4), (10, 200, 20, 210, 4), (100, 100, 150, 150, 4), (30, 30, 40, 40, 4)],
    leaves=[(0, 0, 128, 64, 7, 6, [3, 0]),
            (0, 64, 128, 128, 7, 6, [2]),
            (128, 0, 256, 128, 7, 7, [2]),
            (0, 128, 256, 256, 8, 7, [2, 1])],
)

# ---- T4: the 8-pixel floor keeps more than 126 shapes ------------------------------------------------------------
# Canvas 256x256, 130 copies of the box (0,0,256,256).  #0..#125 fill the root leaf.  #126 splits it (vSplit at 128);
# every old shape spans every cut (0 < cut < 256), so each refill puts all 126 into both children, which are full
# again, and #126 arriving in a child splits that in turn: HLeaf (256 wide) hSplit, VLeaf (128 high) vSplit, ... down
# to 16x8 HLeaves, whose hSplit gives 8x8 VLeaves: height 8 <= 8, the floor (R3) — they take #126 as their 127th
# shape, and #127..#129 after it.  Result: 32 x 32 leaves of 8x8, depth (3,3) (8 = 2^3; each axis lost 5 levels),
# every one holding [129..0].  Order: at every level top-left, top-right, bottom-left, bottom-right, i.e. leaf k sits
# at the position whose y bits are the odd bits of k and whose x bits are the even bits (5 bits each).
def _deinterleave(k):
    x = y = 0
    for b in range(5):
        x |= ((k >> (2 * b)) & 1) << b
        y |= ((k >> (2 * b + 1)) & 1) << b
    return x, y


CASES["T4_floor_keeps_130"] = dict(
    spec=CANONICAL, canvas=(256, 256),
    shapes=[(0, 0, 256, 256, 4)] * 130,
    leaves=[(8 * x, 8 * y, 8 * x + 8, 8 * y + 8, 3, 3, _down(129, 0)) for x, y in map(_deinterleave, range(1024))],
)

# ---- T5: a split child that has to split again while it is being refilled ----------------------------------------
# maxStrandsPerTile = 12, canvas 256x256.  #0, #1 both (10,10,20,20) (8 strands), #2 (10,200,20,210): 8+4 = 12 fails
# -> vSplit at 128, refill #0 -> T (4), #1 -> T (8); #2 -> B.  T = [1,0], B = [2].
#   #3 (12,12,18,18): T: 8+4 fails -> hSplit at 128: TL, TR (7,7); refill #0, #1 -> TL (8 strands); #3 -> TL: fails,
#   vSplit at 64: TLt (0,0,128,64), TLb (7,6); refill #0, #1 -> TLt (8); #3 -> TLt: HLeaf, 8+4 fails, width 128 > 8 ->
#   hSplit at 64: (0,0,64,64), (64,0,128,64), depth (6,6); refill #0, #1 -> left (8); #3 -> left: VLeaf fails, height
#   64 > 8 -> vSplit at 32 -> HLeaves (0,0,64,32), (0,32,64,64), (6,5); #0,#1 -> top (8); #3 -> top: hSplit at 32 ->
#   (0,0,32,32), (32,0,64,32) (5,5); #0,#1 left; #3 left: vSplit at 16 -> (0,0,32,16) (5,4), (0,16,32,32): here
#   #0, #1 (10..20) span the cut: top 10 < 16 and bottom 20 > 16 -> both halves get both (8 strands each); #3 (12..18)
#   spans it too: top half 8+4 fails -> HLeaf (0,0,32,16) hSplit at 16 -> (0,0,16,16) (4,4), (16,0,32,16); #0, #1:
#   left 10 < 16 and right 20 > 16 -> both (8 each); #3 left 12 < 16 -> left VLeaf (0,0,16,16): fails, height 16 > 8 ->
#   vSplit at 8 -> (0,0,16,8), (0,8,16,16) (4,3); #0, #1: top 10 < 8 no, bottom -> lower only (8); #3: top 12 < 8 no ->
#   lower: 8+4 fails, width 16 > 8 -> hSplit at 8 -> (0,8,8,16), (8,8,16,16) (3,3); #0,#1: left 10 < 8 no; right -> R (8);
#   #3: left 12 < 8 no; right: VLeaf (8,8,16,16) 8+4 fails but height 8 <= 8: floor, taken: [3,1,0].
#   Back up: #3 right 18 > 16 -> VLeaf (16,0,32,16): [1,0] 8+4 fails, height 16 > 8 -> vSplit at 8: (16,0,32,8),
#   (16,8,32,16) (4,3): #0, #1 top 10 < 8 no -> lower (8); #3 -> lower: fails, width 16 > 8 -> hSplit at 24:
#   (16,8,24,16), (24,8,32,16) (3,3): #0, #1: left 10 < 24 -> L; right 20 > 24 no. #3: left 12 < 24 -> L: floor (height
#   8): [3,1,0]; right 18 > 24 no.
#   And the bottom half of the cut at 16, HLeaf (0,16,32,32) (5,4) holding [1,0]: #3 bottom 18 > 16 -> 8+4 fails, width
#   32 > 8 -> hSplit at 16: (0,16,16,32), (16,16,32,32) (4,4): #0,#1 both sides (10 < 16, 20 > 16); #3 left 12 < 16 ->
#   VLeaf (0,16,16,32): fails, height 16 > 8 -> vSplit at 24: (0,16,16,24), (0,24,16,32) (4,3): #0,#1 top 10 < 24 -> upper;
#   bottom 20 > 24 no.  #3: top 12 < 24 -> upper HLeaf (0,16,16,24): fails, width 16 > 8 -> hSplit at 8: (0,16,8,24),
#   (8,16,16,24) (3,3): #0,#1 left 10 < 8 no, right -> R; #3 -> R: floor: [3,1,0]; bottom 18 > 24 no.
#   #3 right 18 > 16 -> VLeaf (16,16,32,32) [1,0]: fails -> vSplit at 24: (16,16,32,24), (16,24,32,32) (4,3): #0,#1 upper;
#   #3 upper: fails -> hSplit at 24: (16,16,24,24), (24,16,32,24) (3,3): #0,#1 left (10 < 24), not right (20 > 24 no);
#   #3 left: floor [3,1,0]; right 18 > 24 no.
# Leaves in traversal order (V128: top first; H128: left first; V64; H64; V32; H32; V16; ...):
CASES["T5_cascade_to_the_floor"] = dict(
    spec=dict(CANONICAL, max_strands_per_tile=12), canvas=(256, 256),
    shapes=[(10, 10, 20, 20, 4), (10, 10, 20, 20, 4), (10, 200, 20, 210, 4), (12, 12, 18, 18, 4)],
    leaves=[
        # inside (0,0,32,16) = H16{ V8{ (0,0,16,8), H8{(0,8,8,16),(8,8,16,16)} }, V8{ (16,0,32,8), H24{..} } }
        (0, 0, 16, 8, 4, 3, []),
        (0, 8, 8, 16, 3, 3, []),
        (8, 8, 16, 16, 3, 3, [3, 1, 0]),
        (16, 0, 32, 8, 4, 3, []),
        (16, 8, 24, 16, 3, 3, [3, 1, 0]),
        (24, 8, 32, 16, 3, 3, []),
        # inside (0,16,32,32) = H16{ V24{ H8{(0,16,8,24),(8,16,16,24)}, (0,24,16,32) }, V24{ H24{..}, (16,24,32,32) } }
        (0, 16, 8, 24, 3, 3, []),
        (8, 16, 16, 24, 3, 3, [3, 1, 0]),
        (0, 24, 16, 32, 4, 3, []),
        (16, 16, 24, 24, 3, 3, [3, 1, 0]),
        (24, 16, 32, 24, 3, 3, []),
        (16, 24, 32, 32, 4, 3, []),
        # the empty siblings on the way up
        (32, 0, 64, 32, 5, 5, []),
        (0, 32, 64, 64, 6, 5, []),
        (64, 0, 128, 64, 6, 6, []),
        (0, 64, 128, 128, 7, 6, []),
        (128, 0, 256, 128, 7, 7, []),
        (0, 128, 256, 256, 8, 7, [2]),
    ],
)

# ---- T6: job packing with the swapped arguments (R5) ---------------------------------------------------------------
# Spec: tile 64, threads_per_tile 64, max_tiles_per_call 128.  Canvas 1024x512: the tree covers 1024x1024 (R1):
# canvasDepth 10, tileDepth 6: 16 x 16 = 256 leaves, depth (6,6), in the order of T4 (y bits odd, x bits even, 4 bits
# each).  One shape, (60,60,70,70): crosses x = 64 and y = 64, so it is in leaves (0,0) (1,0) (0,1) (1,1) = k 0,1,2,3.
# Jobs (R5): a job closes after threads_per_tile = 64 tiles -> 4 jobs of 64 leaves; within a job the i-th tile has
# column_allocation = i * max_tiles_per_call = 128 i, the job's total is 64 * 128 = 8192; the list comes back last
# created first: jobs[0] holds leaves 192..255, jobs[3] leaves 0..63 and the only four shape references.
def _deinterleave4(k):
    x = y = 0
    for b in range(4):
        x |= ((k >> (2 * b)) & 1) << b
        y |= ((k >> (2 * b + 1)) & 1) << b
    return x, y


CASES["T6_job_packing_swapped_arguments"] = dict(
    spec=dict(max_tile_size=64, threads_per_tile=64, max_tiles_per_call=128, max_thresholds=256,
              max_strands_per_tile=254, max_shapes=127),
    canvas=(1024, 512),
    shapes=[(60, 60, 70, 70, 4)],
    leaves=[(64 * x, 64 * y, 64 * x + 64, 64 * y + 64, 6, 6, [0] if k < 4 else [])
            for k, (x, y) in enumerate(map(_deinterleave4, range(256)))],
    jobs=dict(count=4, tiles_per_job=64, column_step=128, columns_per_job=8192,
              first_leaf_of_job=[192, 128, 64, 0]),
)
