// raster_kernels.cu — the raster kernels and their launch wrappers.  Device arithmetic is in raster_device.cuh
// (per-column-thread functions shared with the replay), raster_warp.cuh (generate) and raster_split.cuh (render).
#include "raster_kernels.cuh"

#include <algorithm>

#include "raster_sort.cuh"
#include "raster_split.cuh"

using namespace gudni_dev;

#ifndef GUDNI_SLICE_WARPS
#define GUDNI_SLICE_WARPS 4
#endif
#ifndef GUDNI_COLOR_WARPS
#define GUDNI_COLOR_WARPS 4
#endif
#ifndef GUDNI_MAX_LANE_SHIFT
#define GUDNI_MAX_LANE_SHIFT 0       // whole-warp units.  Narrower ones measured slower on a 2,048-row strip of S5 (3.81 ms with
#endif                               // 32 lanes, 3.96 with 16, 4.87 with 8): a unit's time is the latency of its longest column's
                                     // chain of records, not the number of lanes that diverge, so narrower units only add waves
constexpr int kSliceWarpsPerCta = GUDNI_SLICE_WARPS;
constexpr int kColorWarpsPerCta = GUDNI_COLOR_WARPS;

// A frame (or a launch of up to kLaunchTiles tiles of it) goes through these kernels on one stream (on two, as two batches
// of tiles, when its pixels leave the GPU: rasterTiles); every one sizes
// its grid to what the chip holds resident and pulls its work from a global counter, expensive tiles first:
//   raster_generate_kernel    generateThresholds (K.cl:2030-2082), one CTA per tile: the column-threads' threshold
//                             queues packed into HBM (20 B per threshold)
//   raster_sort_kernel        sortThresholds (K.cl:2084-2115): a warp rank-sorts the 32 queues of a unit in place
//   raster_slice_kernel       the state machine of renderThresholds (K.cl:2117-2167) without colours: one
//                             section stream per column-thread
//   raster_resolve_kernel     the streams' shape stacks numbered (deduplicated per warp)
//   raster_composite_kernel   every numbered stack composited once (determineColor, K.cl:1447-1513)
//   raster_accumulate_kernel  colour * area per section, pixels stored
//   raster_picture_kernel     tiles with picture substances: one pass, colours per pixel
// and raster_spill_kernel replays, lane-privately against an HBM queue, the column-threads that did not fit the
// on-chip structures.  Round 1 ran renderThresholds as one kernel; split this way each phase has the occupancy
// and the lane utilisation it can reach (profiles/README.md).
// generateThresholds (K.cl:2030-2082), one CTA per tile: blockDim.x = threadsPerTile, thread =
// column-thread.  Dynamic shared memory: the tile's staged strand headers, then one queue window per warp.
__global__ void __launch_bounds__(1024) raster_generate_kernel(const FrameParams P, int tileBase, int nTiles) {
#ifdef GUDNI_HOST_EMULATION
    unsigned char* smemRaw = cuemu::dynamicShared;
#else
    extern __shared__ __align__(16) unsigned char smemRaw[];
#endif
    TileStage& S = *reinterpret_cast<TileStage*>(smemRaw);
    SearchSummaries R;
    {
        unsigned char* p = smemRaw + ((sizeof(TileStage) + 15) & ~(size_t)15);
        R.yRange = reinterpret_cast<float2*>(p);
        R.flags = reinterpret_cast<uint32_t*>(p + (size_t)kGenItems * 8);
    }
    GenScratch* scratch = reinterpret_cast<GenScratch*>(smemRaw + ((sizeof(TileStage) + 15) & ~(size_t)15) + searchSummariesBytes());
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    QueueCold<kQueueCap - kGenQueueHot> cold;
    GenQueue q;
    q.limit = min(kQueueCap, P.maxThresholds);
    q.cold = &cold;
    q.thrHot = scratch[warp].qThr + lane;
    q.hdrHot = scratch[warp].qHdr + lane;
    const uint32_t denseCap = (uint32_t)min(P.maxShape, kWarpTableCap);
    unsigned int* workCounter = P.work + kWorkGenerate;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { S.tileSlot = (int)atomicAdd(workCounter, 1u); S.anyPicture = 0; }
        __syncthreads();
        const int tileSlot = S.tileSlot;
        if (tileSlot >= nTiles) break;
        const int tileIndex = (int)P.tileOrder[tileBase + tileSlot * P.batchStride + P.batchIndex];
        const gudni_tile tile = P.tiles[tileIndex];
        const int column = (int)threadIdx.x;   // the reference's thread number inside the tile
        const ThreadGeom g = threadGeom(P, tile, column);
        int generated = -1;
        int failed = 0;
        bool exhausted = false;
        if (tile.shape_count <= denseCap) {
            GenThread t;
            t.stack = ShapeStack{0ull, 0ull};
            t.added = ShapeStack{0ull, 0ull};
            t.failed = false;
            q.init();
            generateTileThresholds(P, S, R, q, t, tile, g);
            failed = packWarp(P, q, g, t.stack, t.bits(), t.failed, tileIndex, column, generated, exhausted, S.anyPicture != 0);
        } else {
            // a tile that stopped splitting at the 8-pixel floor with more shapes than stack bits:
            // its threads take the lane-private replay path (bit -> shape table, HBM queue)
            ThreadRec rec{};
            rec.count = kRecInactive;
            P.threadRecs[((size_t)tileIndex << P.computeDepth) + (size_t)column] = rec;
            failed = g.active ? 1 : 0;
        }
        // statistics: thresholds of lanes that completed here (replayed lanes are counted by the replay)
        unsigned int mine = (!failed && generated > 0) ? (unsigned)generated : 0u;
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(full, mine, d);
        if (lane == 0 && mine) atomicAdd(&P.counters[kCntThresholds], (unsigned long long)mine);
        if (exhausted) atomicAdd(&P.counters[kCntExhausted], 1ull);
        if (failed) registerSpill(P, tileIndex, column);
    }
}

// sortThresholds (K.cl:2084-2115): the 32 queues of a unit rank-sorted by one warp (raster_sort.cuh).
#ifndef GUDNI_SORT_WARPS
#define GUDNI_SORT_WARPS 2
#endif
constexpr int kSortWarpsPerCta = GUDNI_SORT_WARPS;
__global__ void __launch_bounds__(kSortWarpsPerCta * 32) raster_sort_kernel(const FrameParams P, int tileBase, int nTiles);

// The sweep's state machine alone (raster_split.cuh): sorted thresholds in, section streams out.
#ifndef GUDNI_SLICE_MIN_CTAS
#define GUDNI_SLICE_MIN_CTAS 8      // 64 registers: 32 warps per SM hide the divergent kernel's latencies (96 registers / 20 warps: +1.1 ms on S4)
#endif
// Persistent-warp loop over the units of a launch, most expensive tiles first.  A unit is 32 >> laneShift
// neighbouring column-threads of a tile (the reference's work-group sliced by warps, `Work2D numTiles
// threadsPerTile`, OpenCL/CallKernels.hs:141-142).  laneShift > 0 (narrower units for launches with few tiles) is
// kept as a build option; it measured slower, see GUDNI_MAX_LANE_SHIFT.
// body(tileIndex, tile, rec, column) is called by the whole warp for every unit of a dense tile; rec is the
// lane's thread record, or null for a lane beyond the unit's width.
template <class F>
__device__ __forceinline__ void forEachUnit(const FrameParams& P, int tileBase, int nTiles, int counterSlot, F body) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int unitShift = P.computeDepth - 5 + P.laneShift;            // units per tile
    const int lanesPerUnit = 32 >> P.laneShift;
    const unsigned totalUnits = (unsigned)nTiles << unitShift;
    const uint32_t denseCap = (uint32_t)min(P.maxShape, kWarpTableCap);
    unsigned int* workCounter = P.work + counterSlot;
    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd(workCounter, 1u);
        unit = __shfl_sync(full, unit, 0);
        if (unit >= totalUnits) break;
        const int tileIndex = (int)P.tileOrder[tileBase + (int)(unit >> unitShift) * P.batchStride + P.batchIndex];
        const unsigned unitInTile = unit & ((1u << unitShift) - 1u);
        const gudni_tile tile = P.tiles[tileIndex];
        if (tile.shape_count > denseCap) continue;   // replayed lane-privately
        const int column = (int)(unitInTile * (unsigned)lanesPerUnit) + min(lane, lanesPerUnit - 1);
        ThreadRec* rec = lane < lanesPerUnit ? P.threadRecs + (((size_t)tileIndex << P.computeDepth) + (size_t)column) : nullptr;
        body(tileIndex, tile, rec, column);
    }
}

__global__ void __launch_bounds__(kSortWarpsPerCta * 32) raster_sort_kernel(const FrameParams P, int tileBase, int nTiles) {
    __shared__ SortScratch scratch[kSortWarpsPerCta];
    SortScratch& W = scratch[threadIdx.x >> 5];
    BulkStage B;
    bulkInit(W, B);
    forEachUnit(P, tileBase, nTiles, kWorkSort, [&](int, const gudni_tile&, ThreadRec* rec, int) {
        sortWarp(P, W, B, rec);
        __syncwarp();
    });
}

__global__ void __launch_bounds__(kSliceWarpsPerCta * 32, GUDNI_SLICE_MIN_CTAS) raster_slice_kernel(const FrameParams P, int tileBase, int nTiles) {
    __shared__ SliceScratch scratch[kSliceWarpsPerCta];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SliceScratch& W = scratch[warp];
    if (lane == 0) { W.slabNext = 0u; W.slabEnd = 0u; }
    __syncwarp();
    ActiveRun q;
    q.thr = W.aThr + lane;
    q.hdr = W.aHdr + lane;
    forEachUnit(P, tileBase, nTiles, kWorkSlice, [&](int tileIndex, const gudni_tile& tile, ThreadRec* rec, int column) {
        const unsigned int count = rec ? rec->count : 0u;
        bool exhausted = false, wide = false;
        const int failed = sliceWarp<kActiveCap>(P, W, q, tile, rec, column, exhausted, wide);
        if (failed) {
            // its thresholds were counted by the generate kernel; the replay counts them again
            atomicAdd(&P.counters[kCntThresholds], 0ull - (unsigned long long)count);
            if (exhausted) atomicAdd(&P.counters[kCntExhausted], 1ull);
            registerSpill(P, tileIndex, column);
        }
        // a unit with a thread for the wide pass is listed once (the flags are in the thread records)
        const unsigned wideLanes = __ballot_sync(0xffffffffu, wide);
        if (wideLanes && lane == __ffs((int)wideLanes) - 1) {
            const unsigned int unitInTile = (unsigned int)column >> (5 - P.laneShift);
            P.wideList[atomicAdd(P.work + kWorkWideCount, 1u)] = ((unsigned int)tileIndex << 8) | unitInTile;
        }
    });
}

// The threads the slice kernel flagged kRecWide, sliced again with room for a run of kActiveCapWide thresholds per lane.
// One warp per CTA (the scratch is most of the 48 KB of static shared memory); launched after every slice pass, it finds
// its list empty on all but pathological frames.
__global__ void __launch_bounds__(32) raster_slice_wide_kernel(const FrameParams P) {
    __shared__ SliceScratchT<kActiveCapWide> W;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned int listed = P.work[kWorkWideCount];
    if (listed == 0u) return;
    if (lane == 0) { W.slabNext = 0u; W.slabEnd = 0u; }
    __syncwarp();
    ActiveRun q;
    q.thr = W.aThr + lane;
    q.hdr = W.aHdr + lane;
    const int lanesPerUnit = 32 >> P.laneShift;
    for (;;) {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(P.work + kWorkWideCursor, 1u);
        i = __shfl_sync(full, i, 0);
        if (i >= listed) break;
        const unsigned int entry = P.wideList[i];
        const int tileIndex = (int)(entry >> 8);
        const gudni_tile tile = P.tiles[tileIndex];
        const int column = (int)((entry & 0xFFu) * (unsigned)lanesPerUnit) + min(lane, lanesPerUnit - 1);
        ThreadRec* rec = lane < lanesPerUnit ? P.threadRecs + (((size_t)tileIndex << P.computeDepth) + (size_t)column) : nullptr;
        if (rec && !(rec->pad1 & kRecWide)) rec = nullptr;   // the unit's other threads are done
        if (rec) rec->pad1 &= ~kRecWide;
        const unsigned int count = rec ? rec->count : 0u;
        bool exhausted = false, wide = false;
        const int failed = sliceWarp<kActiveCapWide>(P, W, q, tile, rec, column, exhausted, wide);
        if (failed) {
            atomicAdd(&P.counters[kCntThresholds], 0ull - (unsigned long long)count);
            if (exhausted) atomicAdd(&P.counters[kCntExhausted], 1ull);
            registerSpill(P, tileIndex, column);
        }
        __syncwarp();
    }
}

#ifndef GUDNI_RESOLVE_WARPS
#define GUDNI_RESOLVE_WARPS 4
#endif
#ifndef GUDNI_COMPOSITE_WARPS
#define GUDNI_COMPOSITE_WARPS 4
#endif
#ifndef GUDNI_ACCUMULATE_WARPS
#define GUDNI_ACCUMULATE_WARPS 4
#endif
constexpr int kResolveWarpsPerCta = GUDNI_RESOLVE_WARPS;
constexpr int kCompositeWarpsPerCta = GUDNI_COMPOSITE_WARPS;
constexpr int kAccumulateWarpsPerCta = GUDNI_ACCUMULATE_WARPS;

// Section streams -> numbered shape stacks (raster_split.cuh).
__global__ void __launch_bounds__(kResolveWarpsPerCta * 32) raster_resolve_kernel(const FrameParams P, int tileBase, int nTiles) {
    __shared__ ResolveScratch scratch[kResolveWarpsPerCta];
    ResolveScratch& W = scratch[threadIdx.x >> 5];
    RefSlab slab{kRefNone, 0u, -1};
    // (A warp taking the slabs of a tile in groups of 2 / 4 / 8, so that its stack cache serves what lies above and below a
    // slab border, measured slower on S4: 10.69 / 11.03 / 11.37 against 10.61 ms — S4's outlines cross every pixel about
    // twice, a slab's stacks are its own.)
    forEachUnit(P, tileBase, nTiles, kWorkResolve, [&](int tileIndex, const gudni_tile& tile, ThreadRec* rec, int column) {
        if (unitHasPictures(rec)) return;   // raster_picture_kernel
        const unsigned int count = rec ? rec->count : 0u;
        if (resolveWarp(P, W, slab, tileIndex, rec)) {
            atomicAdd(&P.counters[kCntThresholds], 0ull - (unsigned long long)count);
            atomicAdd(&P.counters[kCntExhausted], 1ull);
            registerSpill(P, tileIndex, column);
        }
    });
    closeSlab(P, slab);
}

// Every numbered stack composited once: a warp per slab of numbers.
__global__ void __launch_bounds__(kCompositeWarpsPerCta * 32) raster_composite_kernel(const FrameParams P) {
    __shared__ TileTable tables[kCompositeWarpsPerCta];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    TileTable& T = tables[threadIdx.x >> 5];
    // the slabs the batch's resolve pass filled: its region of the table, from the start (the table is reused launch by launch)
    const unsigned int drawn = P.work[kWorkRefSlabs];
    const unsigned int nSlabs = min(drawn, P.refCapSlabs);
    // what the whole table would have to hold for this launch: every batch as much as the hungriest one
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicMax(&P.counters[kCntRefSlabs], (unsigned long long)drawn * (unsigned long long)P.batchCount);
    unsigned int* workCounter = P.work + kWorkComposite;
    int tableTile = -1;
    int mode = 0;
    for (;;) {
        unsigned int s = 0;
        if (lane == 0) s = atomicAdd(workCounter, 1u);
        s = __shfl_sync(full, s, 0);
        if (s >= nSlabs) break;
        compositeSlab(P, T, tableTile, mode, P.refSlabBase + s);
    }
}

// The resolved streams with the colours at hand: accumulation and pixel stores.
__global__ void __launch_bounds__(kAccumulateWarpsPerCta * 32) raster_accumulate_kernel(const FrameParams P, int tileBase, int nTiles) {
    __shared__ AccumScratch scratch[kAccumulateWarpsPerCta];
    AccumScratch& W = scratch[threadIdx.x >> 5];
    forEachUnit(P, tileBase, nTiles, kWorkAccumulate, [&](int, const gudni_tile& tile, ThreadRec* rec, int column) {
        if (unitHasPictures(rec)) return;
        accumulateWarp(P, W, tile, rec, column);
        __syncwarp();
    });
}

// Tiles with picture substances: one pass, every lane composites its own sections (the colour depends on the pixel).
__global__ void __launch_bounds__(kColorWarpsPerCta * 32) raster_picture_kernel(const FrameParams P, int tileBase, int nTiles) {
    __shared__ TileTable tables[kColorWarpsPerCta];
    TileTable& T = tables[threadIdx.x >> 5];
    forEachUnit(P, tileBase, nTiles, kWorkPicture, [&](int, const gudni_tile& tile, ThreadRec* rec, int column) {
        if (!unitHasPictures(rec)) return;
        __syncwarp();
        bool anyPicture, anyWild;
        buildTileTable(P, T, tile, anyPicture, anyWild);
        __syncwarp();
        if (anyPicture) pictureWarp(P, T, tile, rec, column);
    });
}

// Replay of spilled column-threads.  Persistent: each thread owns one HBM queue slot and walks the
// spill list with a grid stride, so the scratch footprint is fixed (slots x MAXTHRESHOLDS x 20 B)
// regardless of how many threads spilled.
__global__ void __launch_bounds__(128) raster_spill_kernel(const FrameParams P, float4* thr, uint32_t* hdr, int slots) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned long long n = P.counters[kCntSpilled];
    if (n > (unsigned long long)P.spillCapacity) n = (unsigned long long)P.spillCapacity;
    // (warp-uniform trip count: the lanes of a warp colour their sections together, rasterThread)
    for (unsigned long long first = (unsigned long long)(slot - lane); first < n; first += (unsigned long long)slots) {
        const unsigned long long i = first + (unsigned long long)lane;
        const bool active = i < n;
        const unsigned long long packed = active ? P.spillList[i] : 0ull;
        const int tileIndex = (int)(packed >> 32), column = (int)(packed & 0xFFFFFFFFull);
        const gudni_tile tile = P.tiles[tileIndex];
        const ThreadGeom g = threadGeom(P, tile, column);
        HbmQueue q;
        q.thr = thr + slot;
        q.hdr = hdr + slot;
        q.stride = (size_t)slots;
        q.cap = P.maxThresholds;
        // the first kernel does not count the thresholds of a thread it hands over
        int generated;
        const bool ok = rasterThread(P, g, q, P.tileThreadBase[tileIndex] + column, generated, active);
        if (active && generated > 0) atomicAdd(&P.counters[kCntThresholds], (unsigned long long)generated);
        if (active && !ok) atomicAdd(&P.counters[kCntOverflow], 1ull);
    }
}

// Counting sort of the launch's tiles by shape count, descending (one CTA; 256 bins, counts >= 255
// share the first bin).  Longest-processing-time-first order for the persistent kernels.
// It is also where a frame whose geometry holds a point at infinity is defused (kCntNonFinite, set by
// strand_bounds_kernel earlier on the stream): the tiles of the launch lose their shape lists, so the
// raster kernels that follow never walk a strand and only paint background; frame_end reports the error.
__global__ void __launch_bounds__(1024) tile_order_kernel(gudni_tile* __restrict__ tiles, int tileBase, int nTiles,
                                                          uint32_t* __restrict__ order, unsigned long long* __restrict__ counters) {
    __shared__ unsigned int bins[256];
    __shared__ unsigned int starts[256];
    if (counters[kCntNonFinite])
        for (int i = threadIdx.x; i < nTiles; i += blockDim.x) tiles[tileBase + i].shape_count = 0u;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bins[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < nTiles; i += blockDim.x)
        atomicAdd(&bins[255u - min(tiles[tileBase + i].shape_count, 255u)], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        for (int b = 0; b < 256; b++) { starts[b] = acc; acc += bins[b]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nTiles; i += blockDim.x) {
        const unsigned int b = 255u - min(tiles[tileBase + i].shape_count, 255u);
        order[tileBase + atomicAdd(&starts[b], 1u)] = (uint32_t)(tileBase + i);
    }
}

// (min y, max y) over all points of every strand of `count` shape records (`stride` bytes apart, each
// starting with {u64 tag, u32 geo_start, u32 num_strands}); strand layout K.cl:1365-1376.
__global__ void strand_bounds_kernel(const uint8_t* __restrict__ geometry, size_t geometryBytes, const uint8_t* __restrict__ records,
                                     int stride, int count, float2* __restrict__ bounds, unsigned long long* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    bool infinite = false, outside = false;
    const uint4 rec = __ldg(reinterpret_cast<const uint4*>(records + (size_t)i * stride));
    // The raster kernels trust geo_start, the strand count and the size words: this is the one place every strand
    // header is read before them, so a record that points outside the heap is caught here (and the frame defused
    // like one with an infinite coordinate: tile_order_kernel) instead of being walked.
    size_t at = 16ull * rec.z;
    for (uint32_t s = 0; s < rec.w; s++) {
        if (at + 32 > geometryBytes) { outside = true; break; }
        const uint8_t* strand = geometry + at;
        const uint32_t size = __ldg(reinterpret_cast<const uint32_t*>(strand)) & 0xFFFFu;   // in 8-byte units
        if (size < 4u || (size & 1u) || at + 8ull * size > geometryBytes) { outside = true; break; }   // 16-byte records: K.cl:1365-1376
        float lo = FLT_MAX, hi = -FLT_MAX;
        for (uint32_t k = 1; k < size; k++) {
            const float2 p = __ldg(reinterpret_cast<const float2*>(strand + 8 * k));
            lo = fminf(lo, p.y);
            hi = fmaxf(hi, p.y);
            infinite |= fabsf(p.x) == INFINITY || fabsf(p.y) == INFINITY;   // NaN terminates (and compares false)
#ifdef GUDNI_SEARCH_X
            // the x of every tree node (records 2.. of the strand), in the slot of this array that belongs to the node's
            // record: an x-only copy of the tree for the searches, which compare nothing else (strandSearchX)
            if (k >= 4 && !(k & 1)) bounds[(at >> 4) + (k >> 1)].x = p.x;
#endif
        }
        bounds[at >> 4] = make_float2(lo, hi);
        at += 8ull * size;
    }
    if (infinite) atomicOr(&counters[kCntNonFinite], 1ull);
    if (outside) atomicOr(&counters[kCntNonFinite], 2ull);
}

// div3 against the compiler's IEEE division on pseudo-random operands (and the edge cases around
// the fast-path window); counts bit mismatches.
__global__ void selftest_div3_kernel(unsigned long long n, unsigned long long seed, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + i * 0x9E3779B97F4A7C15ull;
        float v[4];
        for (int k = 0; k < 4; k++) {
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const unsigned mode = (unsigned)(z >> 60);
            float f = (float)((z >> 20) & 0xFFFFFFull) * (1.0f / 16777216.0f);   // [0,1)
            if (mode == 0) f = __uint_as_float((unsigned)(z >> 8) & 0x7FFFFFFFu);    // any non-negative bit pattern
            else if (mode == 1) f = f * 0x1p-58f;                                    // around the window's lower edge
            else if (mode == 2) f = 0.0f;
            else if (mode == 3) f = 1.0f;
            v[k] = f;
        }
        const float d = v[3];
        if (!(d > 0.0f) || isnan(v[0]) || isnan(v[1]) || isnan(v[2]) || isnan(d)) continue;
        float qx, qy, qz;
        div3<true>(v[0], v[1], v[2], d, qx, qy, qz);
        const float rx = v[0] / d, ry = v[1] / d, rz = v[2] / d;
        bad += (__float_as_uint(qx) != __float_as_uint(rx)) + (__float_as_uint(qy) != __float_as_uint(ry)) +
               (__float_as_uint(qz) != __float_as_uint(rz));
    }
    if (bad) atomicAdd(mismatches, bad);
}

#ifndef GUDNI_HOST_EMULATION   // the emulator drives the kernels itself (tests/native/raster_emu.cpp)
namespace gudni_launch {
int strandBounds(gudni_ctx* ctx, const void* geometry, const void* records, int stride, int count, float2* bounds) {
    if (count <= 0) return GUDNI_OK;
    unsigned long long* counters = ctx->counters.as<unsigned long long>();
    strand_bounds_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(geometry), ctx->geometryBytes,
                                                                     static_cast<const uint8_t*>(records), stride, count, bounds, counters);
    ctx->launches++;
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    return GUDNI_OK;
}
int selftestDiv3(gudni_ctx* ctx, unsigned long long n, unsigned long long seed, unsigned long long* devMismatches) {
    selftest_div3_kernel<<<148 * 8, 256, 0, ctx->stream>>>(n, seed, devMismatches);
    ctx->launches++;
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    return GUDNI_OK;
}

// streams and events of the batches beyond the first (which runs on the context's own stream)
static int ensureBatchStreams(gudni_ctx* ctx, int batches) {
    while ((int)ctx->batchStreams.size() < batches - 1) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        int lowest = 0, highest = 0;   // the first batch (the context's own stream) keeps the default priority: it is the critical one
        GUDNI_CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lowest, &highest));
        GUDNI_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, lowest));
        ctx->batchStreams.push_back(st);
        GUDNI_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->evJoin.push_back(ev);
    }
    if (batches > 1 && !ctx->evFork) GUDNI_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
    return GUDNI_OK;
}

int rasterTiles(gudni_ctx* ctx, const FrameParams& frame, int tileBase, int nTiles) {
    if (nTiles <= 0) return GUDNI_OK;
    FrameParams P = frame;
    // occupancy of the persistent kernels on THIS context's device (a process may hold several contexts)
    const size_t genSmem = ((sizeof(TileStage) + 15) & ~(size_t)15) + searchSummariesBytes() +
                           (size_t)(ctx->spec.threads_per_tile / 32) * sizeof(GenScratch);
    if (!ctx->occupancyKnown) {
        // the attribute belongs to the function, not to the context: allow what the largest spec (1,024 threads per tile) needs
        const size_t genSmemMax = ((sizeof(TileStage) + 15) & ~(size_t)15) + searchSummariesBytes() + (size_t)32 * sizeof(GenScratch);
        GUDNI_CUDA_TRY(ctx, cudaFuncSetAttribute(raster_generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)std::min<size_t>(genSmemMax, (size_t)227 * 1024)));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->genCtasPerSm, raster_generate_kernel,
                                                                          ctx->spec.threads_per_tile, genSmem));
        GUDNI_CUDA_TRY(ctx, cudaFuncSetAttribute(raster_sort_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->sortCtasPerSm, raster_sort_kernel,
                                                                          kSortWarpsPerCta * 32, 0));
        ctx->sortCtasPerSm = std::max(ctx->sortCtasPerSm, 1);
        GUDNI_CUDA_TRY(ctx, cudaFuncSetAttribute(raster_slice_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->sliceCtasPerSm, raster_slice_kernel,
                                                                          kSliceWarpsPerCta * 32, 0));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->colorCtasPerSm, raster_picture_kernel,
                                                                          kColorWarpsPerCta * 32, 0));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->resolveCtasPerSm, raster_resolve_kernel,
                                                                          kResolveWarpsPerCta * 32, 0));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->compositeCtasPerSm, raster_composite_kernel,
                                                                          kCompositeWarpsPerCta * 32, 0));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->accumulateCtasPerSm, raster_accumulate_kernel,
                                                                          kAccumulateWarpsPerCta * 32, 0));
        ctx->resolveCtasPerSm = std::max(ctx->resolveCtasPerSm, 1);
        ctx->compositeCtasPerSm = std::max(ctx->compositeCtasPerSm, 1);
        ctx->accumulateCtasPerSm = std::max(ctx->accumulateCtasPerSm, 1);
        GUDNI_CUDA_TRY(ctx, cudaDeviceGetAttribute(&ctx->numSms, cudaDevAttrMultiProcessorCount, ctx->device));
        ctx->genCtasPerSm = std::max(ctx->genCtasPerSm, 1);
        ctx->sliceCtasPerSm = std::max(ctx->sliceCtasPerSm, 1);
        ctx->colorCtasPerSm = std::max(ctx->colorCtasPerSm, 1);
        ctx->occupancyKnown = true;
    }
    const int numSms = ctx->numSms;
    // Batches (FrameParams::work).  One batch per ~kBatchTiles tiles, at most kMaxBatches, or what GUDNI_BATCHES says.
    // One batch, except when the frame is stored into a canvas that is not this context's own (gudni_b200_frame_target: the
    // presenting GPU's, over NVLink).  Then all ranks' accumulate kernels would push their strips into that GPU's inbound links
    // at the same moment, at the end of the frame; with the cheap half of the tiles a batch of its own, that half's pixels
    // cross while the expensive half is still being sliced (S5 on 8 GPUs: 3.95 -> 3.63 ms per frame; 4 / 8 batches 3.78 / 3.75).
    // On a single GPU batches measured neutral to slower (profiles/README.md).
    int batches = ctx->batches > 0 ? ctx->batches : (ctx->externalTarget ? 2 : 1);
    const bool ordered = ctx->batchOrdered >= 0 ? ctx->batchOrdered != 0 : true;
    batches = std::max(1, std::min({batches, gudni_dev::kMaxBatches, nTiles}));
    GUDNI_TRY(ensureBatchStreams(ctx, batches));
    // work cursors of the kernels (the threshold store and stream cursors run on across the launches of a frame)
    unsigned int* work = reinterpret_cast<unsigned int*>(ctx->counters.as<unsigned long long>() + 32);
    GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(work, 0, (size_t)gudni_dev::kMaxBatches * gudni_dev::kWorkWords * sizeof(unsigned int), ctx->stream));
    tile_order_kernel<<<1, 1024, 0, ctx->stream>>>(const_cast<gudni_tile*>(P.tiles), tileBase, nTiles, const_cast<uint32_t*>(P.tileOrder), P.counters);
    ctx->launches++;
    if (batches > 1) GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evFork, ctx->stream));
    const unsigned int regionSlabs = P.refCapSlabs / (unsigned int)batches;
    size_t tilesBefore = 0;   // (the batches' regions of the wide list lie end to end)
    for (int b = 0; b < batches; b++) {
        cudaStream_t st = b == 0 ? ctx->stream : ctx->batchStreams[b - 1];
        if (b > 0) GUDNI_CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->evFork, 0));
        // interleaved: places b, b + B, ... of the launch's cost order; ordered: the b-th run of it (batch 0 the most expensive tiles)
        int tilesHere = (nTiles - b + batches - 1) / batches;
        int firstTile = tileBase;
        P.work = work + (size_t)b * gudni_dev::kWorkWords;
        P.batchCount = batches;
        P.batchStride = batches;
        P.batchIndex = b;
        if (ordered) {
            int lo = (int)((long long)nTiles * b / batches), hi = (int)((long long)nTiles * (b + 1) / batches);
            if (batches == 2 && ctx->batchSplitPercent > 0) {   // the expensive batch's share of the tiles
                const int cut = std::max(1, std::min(nTiles - 1, (int)((long long)nTiles * ctx->batchSplitPercent / 100)));
                lo = b == 0 ? 0 : cut;
                hi = b == 0 ? cut : nTiles;
            }
            firstTile = tileBase + lo;
            tilesHere = hi - lo;
            P.batchStride = 1;
            P.batchIndex = 0;
        }
        P.refSlabBase = (unsigned int)b * regionSlabs;
        P.refCapSlabs = regionSlabs;
        P.wideList = ctx->wideList.as<unsigned int>() + tilesBefore * (size_t)(ctx->spec.threads_per_tile / 8);
        tilesBefore += (size_t)tilesHere;
        // units narrower than a warp when whole-warp units would not go round (see forEachUnit)
        long long units = (long long)tilesHere * (ctx->spec.threads_per_tile / 32);
        P.laneShift = 0;
        while (P.laneShift < GUDNI_MAX_LANE_SHIFT && units < 4ll * numSms * ctx->sliceCtasPerSm * kSliceWarpsPerCta) { P.laneShift++; units *= 2; }
        auto grid = [&](int ctasPerSm, int warpsPerCta) {
            return (int)std::min<long long>((long long)ctasPerSm * numSms, (units + warpsPerCta - 1) / warpsPerCta);
        };
        raster_generate_kernel<<<std::min(ctx->genCtasPerSm * numSms, tilesHere), ctx->spec.threads_per_tile, genSmem, st>>>(P, firstTile, tilesHere);
        raster_sort_kernel<<<grid(ctx->sortCtasPerSm, kSortWarpsPerCta), kSortWarpsPerCta * 32, 0, st>>>(P, firstTile, tilesHere);
        raster_slice_kernel<<<grid(ctx->sliceCtasPerSm, kSliceWarpsPerCta), kSliceWarpsPerCta * 32, 0, st>>>(P, firstTile, tilesHere);
        raster_slice_wide_kernel<<<numSms, 32, 0, st>>>(P);
        raster_resolve_kernel<<<grid(ctx->resolveCtasPerSm, kResolveWarpsPerCta), kResolveWarpsPerCta * 32, 0, st>>>(P, firstTile, tilesHere);
        raster_composite_kernel<<<ctx->compositeCtasPerSm * numSms, kCompositeWarpsPerCta * 32, 0, st>>>(P);
        raster_accumulate_kernel<<<grid(ctx->accumulateCtasPerSm, kAccumulateWarpsPerCta), kAccumulateWarpsPerCta * 32, 0, st>>>(P, firstTile, tilesHere);
        raster_picture_kernel<<<grid(ctx->colorCtasPerSm, kColorWarpsPerCta), kColorWarpsPerCta * 32, 0, st>>>(P, firstTile, tilesHere);
        ctx->launches += 8;
        GUDNI_CUDA_TRY(ctx, cudaGetLastError());
        if (b > 0) {
            GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evJoin[b - 1], st));
            GUDNI_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evJoin[b - 1], 0));
        }
    }
    return GUDNI_OK;
}

int rasterSpill(gudni_ctx* ctx, const FrameParams& P) {
    const int threads = 128;
    const int blocks = ctx->spillSlots / threads;
    raster_spill_kernel<<<blocks, threads, 0, ctx->stream>>>(P, ctx->spillThr.as<float4>(), ctx->spillHdr.as<uint32_t>(),
                                                             ctx->spillSlots);
    ctx->launches++;
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    return GUDNI_OK;
}

}  // namespace gudni_launch
#endif  // GUDNI_HOST_EMULATION
