// raster_warp.cuh — the generate kernel's body: one CTA per tile, generateThresholds + sortThresholds
// (Kernels.cl:2030-2115) for the tile's column-threads, their sorted queues packed into the frame-wide
// threshold store.  The per-column arithmetic is in raster_device.cuh; the render phase is in raster_split.cuh.
#pragma once
#include "raster_device.cuh"

namespace gudni_dev {

constexpr int kWarpTableCap = 128;                  // shapes a dense tile may list (stack bits)
#ifndef GUDNI_QUEUE_CAP
#define GUDNI_QUEUE_CAP 256
#endif
constexpr int kQueueCap = GUDNI_QUEUE_CAP;          // thresholds per column-thread before the HBM replay takes over
#ifndef GUDNI_GEN_QUEUE_HOT
#define GUDNI_GEN_QUEUE_HOT 4
#endif
constexpr int kGenQueueHot = GUDNI_GEN_QUEUE_HOT;       // ... (generate kernel)

// ---- sort + pack -----------------------------------------------------------------------------------
// The warp packs the threshold queues of its column-threads, in the order they were built, into the frame-wide
// store with one atomic (a warp prefix sum gives each lane its offset); raster_sort_kernel sorts them there
// (K.cl:2084-2115).  Returns per lane 1 if the thread must be replayed against the HBM queue.
struct GenScratch {
    float4 qThr[kGenQueueHot * 32];
    uint32_t qHdr[kGenQueueHot * 32];
};
typedef WarpQueue<kQueueCap, kGenQueueHot> GenQueue;

// `column` is the reference's thread number inside the tile; the thread's record goes to threadRecs[tile * threadsPerTile + column]
__device__ __forceinline__ int packWarp(const FrameParams& P, GenQueue& q, const ThreadGeom& g, const ShapeStack& stack,
                                        uint32_t bits, bool failed, int tileIndex, int column, int& generated,
                                        bool& exhausted, bool tilePictures) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool spilled = false;
    generated = -1;
    int count = 0;
    if (g.active) {
        if (failed) {
            spilled = true;
        } else {
            generated = q.len;
            const int threadId = P.tileThreadBase[tileIndex] + column;
            if (P.dbgThresholds) P.dbgThresholds[threadId] = q.len;
            if (P.dbgShapeBits) P.dbgShapeBits[threadId] = (int32_t)bits;
            // the queue leaves in the order it was built; raster_sort_kernel sorts it in the store (raster_sort.cuh)
            count = q.len;
        }
    }
    // the queues lie end to end: the sort works in place and the slice kernel only reads
    const int reserve = count;
    int incl = reserve;
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(full, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(full, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(&P.counters[kCntStoreCursor], (unsigned long long)total);
    base = __shfl_sync(full, base, 0);
    exhausted = false;
    if (base + (unsigned long long)total > P.storeCap) {   // store exhausted: the replay kernel takes the warp's threads
        exhausted = g.active && !spilled;
        spilled = spilled || g.active;
        count = 0;
    }
    const unsigned int offset = (unsigned int)(base + (unsigned long long)(incl - reserve));
    for (int i = 0; i < count; i++) {
        const Thr t = q.getT(i);
        P.thrStore[offset + i] = make_float4(t.top, t.bottom, t.left, t.right);
        P.hdrStore[offset + i] = q.getH(i);
    }
    ThreadRec rec;
    rec.hi = stack.hi; rec.lo = stack.lo;
    rec.offset = offset;
    rec.count = (g.active && !spilled) ? (unsigned int)count : kRecInactive;
    rec.chunk = 0u; rec.pad1 = tilePictures ? kRecTilePictures : 0u;
    P.threadRecs[((size_t)tileIndex << P.computeDepth) + (size_t)column] = rec;
    return spilled ? 1 : 0;
}

// ---- generate kernel body, one CTA per tile ---------------------------------------------------------
// The CTA's threads are the tile's column-threads (blockDim = threadsPerTile), so every thread walks the same
// shape list.  The reference has each work-item fetch every strand's header from global memory itself
// (K.cl:1557-1583: size word, right end, left end + control); here the CTA stages the headers of a run of
// consecutive shapes in shared memory once — one thread per shape walks that shape's strands (the next
// strand's address hangs on the size word of the one before) and writes them at the offset a scan of the
// strand counts gives — and then every thread runs down the table: the range tests read shared memory, and only
// a column the strand really crosses goes on to the strand's tree in global memory.
struct StrandEntry {   // 48 bytes
    float4 lc;              // left end, control point of the first curve (K.cl:1371-1373)
    float2 right;           // right end
    float2 yb;              // (min y, max y) over the strand's points (strand_bounds_kernel)
    uint32_t sizeWord;
    uint32_t offset16;      // the strand's place in the geometry heap, in 16-byte units
    uint32_t shapeAndFlags; // position of the shape in the tile's list
    uint32_t pad;
};
constexpr uint32_t kEntryShapeMask = 0xFFu;
#ifndef GUDNI_STRAND_TABLE
#define GUDNI_STRAND_TABLE 64
#endif
constexpr int kStrandTableCap = GUDNI_STRAND_TABLE;   // strands staged at a time (at most 64: one bit each in a lane's to-do mask)
#ifndef GUDNI_GEN_ITEMS
#define GUDNI_GEN_ITEMS 2048
#endif
constexpr int kGenItems = GUDNI_GEN_ITEMS;            // (strand, column) pairs searched at a time
struct TileStage {
    StrandEntry entry[kStrandTableCap];
    uint32_t shapeBase[kWarpTableCap + 1];   // exclusive scan of the strand counts of the tile's shapes
    int tileSlot;                            // the CTA's current tile
    int anyPicture;                          // ... lists a picture substance
};

struct GenThread {   // what buildThresholdArray carries through the tile's shape list (K.cl:1540-1595)
    ShapeStack stack;   // bit n = enclosedByShape of shape n: the XOR over its strands of their `enclosed` (K.cl:1582-1586)
    ShapeStack added;   // bit n: shape n stored a threshold in this thread (K.cl:1587)
    bool failed;
    // ShapeState.shapeBits as the reference would have stored it: shapes that added a threshold or enclose the slab's top
    __device__ __forceinline__ uint32_t bits() const {
        return (uint32_t)(__popcll((long long)(stack.lo | added.lo)) + __popcll((long long)(stack.hi | added.hi)));
    }
    __device__ __forceinline__ void note(uint32_t n, const GenFlags& f) {
        if (f.enclosed) stack.flip(n);
        if (f.added) {
            if (n < 64) added.lo |= 1ull << n;
            else added.hi |= 1ull << (n & 63);
        }
    }
};

// What the search of a (strand, column) pair leaves for the column's threads to classify the strand with:
// structure of arrays, item = strandInChunk * tileWidth + column.
struct SearchSummaries {
    float2* yRange;    // min and max over the six y of the two pieces the searches ended on
    uint32_t* flags;   // kItem*
};
constexpr uint32_t kItemInRange = 1u, kItemPersistentAbove = 2u, kItemNoNaN = 4u;
// where the two searches ended (Trav::idx), when both fit a byte: bits 8-15 and 16-23 (strandSearchFrom)
constexpr uint32_t kItemEnds = 8u;
__host__ __device__ constexpr size_t searchSummariesBytes() { return (size_t)kGenItems * 12; }

// Whole CTA.  Afterwards q holds the thread's thresholds in push order (or t.failed is set).
//
// The tile's column-threads are tileWidth columns x S slabs (S = threadsPerTile / tileWidth, K.cl:1692-1722), a
// warp being 32 neighbouring columns of one slab.  The S threads of a column all need the same two tree searches
// for every strand — the searches compare x only — and all but one or two of them only need to know that the
// strand passes above or below them.  So the strands are taken in chunks of E (as many as the header table and
// kGenItems / tileWidth allow):
//   A. search: the CTA's threads share out the chunk's (strand, column) pairs; each searches (strandSearch) and
//      leaves the y range of the two pieces and two x-only facts in shared memory;                        barrier
//   B. classify: every column-thread runs down the chunk — a few instructions per strand: out of the column's
//      range or below the slab: nothing; above the slab: the enclosure parity flips if the crossing is persistent
//      (strandPersistentAbove); else the strand goes into a bit mask;
//   C. spawn: while any lane of the warp has a bit set, every lane takes its next strand, searches it again for
//      its own column, subtracts its oy from the pieces' y (bit for bit the reference's `- threadDelta`) and
//      spawns (K.cl:1264-1333).  A strand that crosses a slab does so in a few neighbouring columns; taking the
//      strands of a chunk together is what keeps more than a handful of lanes busy in the curve bisection. barrier
// The reference resets `enclosed` per strand, XORs it into the shape's parity and flips the shape's stack bit at the
// shape's last strand (K.cl:1557-1592); since every shape has its own bit that is the same as flipping the bit once
// per enclosing strand, in any order — which is what lets B and C take the strands out of order.
__device__ __forceinline__ void generateTileThresholds(const FrameParams& P, TileStage& S, const SearchSummaries& R, GenQueue& q,
                                                       GenThread& t, const gudni_tile& tile, const ThreadGeom& g) {
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, nThreads = blockDim.x;
    const float ox = (float)g.originX, oy = (float)g.originY, floatHeight = (float)g.intHeight;
    const int tileWidth = 1 << tile.h_depth;
    const int column = tid & (tileWidth - 1);
    // A tile of one or two slabs has nothing to share: its threads run down the staged headers on their own
    // (strandThresholds: search, classify, spawn per strand), a whole table of strands between two barriers.
    const bool shareSearches = (nThreads >> tile.h_depth) > 2;
    const uint32_t chunkCap = shareSearches ? (uint32_t)max(1, min(kStrandTableCap, kGenItems >> tile.h_depth)) : (uint32_t)kStrandTableCap;
    const bool haveBounds = P.strandBounds != nullptr;
    const float below = floatHeight + kCullMargin;
    const uint32_t numShapes = tile.shape_count;   // <= kWarpTableCap here
    // strand counts -> exclusive scan (a handful of shapes per thread of the first warp)
    __syncthreads();
    for (uint32_t i = (uint32_t)tid; i < numShapes; i += (uint32_t)nThreads)   // (S.anyPicture was cleared with the tile fetch)
        if (tagMeta(__ldg(&P.shapes[tile.shape_start + i].tag)) & kMetaPicture) S.anyPicture = 1;
    if (tid < 32) {
        uint32_t mine[(kWarpTableCap + 31) / 32];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < (kWarpTableCap + 31) / 32; k++) {
            const uint32_t i = (uint32_t)tid * ((kWarpTableCap + 31) / 32) + k;
            mine[k] = i < numShapes ? __ldg(&P.shapes[tile.shape_start + i].num_strands) : 0u;
            sum += mine[k];
        }
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(full, incl, d);
            if (tid >= d) incl += v;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < (kWarpTableCap + 31) / 32; k++) {
            const uint32_t i = (uint32_t)tid * ((kWarpTableCap + 31) / 32) + k;
            if (i <= numShapes) S.shapeBase[i] = run;
            run += mine[k];
        }
        if (tid == 31 && numShapes == (uint32_t)kWarpTableCap) S.shapeBase[kWarpTableCap] = run;
    }
    __syncthreads();
    uint32_t shapeBegin = 0;
    while (shapeBegin < numShapes) {
        // the longest run of shapes whose strands fit a chunk (uniform across the CTA)
        const uint32_t first = S.shapeBase[shapeBegin];
        uint32_t shapeEnd = shapeBegin;
        while (shapeEnd < numShapes && S.shapeBase[shapeEnd + 1] - first <= chunkCap) shapeEnd++;
        if (shapeEnd == shapeBegin) {
            // one shape with more strands than a chunk holds: every thread reads its headers and searches itself
            const uint4 rec = __ldg(reinterpret_cast<const uint4*>(P.shapes + tile.shape_start + shapeBegin));
            const uint8_t* strand = P.geometry + 16ull * rec.z;
            if (g.active && !t.failed) {
                GenFlags f{false, false};
                for (uint32_t k = 0; k < rec.w; k++) {
                    const float4 h0 = __ldg(reinterpret_cast<const float4*>(strand));
                    const float4 lc = __ldg(reinterpret_cast<const float4*>(strand + 16));
                    const float2 yb = haveBounds ? __ldg(P.strandBounds + ((size_t)(strand - P.geometry) >> 4)) : make_float2(0.f, 0.f);
                    const uint32_t sizeWord = __float_as_uint(h0.x);
                    f.enclosed = false;
                    strandThresholds(q, strand, sizeWord, ox, oy, floatHeight, shapeBegin, f, make_float2(h0.z, h0.w), lc, haveBounds, yb);
                    if (q.failed()) { t.failed = true; break; }
                    t.note(shapeBegin, f);
                    strand += 8u * (sizeWord & 0xFFFFu);
                }
            }
            shapeBegin += 1;
            continue;
        }
        // stage the strand headers: one thread per shape of the run
        for (uint32_t i = shapeBegin + tid; i < shapeEnd; i += nThreads) {
            const uint4 rec = __ldg(reinterpret_cast<const uint4*>(P.shapes + tile.shape_start + i));
            const uint8_t* strand = P.geometry + 16ull * rec.z;
            StrandEntry* e = S.entry + (S.shapeBase[i] - first);
            for (uint32_t k = 0; k < rec.w; k++) {
                const float4 h0 = __ldg(reinterpret_cast<const float4*>(strand));
                e[k].lc = __ldg(reinterpret_cast<const float4*>(strand + 16));
                e[k].right = make_float2(h0.z, h0.w);
                const uint32_t off16 = (uint32_t)((size_t)(strand - P.geometry) >> 4);
                e[k].yb = haveBounds ? __ldg(P.strandBounds + off16) : make_float2(0.f, 0.f);
                e[k].sizeWord = __float_as_uint(h0.x);
                e[k].offset16 = off16;
                e[k].shapeAndFlags = i;
                strand += 8u * (__float_as_uint(h0.x) & 0xFFFFu);
            }
        }
        __syncthreads();
        const int count = (int)(S.shapeBase[shapeEnd] - first);
        if (!shareSearches) {
            if (g.active && !t.failed) {
                for (int k = 0; k < count; k++) {
                    const StrandEntry& en = S.entry[k];
                    GenFlags f{false, false};
                    strandThresholds(q, P.geometry + 16ull * en.offset16, en.sizeWord, ox, oy, floatHeight, en.shapeAndFlags & kEntryShapeMask,
                                     f, en.right, en.lc, haveBounds, en.yb);
                    if (q.failed()) { t.failed = true; break; }
                    t.note(en.shapeAndFlags & kEntryShapeMask, f);
                }
            }
            __syncthreads();
            shapeBegin = shapeEnd;
            continue;
        }
        // ---- A: the chunk's (strand, column) pairs, shared out over the CTA ----------------------------------
        for (int item = tid; item < count * tileWidth; item += nThreads) {
            const StrandEntry& en = S.entry[item >> tile.h_depth];
            const int itemColumn = item & (tileWidth - 1);
            uint32_t flags = 0u;
            if (tile.left + itemColumn < P.width) {
                Trav l, r;
#ifdef GUDNI_SEARCH_X
                if (strandSearchX(P.geometry + 16ull * en.offset16, P.strandBounds + en.offset16 + 2, en.sizeWord, (float)(tile.left + itemColumn), en.right, en.lc, l, r)) {
#else
                if (strandSearch(P.geometry + 16ull * en.offset16, en.sizeWord, (float)(tile.left + itemColumn), en.right, en.lc, l, r)) {
#endif
                    const float ymin = fminf(fminf(fminf(l.ly, l.cy), fminf(l.ry, r.ly)), fminf(r.cy, r.ry));
                    const float ymax = fmaxf(fmaxf(fmaxf(l.ly, l.cy), fmaxf(l.ry, r.ly)), fmaxf(r.cy, r.ry));
                    const bool noNaN = (l.ly == l.ly) && (l.cy == l.cy) && (l.ry == l.ry) && (r.ly == r.ly) && (r.cy == r.cy) && (r.ry == r.ry);
                    flags = kItemInRange | (strandPersistentAbove(l, r) ? kItemPersistentAbove : 0u) | (noNaN ? kItemNoNaN : 0u);
#ifndef GUDNI_NO_SEARCH_ENDS
                    if (l.idx < 256 && r.idx < 256) flags |= kItemEnds | ((uint32_t)l.idx << 8) | ((uint32_t)r.idx << 16);
#endif
                    R.yRange[item] = make_float2(ymin, ymax);
                }
            }
            R.flags[item] = flags;
        }
        __syncthreads();
        // ---- B: classify the chunk's strands for this thread's slab ---------------------------------------------
        unsigned long long todo = 0ull;
        if (g.active && !t.failed) {
            for (int k = 0; k < count; k++) {
                const int item = (k << tile.h_depth) + column;
                const uint32_t flags = R.flags[item];
                if (!(flags & kItemInRange)) continue;
                const StrandEntry& en = S.entry[k];
                const float2 yb = en.yb;
                if (haveBounds && (yb.x - oy) >= below) continue;   // the strand's whole y range below the slab
                const float2 yr = R.yRange[item];
                const bool noNaN = (flags & kItemNoNaN) != 0u;
                // all six y of the two pieces >= below  <=>  the smallest is (x -> fl(x - oy) is monotone)
                if (noNaN && (yr.x - oy) >= below) continue;
                if ((noNaN && (yr.y - oy) <= -kCullMargin) || (haveBounds && (yb.y - oy) <= -kCullMargin)) {
                    if (flags & kItemPersistentAbove) t.stack.flip(en.shapeAndFlags & kEntryShapeMask);
                    continue;
                }
                todo |= 1ull << k;
            }
        }
        // ---- C: spawn; every lane its own strands, in order --------------------------------------------------------
        while (__any_sync(full, todo != 0ull)) {
            if (todo) {
                const int k = __ffsll((long long)todo) - 1;
                todo &= todo - 1ull;
                const StrandEntry& en = S.entry[k];
                const uint32_t n = en.shapeAndFlags & kEntryShapeMask;
                Trav l, r;
                const uint32_t found = R.flags[(k << tile.h_depth) + column];
                if (found & kItemEnds)
                    strandSearchFrom(P.geometry + 16ull * en.offset16, ox, en.right, en.lc, (int)((found >> 8) & 0xFFu),
                                     (int)((found >> 16) & 0xFFu), l, r);
                else
                    strandSearch(P.geometry + 16ull * en.offset16, en.sizeWord, ox, en.right, en.lc, l, r);   // in range: see B
                l.ly -= oy; l.cy -= oy; l.ry -= oy;
                r.ly -= oy; r.cy -= oy; r.ry -= oy;
                GenFlags f{false, false};
                strandSpawnCore(q, l, r, floatHeight, n, f);
                if (q.failed()) { t.failed = true; todo = 0ull; }
                else t.note(n, f);
            }
        }
        __syncthreads();
        shapeBegin = shapeEnd;
    }
}

}  // namespace gudni_dev
