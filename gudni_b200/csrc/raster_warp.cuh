// raster_warp.cuh — warp-cooperative rasterization of one (tile, 32-column group).
//
// The reference gives each work-item the whole job of its column: generate, sort, sweep, and for
// every section of the sweep a walk down the shape stack compositing every translucent layer
// (Kernels.cl:1881-1916, 1447-1513).  On deep scenes that walk is > 80 % of the arithmetic, it is
// redundant (the colour of a section is a pure function of the set of shapes present, and the
// sections of a pixel, of the pixel below and of the neighbouring columns keep meeting the same
// sets), and its length differs from lane to lane.
//
// For "dense" tiles (no more shapes than MAXSHAPE: every tile above the 8-pixel floor) a shape's
// stack bit is its position in the tile's list, so the 128-bit stack is the same key in every lane
// of the warp.  The warp keeps in shared memory
//   * the tile's substance table (premultiplied colour + meta word per bit),
//   * a direct-mapped cache of colours keyed by the stack, and
//   * a list of stacks met but not composited yet ("pending").
// Lanes sweep lane-privately in band-synchronous rounds (the branchy band bookkeeping is reached by
// all lanes together).  A section whose stack hits the cache is accumulated at once; otherwise the
// stack joins the pending list (deduplicated) and the lane appends {reference, area} to a private
// log and keeps sweeping.  When about a warp's worth of stacks is pending they are composited, one
// per lane with all lanes busy (determineColor, operation for operation as the reference), and
// every lane replays its log.  Each lane still adds colour * area into its own accumulators in
// section order (K.cl:1904), so pixels are bit-identical to the reference; what changes is how often
// a colour is recomputed and how many lanes work while it is.
#pragma once
#include "raster_device.cuh"

namespace gudni_dev {

constexpr int kWarpTableCap = 128;
#ifndef GUDNI_QUEUE_CAP
#define GUDNI_QUEUE_CAP 256
#endif
constexpr int kQueueCap = GUDNI_QUEUE_CAP;          // thresholds per column-thread before the HBM replay takes over
#ifndef GUDNI_QUEUE_HOT
#define GUDNI_QUEUE_HOT 8
#endif
constexpr int kQueueHot = GUDNI_QUEUE_HOT;           // head window in shared memory (sweep kernel; power of two)
#ifndef GUDNI_GEN_QUEUE_HOT
#define GUDNI_GEN_QUEUE_HOT 4
#endif
constexpr int kGenQueueHot = GUDNI_GEN_QUEUE_HOT;       // ... (generate kernel)
#ifndef GUDNI_CACHE_LINES
#define GUDNI_CACHE_LINES 128
#endif
constexpr int kColorCacheLines = GUDNI_CACHE_LINES;  // direct mapped
#ifndef GUDNI_SECTIONS_PER_ROUND
#define GUDNI_SECTIONS_PER_ROUND 6
#endif
constexpr int kSectionsPerRound = GUDNI_SECTIONS_PER_ROUND;   // section records a lane may hand to the resolver per round
#ifndef GUDNI_EVAL_PAIR
#define GUDNI_EVAL_PAIR 1
#endif
#ifndef GUDNI_PENDING_CAP
#define GUDNI_PENDING_CAP (GUDNI_EVAL_PAIR ? 80 : 48)
#endif
constexpr int kPendingCap = GUDNI_PENDING_CAP;   // stacks waiting to be composited
#ifndef GUDNI_PENDING_FLUSH
#define GUDNI_PENDING_FLUSH (GUDNI_EVAL_PAIR ? 48 : 27)
#endif
constexpr int kPendingFlush = GUDNI_PENDING_FLUSH;      // composite when this many are waiting (one per lane, most lanes busy)
#ifndef GUDNI_LOG_CAP
#define GUDNI_LOG_CAP (GUDNI_EVAL_PAIR ? 72 : 40)
#endif
constexpr int kLogCap = GUDNI_LOG_CAP;            // per-lane log entries between flushes
#ifndef GUDNI_STORE_SLACK
#define GUDNI_STORE_SLACK 8
#endif
constexpr int kStoreSlack = GUDNI_STORE_SLACK;   // free store entries in front of every queue (see HeadQueue)
#ifndef GUDNI_LINELESS_SLOTS
#define GUDNI_LINELESS_SLOTS 64
#endif
constexpr int kLinelessSlots = GUDNI_LINELESS_SLOTS;   // hash slots for pending stacks without a cache line (power of two, at most 128)
constexpr uint8_t kLinelessNone = 0xFF;
constexpr uint8_t kLogInline = 0xFF;   // entry carries its colour
constexpr uint8_t kLogPixelEnd = 0x7F; // markers kLogPixelEnd + n, n = 1 .. kMaxBlankRun: store the pixel n times
constexpr int kMaxBlankRun = 0xFE - 0x7F;
static_assert(kPendingCap <= 0x7F, "log tags below kLogPixelEnd are pending indices");

struct WarpScratch {
    float4 premul[kWarpTableCap];                    // 2,048 B  tile substance table
    uint32_t meta[kWarpTableCap];                    //   512 B
    ulonglong2 cacheKey[kColorCacheLines];           // 2,048 B  colour cache: stack (lo, hi)
    float4 cacheColor[kColorCacheLines];             // 2,048 B  colour; w < 0 marks an empty line
    uint8_t cacheClaim[kColorCacheLines];            //   128 B
    uint8_t lineless[kLinelessSlots];                //    64 B  pending stacks without a cache line: hash -> pending index
    ulonglong2 pendKey[kPendingCap];                 // 1,024 B  stacks to composite
    float4 pendColor[kPendingCap];                   // 1,024 B  ... and their colours once composited
    ulonglong2 recKey[32 * kSectionsPerRound];       // 2,048 B  records of the round: stack, then colour / reference
    float recArea[32 * kSectionsPerRound];           //   512 B
    float4 qThr[kQueueHot * 32];                     // 4,096 B  hot part of the 32 threshold queues
    uint32_t qHdr[kQueueHot * 32];                   // 1,024 B
};
typedef HeadQueue<kQueueCap, kQueueHot> LaneQueue;

// per-lane log of sections whose colour was not known when they were swept
struct LaneLog {
    float4 rec[kLogCap];    // inline: (r, g, b, area); pending reference: (-, -, -, area)
    uint8_t tag[kLogCap];   // kLogInline | kLogPixelEnd | pending index
};

// determineColor (K.cl:1447-1513) for a dense tile: table index = stack bit.
static __device__ __noinline__ float4 denseColor(const FrameParams& P, const WarpScratch& W, uint64_t hi, uint64_t lo,
                                                 float4 bgPremul, int absX, int absY) {
    float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t lastId = 0xFFFFFFFFu;
    // walk the set bits from the top, 32 bits at a time (FLO works on 32-bit registers)
    uint32_t word = (uint32_t)(hi >> 32);
    int wordBase = 96;
    for (;;) {
        while (word == 0u) {
            if (wordBase == 0) return compositeOverPremul(base, bgPremul);
            wordBase -= 32;
            word = (wordBase == 64) ? (uint32_t)hi : (wordBase == 32) ? (uint32_t)(lo >> 32) : (uint32_t)lo;
        }
        const int b = 31 - __clz((int)word);
        word ^= (1u << b);
        const int bit = wordBase + b;
        const uint32_t meta = W.meta[bit];
        const uint32_t id = meta & kMetaIdMask;
        if (id != lastId && (meta & kMetaSet)) {
            float4 pm = W.premul[bit];
            if (meta & kMetaPicture) pm = premultiply(readPicture(P, id, absX, absY));
            base = compositeOverPremul(base, pm);
            if (base.w == 1.0f) return base;
        }
        lastId = id;
    }
}

// The same walk for tiles whose substances are all solid and tame (substanceIsTame): no picture
// branch, unchecked shared-reciprocal division.  This loop is where the sweep kernel spends most of
// its instructions, so it is kept as lean as the reference's operation order allows.
static __device__ __noinline__ float4 denseColorTame(const WarpScratch& W, uint64_t hi, uint64_t lo, float4 bgPremul) {
    float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t lastId = 0xFFFFFFFFu;
    uint32_t word = (uint32_t)(hi >> 32);
    int wordBase = 96;
    for (;;) {
        while (word == 0u) {
            if (wordBase == 0) return compositeOverPremulT<false>(base, bgPremul);
            wordBase -= 32;
            word = (wordBase == 64) ? (uint32_t)hi : (wordBase == 32) ? (uint32_t)(lo >> 32) : (uint32_t)lo;
        }
        const int b = 31 - __clz((int)word);
        word ^= (1u << b);
        const int bit = wordBase + b;
        const uint32_t meta = W.meta[bit];
        const uint32_t id = meta & kMetaIdMask;
        if (id != lastId && (meta & kMetaSet)) {
            base = compositeOverPremulT<false>(base, W.premul[bit]);
            if (base.w == 1.0f) return base;
        }
        lastId = id;
    }
}

// Two independent stacks per lane, composited in one branch-free loop so that the two dependent chains
// (table lookup -> products -> reciprocal -> corrections) overlap: the walk is latency bound, not issue
// bound.  Each chain performs exactly the operations of denseColorTame on its own stack; a finished
// chain idles on selects until the other one is done.  The background is folded in as a last virtual
// layer.
struct TameChain {
    uint32_t w3, w2, w1, w0;   // remaining stack bits
    float4 base;
    uint32_t lastId;
    bool done;
    __device__ __forceinline__ void init(uint64_t hi, uint64_t lo, bool valid) {
        w3 = (uint32_t)(hi >> 32); w2 = (uint32_t)hi; w1 = (uint32_t)(lo >> 32); w0 = (uint32_t)lo;
        base = make_float4(0.f, 0.f, 0.f, 0.f);
        lastId = 0xFFFFFFFFu;
        done = !valid;
    }
    __device__ __forceinline__ void step(const WarpScratch& W, float4 bgPremul) {
        const bool h3 = w3 != 0u, h2 = w2 != 0u, h1 = w1 != 0u, h0 = w0 != 0u;
        const bool none = !(h3 || h2 || h1 || h0);
        const uint32_t word = h3 ? w3 : h2 ? w2 : h1 ? w1 : w0;
        const int wordBase = h3 ? 96 : h2 ? 64 : h1 ? 32 : 0;
        const int b = 31 - __clz((int)(word | 1u));   // word == 0 only when none
        const uint32_t mask = none ? 0u : (1u << b);
        w3 ^= h3 ? mask : 0u;
        w2 ^= (!h3 && h2) ? mask : 0u;
        w1 ^= (!h3 && !h2 && h1) ? mask : 0u;
        w0 ^= (!h3 && !h2 && !h1) ? mask : 0u;
        const int bit = none ? 0 : wordBase + b;
        const uint32_t meta = none ? kMetaSet | 0x3FFFFFFEu : W.meta[bit];
        const float4 pm = none ? bgPremul : W.premul[bit];
        const uint32_t id = meta & kMetaIdMask;
        const bool blend = !done && (id != lastId) && (meta & kMetaSet);
        // composite (K.cl:878-887), computed unconditionally, kept only if this layer is blended
        const float oneMinus = 1.0f - base.w;
        const float alphaOut = base.w + pm.w * oneMinus;
        float4 c;
        div3<false>((base.x * base.w) + (pm.x * oneMinus), (base.y * base.w) + (pm.y * oneMinus),
                    (base.z * base.w) + (pm.z * oneMinus), alphaOut > 0.0f ? alphaOut : 1.0f, c.x, c.y, c.z);
        c.w = alphaOut;
        if (!(alphaOut > 0.0f)) c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (blend) base = c;
        if (!done) lastId = id;
        done = done || none || (blend && base.w == 1.0f);
    }
};
static __device__ __noinline__ void denseColorTamePair(const WarpScratch& W, ulonglong2 keyA, ulonglong2 keyB, bool validB,
                                                       float4 bgPremul, float4& outA, float4& outB) {
    TameChain a, b;
    a.init(keyA.y, keyA.x, true);
    b.init(keyB.y, keyB.x, validB);
    while (!(a.done && b.done)) {
        a.step(W, bgPremul);
        b.step(W, bgPremul);
    }
    outA = a.base;
    outB = b.base;
}

// two candidate lines per stack (never the same line)
__device__ __forceinline__ void stackLines(uint64_t hi, uint64_t lo, uint32_t& line1, uint32_t& line2) {
    uint32_t t = (uint32_t)lo ^ ((uint32_t)(lo >> 32) * 0x85EBCA6Bu) ^ ((uint32_t)hi * 0xC2B2AE35u) ^
                 ((uint32_t)(hi >> 32) * 0x27D4EB2Fu);
    t ^= t >> 15;
    t *= 0x2C1B3C6Du;
    line1 = (t >> 20) & (uint32_t)(kColorCacheLines - 1);
    line2 = line1 ^ (((t >> 9) & (uint32_t)(kColorCacheLines - 1)) | 1u);
}

__device__ __forceinline__ uint32_t linelessHash(uint64_t hi, uint64_t lo) {
    uint32_t t = ((uint32_t)lo * 0x9E3779B1u) ^ ((uint32_t)(lo >> 32) * 0x7FEB352Du) ^ ((uint32_t)hi * 0x846CA68Bu) ^
                 ((uint32_t)(hi >> 32) * 0x58F38DEDu);
    t ^= t >> 16;
    t *= 0x2C1B3C6Du;
    return (t >> 24) & (uint32_t)(kLinelessSlots - 1);
}

// ---- generate kernel body -----------------------------------------------------------------------
// One warp, one (tile, 32-column group) of a dense tile: every lane builds and sorts the threshold
// queue of its column-thread (K.cl:2030-2115) in the warp's shared memory, then the warp packs the
// queues into the frame-wide store with one atomic (a warp prefix sum gives each lane its offset).
// Returns per lane 1 if the thread must be replayed against the HBM queue.
struct GenScratch {
    float4 qThr[kGenQueueHot * 32];
    uint32_t qHdr[kGenQueueHot * 32];
};
typedef WarpQueue<kQueueCap, kGenQueueHot> GenQueue;

// `column` is the reference's thread number inside the tile; the thread's record goes to threadRecs[tile * threadsPerTile + column]
__device__ __forceinline__ int packWarp(const FrameParams& P, GenQueue& q, const ThreadGeom& g, const ShapeStack& stack,
                                        uint32_t bits, bool failed, int tileIndex, int column, int& generated,
                                        bool& exhausted) {
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
#ifdef GUDNI_GEN_SORT_BINARY
            sortQueueBinary(q);
#else
            sortQueue(q);
#endif
            count = q.len;
        }
    }
    // each non-empty queue gets kStoreSlack free entries in front of it: the sweep keeps the part of the
    // queue that is not in shared memory in this slice, and slicing makes a queue grow at its head
    int reserve = count > 0 ? count + kStoreSlack : 0;
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
    const unsigned int offset = (unsigned int)(base + (unsigned long long)(incl - reserve + kStoreSlack));
    for (int i = 0; i < count; i++) {
        const Thr t = q.getT(i);
        P.thrStore[offset + i] = make_float4(t.top, t.bottom, t.left, t.right);
        P.hdrStore[offset + i] = q.getH(i);
    }
    ThreadRec rec;
    rec.hi = stack.hi; rec.lo = stack.lo;
    rec.offset = offset;
    rec.count = (g.active && !spilled) ? (unsigned int)count : kRecInactive;
    rec.chunk = 0u; rec.pad1 = 0u;
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
    const uint32_t chunkCap = (uint32_t)max(1, min(kStrandTableCap, kGenItems >> tile.h_depth));
    const bool haveBounds = P.strandBounds != nullptr;
    const float below = floatHeight + kCullMargin;
    const uint32_t numShapes = tile.shape_count;   // <= kWarpTableCap here
    // strand counts -> exclusive scan (a handful of shapes per thread of the first warp)
    __syncthreads();
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
        // ---- A: the chunk's (strand, column) pairs, shared out over the CTA ----------------------------------
        for (int item = tid; item < count * tileWidth; item += nThreads) {
            const StrandEntry& en = S.entry[item >> tile.h_depth];
            const int itemColumn = item & (tileWidth - 1);
            uint32_t flags = 0u;
            if (tile.left + itemColumn < P.width) {
                Trav l, r;
                if (strandSearch(P.geometry + 16ull * en.offset16, en.sizeWord, (float)(tile.left + itemColumn), en.right, en.lc, l, r)) {
                    const float ymin = fminf(fminf(fminf(l.ly, l.cy), fminf(l.ry, r.ly)), fminf(r.cy, r.ry));
                    const float ymax = fmaxf(fmaxf(fmaxf(l.ly, l.cy), fmaxf(l.ry, r.ly)), fmaxf(r.cy, r.ry));
                    const bool noNaN = (l.ly == l.ly) && (l.cy == l.cy) && (l.ry == l.ry) && (r.ly == r.ly) && (r.cy == r.cy) && (r.ry == r.ry);
                    flags = kItemInRange | (strandPersistentAbove(l, r) ? kItemPersistentAbove : 0u) | (noNaN ? kItemNoNaN : 0u);
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

// ---- sweep kernel body --------------------------------------------------------------------------------
// One warp, one (tile, 32-column group) of a dense tile.  Returns per lane 1 if the lane's threshold
// queue outgrew the on-chip capacity while slicing (caller hands it to the spill list).
__device__ __forceinline__ int sweepWarp(const FrameParams& P, WarpScratch& W, LaneQueue& q, LaneLog& log,
                                         const gudni_tile& tile, unsigned unit, int column) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const ThreadGeom g = threadGeom(P, tile, column);
    // ---- tile substance table + empty colour cache (warp-cooperative) -----------------------------
    bool anyPicture = false, anyWild = false;
    for (uint32_t i = lane; i < tile.shape_count; i += 32) {
        const uint32_t meta = tagMeta(__ldg(&P.shapes[tile.shape_start + i].tag));
        W.meta[i] = meta;
        anyPicture = anyPicture || (meta & kMetaPicture);
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!(meta & kMetaPicture)) {
            c = __ldg(P.substances + (meta & kMetaIdMask));
            anyWild = anyWild || !substanceIsTame(c);
            c = premultiply(c);
        }
        W.premul[i] = c;
    }
    // a picture's colour depends on the pixel, so stacks are only a valid key without pictures
    const bool cacheable = !__any_sync(full, anyPicture);
    const bool tame = cacheable && !__any_sync(full, anyWild) && substanceIsTame(P.background);
    for (int i = lane; i < kColorCacheLines; i += 32) W.cacheColor[i] = make_float4(0.f, 0.f, 0.f, -1000.f);
    if (lane < kLinelessSlots / 4) reinterpret_cast<uint32_t*>(W.lineless)[lane] = 0xFFFFFFFFu;
    // ---- the thread's sorted queue and initial stack, from the generate kernel ------------------------
    const ThreadRec rec = P.threadRecs[(size_t)unit * 32 + lane];
    ShapeStack stack;
    stack.hi = rec.hi;
    stack.lo = rec.lo;
    SweepState st;
    const float floatHeight = (float)g.intHeight;
    bool spilled = false;
    st.alive = false;
    q.init();
    if (rec.count != kRecInactive) {
        q.attach(P.thrStore, P.hdrStore, rec.offset, (int)rec.count, kStoreSlack);
        st.init(floatHeight);
    }
    __syncwarp();
    // ---- sweep -----------------------------------------------------------------------------------------
    const float4 bgPremul = premultiply(P.background);
    uint32_t* outp = P.out + (size_t)(g.originY - P.rowOrigin) * P.width + g.originX;   // only dereferenced when active
    ulonglong2* myKey = W.recKey + lane * kSectionsPerRound;
    float* myArea = W.recArea + lane * kSectionsPerRound;
#ifdef GUDNI_STATS
    unsigned long long nRec = 0, nReady = 0, nPendHit = 0, nNew = 0, nSlow = 0, nRounds = 0, nFlush = 0, nLogged = 0;
#endif
    int logLen = 0;        // entries in this lane's log
    int wrow = 0;          // pixels of the slab stored so far (rows complete in order)
    int pendingCount = 0;  // warp-uniform
    int blankRun = 1;      // pixels the band being swept stands for
    for (;;) {
        const bool anyAlive = __any_sync(full, st.alive);
        // ---- flush: composite the pending stacks, replay the logs ----------------------------------
        const bool logFull = logLen > kLogCap - (kSectionsPerRound + 2);
        if (pendingCount >= kPendingFlush || !anyAlive || __any_sync(full, logFull)) {
            if (GUDNI_EVAL_PAIR && tame) {
                for (int p0 = 0; p0 < pendingCount; p0 += 64) {
                    const int pa = p0 + lane, pb = p0 + 32 + lane;
                    if (pa < pendingCount) {
                        const bool validB = pb < pendingCount;
                        const ulonglong2 keyA = W.pendKey[pa];
                        const ulonglong2 keyB = validB ? W.pendKey[pb] : make_ulonglong2(0ull, 0ull);
                        float4 ca, cb;
                        denseColorTamePair(W, keyA, keyB, validB, bgPremul, ca, cb);
                        W.pendColor[pa] = ca;
                        // un-pin: the line that references this entry (if it got one) now holds the colour
                        uint32_t lineA, lineA2;
                        stackLines(keyA.y, keyA.x, lineA, lineA2);
                        if (W.cacheColor[lineA].w == -(float)(1 + pa)) W.cacheColor[lineA] = make_float4(ca.x, ca.y, ca.z, 1.f);
                        else if (W.cacheColor[lineA2].w == -(float)(1 + pa)) W.cacheColor[lineA2] = make_float4(ca.x, ca.y, ca.z, 1.f);
                        if (validB) {
                            W.pendColor[pb] = cb;
                            uint32_t lineB, lineB2;
                            stackLines(keyB.y, keyB.x, lineB, lineB2);
                            if (W.cacheColor[lineB].w == -(float)(1 + pb)) W.cacheColor[lineB] = make_float4(cb.x, cb.y, cb.z, 1.f);
                            else if (W.cacheColor[lineB2].w == -(float)(1 + pb)) W.cacheColor[lineB2] = make_float4(cb.x, cb.y, cb.z, 1.f);
                        }
                    }
                }
            } else {
                for (int p0 = 0; p0 < pendingCount; p0 += 32) {
                    const int p = p0 + lane;
                    if (p < pendingCount) {
                        const ulonglong2 key = W.pendKey[p];
                        const float4 c = tame ? denseColorTame(W, key.y, key.x, bgPremul) : denseColor(P, W, key.y, key.x, bgPremul, 0, 0);
                        W.pendColor[p] = c;
                        // un-pin: the line that references this entry (if it got one) now holds the colour
                        uint32_t line, line2;
                        stackLines(key.y, key.x, line, line2);
                        if (W.cacheColor[line].w == -(float)(1 + p)) W.cacheColor[line] = make_float4(c.x, c.y, c.z, 1.f);
                        else if (W.cacheColor[line2].w == -(float)(1 + p)) W.cacheColor[line2] = make_float4(c.x, c.y, c.z, 1.f);
                    }
                }
            }
            __syncwarp();
            // replay in section order (K.cl:1904).  The log is in local memory (L2 latency): the next entry is
            // fetched, tag and record together, before the current one is applied.
            uint8_t tagNext = log.tag[0];
            float4 recNext = log.rec[0];
            for (int j = 0; j < logLen; j++) {
                const uint8_t tag = tagNext;
                float4 r = recNext;
                const int jn = min(j + 1, kLogCap - 1);
                tagNext = log.tag[jn];
                recNext = log.rec[jn];
                if (tag > kLogPixelEnd && tag != kLogInline) {
                    const uint32_t word = pixelWord(st.accR, st.accG, st.accB, st.accArea);
                    const int rep = (int)tag - (int)kLogPixelEnd;
                    for (int r = 0; r < rep; r++) outp[(size_t)(wrow + r) * P.width] = word;
                    st.accR = st.accG = st.accB = st.accArea = 0.f;
                    wrow += rep;
                    continue;
                }
                if (tag != kLogInline) {
                    const float4 c = W.pendColor[tag];
                    r.x = c.x; r.y = c.y; r.z = c.z;
                }
                st.accR += r.x * r.w;
                st.accG += r.y * r.w;
                st.accB += r.z * r.w;
                st.accArea += r.w;
            }
#ifdef GUDNI_STATS
            if (lane == 0) nFlush++;
            nLogged += logLen;
#endif
            logLen = 0;
            pendingCount = 0;
            if (lane < kLinelessSlots / 4) reinterpret_cast<uint32_t*>(W.lineless)[lane] = 0xFFFFFFFFu;
            __syncwarp();
            if (!anyAlive) break;
        }
#ifdef GUDNI_STATS
        if (lane == 0) nRounds++;
#endif
        // ---- (A) band boundary: close the pixel, open the next band ---------------------------------
        // A pixel no threshold touches is one band with one section of area exactly 1: its accumulators are
        // colour * 1 and 1.  When the next threshold starts m or more whole pixels further down, the next m
        // pixels of the column are that same pixel (same stack, same arithmetic), so the band is swept once
        // and the pixel stored m times (`blankRun`, set when the band is opened).  Picture substances
        // depend on the row, so their tiles do not take the shortcut.
        if (st.alive && st.ex == 1.0f) {
            if (st.ey >= st.pixelY) {   // calculatePixel's loop condition failed: the pixel is complete
                if (logLen == 0) {      // everything of this pixel is accumulated
                    const uint32_t word = pixelWord(st.accR, st.accG, st.accB, st.accArea);
                    outp[(size_t)wrow * P.width] = word;
                    st.accR = st.accG = st.accB = st.accArea = 0.f;
                    wrow++;
                    if (blankRun > 1) {
                        for (int r = 1; r < blankRun; r++) outp[(size_t)(wrow + r - 1) * P.width] = word;
                        wrow += blankRun - 1;
                    }
                } else {
                    log.tag[logLen++] = (uint8_t)(kLogPixelEnd + blankRun);
                }
                nextPixel(st, floatHeight);
                if (blankRun > 1) {     // ... and the blankRun - 1 pixels after it
                    const float skipped = (float)(blankRun - 1);
                    st.sy += skipped;
                    st.ey = st.sy;
                    st.pixelY += skipped;
                    st.row += blankRun - 1;
                    st.alive = st.pixelY <= floatHeight;
                }
            }
            if (st.alive) {
                sweepVertical(q, stack, st, floatHeight);
                if (q.failed()) { spilled = true; st.alive = false; }
                blankRun = 1;
                if (cacheable && st.sy == st.pixelY - 1.0f) {
                    const float limit = fminf(st.gapTop, floatHeight);
                    if (limit >= st.pixelY + 1.0f) blankRun = min((int)(limit - st.pixelY) + 1, kMaxBlankRun);
                }
            }
        }
        // ---- (B) up to kSectionsPerRound sections of the band ---------------------------------------
        int count = 0;
        while (st.alive && count < kSectionsPerRound) {
            float area;
            uint64_t hi, lo;
            sweepSection(q, stack, st, area, hi, lo);
            if (area != 0.0f) {   // a zero-area section adds colour * 0 = 0 to every accumulator
                myKey[count] = make_ulonglong2(lo, hi);
                myArea[count] = area;
                count++;
            }
            if (st.ex == 1.0f) break;   // band finished: next round starts at (A)
        }
        // ---- (C) resolve the records: cache hit -> colour, else -> reference to a pending stack -------
        if (!cacheable) {
            // picture substances: the colour depends on the pixel; every lane composites its own records
            // (all of them lie in the pixel row the lane is sweeping)
            for (int j = 0; j < count; j++) {
                const ulonglong2 key = myKey[j];
                const float4 c = denseColor(P, W, key.y, key.x, bgPremul, g.originX, g.originY + st.row);
                *reinterpret_cast<float4*>(&myKey[j]) = make_float4(c.x, c.y, c.z, 1.f);
            }
        } else {
            int incl = count;
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += t;
            }
            const int excl = incl - count;
            const int total = __shfl_sync(full, incl, 31);
            __syncwarp();
            for (int f0 = 0; f0 < total; f0 += 32) {
                // record f belongs to the first lane o with incl(o) > f
                const int f = f0 + lane;
                int lo_ = 0, hi_ = 31;
                for (int step = 0; step < 5; step++) {
                    const int mid = (lo_ + hi_) >> 1;
                    const int v = __shfl_sync(full, incl, mid);
                    if (v > f) hi_ = mid; else lo_ = mid + 1;
                }
                const int ownerExcl = __shfl_sync(full, excl, lo_);
                const bool valid = f < total;
                const int slot = valid ? lo_ * kSectionsPerRound + (f - ownerExcl) : 0;
                // Cache line states (cacheColor.w): < 0 empty (-1000) or pending (-(1 + pending index),
                // pinned until the flush); 1 ready.
                ulonglong2 key = make_ulonglong2(0ull, 0ull);
                uint32_t line = 0;
                bool miss = false, slow = false;
                float4 out = make_float4(0.f, 0.f, 0.f, 0.f);   // w = 1: colour; w = -(1 + pending index): reference
                if (valid) {
                    key = W.recKey[slot];
                    uint32_t line2;
                    stackLines(key.y, key.x, line, line2);
                    const float4 c = W.cacheColor[line];
                    const ulonglong2 k = W.cacheKey[line];
#ifdef GUDNI_STATS
                    nRec++;
#endif
                    if (k.x == key.x && k.y == key.y && c.w > -999.f) {
                        out = c;                                             // ready colour or pending reference
                    } else {
                        const float4 c2 = W.cacheColor[line2];
                        const ulonglong2 k2 = W.cacheKey[line2];
                        if (k2.x == key.x && k2.y == key.y && c2.w > -999.f) {
                            out = c2;
                        } else {
                            // a new stack takes an empty line if it has one, else evicts a ready colour; a line
                            // that waits for its colour (pinned until the flush) cannot be taken
                            const bool empty1 = c.w <= -999.f, empty2 = c2.w <= -999.f;
                            const bool pinned1 = c.w < 0.f && !empty1, pinned2 = c2.w < 0.f && !empty2;
                            if (empty1) miss = true;
                            else if (empty2) { miss = true; line = line2; }
                            else if (!pinned1) miss = true;
                            else if (!pinned2) { miss = true; line = line2; }
                            else slow = true;
                        }
                    }
#ifdef GUDNI_STATS
                    if (out.w > 0.f) nReady++; else if (out.w < 0.f) nPendHit++;
#endif
                }
                if (__any_sync(full, miss)) {
                    // new stacks: one claimant per line appends it to the pending list and pins the line
                    if (miss) W.cacheClaim[line] = (uint8_t)lane;
                    __syncwarp();
                    const bool winner = miss && W.cacheClaim[line] == (uint8_t)lane;
                    const unsigned winners = __ballot_sync(full, winner);
                    const int idx = pendingCount + __popc(winners & ((1u << lane) - 1u));
                    if (winner) {
                        if (idx < kPendingCap) {
                            W.pendKey[idx] = key;
                            W.cacheKey[line] = key;
                            W.cacheColor[line] = make_float4(0.f, 0.f, 0.f, -(float)(1 + idx));
                            out.w = -(float)(1 + idx);
#ifdef GUDNI_STATS
                            nNew++;
#endif
                        } else {   // more than kPendingCap new stacks in flight: composite on the spot (rare)
                            const float4 c = denseColor(P, W, key.y, key.x, bgPremul, 0, 0);
                            out = make_float4(c.x, c.y, c.z, 1.f);
                        }
                    }
                    pendingCount = min(pendingCount + __popc(winners), kPendingCap);
                    __syncwarp();
                    if (miss && !winner) {   // lost the line: to the same stack (share it) or to another one
                        const ulonglong2 k = W.cacheKey[line];
                        const float4 c = W.cacheColor[line];
                        if (k.x == key.x && k.y == key.y && c.w > -999.f) out = c;
                        else slow = true;
                    }
                }
                // leftovers: a stack that could not get a cache line is found again through a small hash of
                // the line-less pending entries; what is not there is appended, one distinct stack at a time
                // (a collision in that hash at worst appends a stack twice, which only costs its compositing)
                uint32_t slotL = 0;
                if (slow) {
                    slotL = linelessHash(key.y, key.x);
                    const uint8_t at = W.lineless[slotL];
                    if (at != kLinelessNone) {
                        const ulonglong2 k = W.pendKey[at];
                        if (k.x == key.x && k.y == key.y) { out.w = -(float)(1 + (int)at); slow = false; }
                    }
                }
#ifdef GUDNI_STATS
                if (slow) nSlow++;
#endif
                unsigned todo = __ballot_sync(full, slow);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    const unsigned long long kx = __shfl_sync(full, key.x, src), ky = __shfl_sync(full, key.y, src);
                    const bool mine = slow && key.x == kx && key.y == ky;
                    todo &= ~__ballot_sync(full, mine);
                    const int idx = pendingCount < kPendingCap ? pendingCount : -1;
                    if (idx >= 0) {
                        if (lane == src) { W.pendKey[idx] = key; W.lineless[slotL] = (uint8_t)idx; }
                        pendingCount++;
                    }
                    if (mine) {
                        if (idx >= 0) out.w = -(float)(1 + idx);
                        else { const float4 c = denseColor(P, W, ky, kx, bgPremul, 0, 0); out = make_float4(c.x, c.y, c.z, 1.f); }
                    }
                }
                if (valid) *reinterpret_cast<float4*>(&W.recKey[slot]) = out;
            }
            __syncwarp();
        }
        // ---- (D) accumulate what is known, log the rest, in section order ----------------------------
        for (int j = 0; j < count; j++) {
            const float4 r = *reinterpret_cast<const float4*>(&myKey[j]);
            const float area = myArea[j];
            if (r.w > 0.f && logLen == 0) {
                st.accR += r.x * area;
                st.accG += r.y * area;
                st.accB += r.z * area;
                st.accArea += area;
            } else {
                log.rec[logLen] = make_float4(r.x, r.y, r.z, area);
                log.tag[logLen] = (r.w > 0.f) ? kLogInline : (uint8_t)(int)(-r.w - 1.f);
                logLen++;
            }
        }
        __syncwarp();
    }
#ifdef GUDNI_STATS
    atomicAdd(&P.counters[8], nRec); atomicAdd(&P.counters[9], nReady); atomicAdd(&P.counters[10], nPendHit);
    atomicAdd(&P.counters[11], nNew); atomicAdd(&P.counters[12], nSlow); atomicAdd(&P.counters[13], nRounds);
    atomicAdd(&P.counters[14], nFlush); atomicAdd(&P.counters[15], nLogged);
#endif
    return spilled ? 1 : 0;
}

}  // namespace gudni_dev
