// raster_split.cuh — renderThresholds (Kernels.cl:2117-2167) as two kernels with a section stream between them.
//
// The reference's render phase does two very different things per column-thread: a branchy, sequential walk
// down the sorted threshold queue that cuts the column into bands and sections (verticalAdvance /
// horizontalAdvance / sliceActive, K.cl:1007-1077, 1744-1851), and for every section a walk down the shape
// stack compositing translucent layers (determineColor, K.cl:1447-1513).  The first is integer / queue work in
// which neighbouring lanes rarely agree on the branch; the second is a long dependent fp32 chain that is the
// same code in every lane.  Run in one kernel they share one register budget and one shared-memory budget
// (12 warps per SM in round 1) and the divergent half drags the colour state through every branch.
//
//   raster_slice_kernel   one lane per column-thread: the sweep's state machine only.  No shape stack, no
//                         colour: it never needs to know WHICH shapes are present to decide where bands and
//                         sections lie.  Emits the column's *section stream*: 8-byte records
//                           SEC  (toggle bit | first-of-band, area)   one section; afterwards the bit flips
//                           FLIP (bit)                                a persistent threshold crossing a band border
//                           PEND (n)                                  pixel complete; store it n times (runs of untouched rows)
//                           LINK (next chunk) / END
//                         into a chain of 128-byte chunks (16 records, the last one the link) drawn from one pool.
//   raster_resolve_kernel / raster_composite_kernel / raster_accumulate_kernel
//                         replay the streams: number the distinct shape stacks, composite each once, add
//                         colour * area per section and store the pixels (see "colours" below).
//
// Every lane still adds colour * area into its own accumulators in section order (K.cl:1904), so pixels are
// bit-identical to the reference.
#pragma once
#include "raster_warp.cuh"

namespace gudni_dev {

// ---- the section stream ---------------------------------------------------------------------------------------
constexpr int kChunkRecs = 16;                     // 128 bytes: one L2 line
constexpr uint32_t kRecKindMask = 0xC0000000u;
constexpr uint32_t kRecSection = 0x00000000u;      // .x = kind | first | bit, .y = area (f32 bits)
constexpr uint32_t kRecFlip = 0x40000000u;         // .x = kind | bit
constexpr uint32_t kRecPixelEnd = 0x80000000u;     // .x = kind | repeat count
constexpr uint32_t kRecLink = 0xC0000000u;         // .y = next chunk, kStreamEnd: the stream ends here
constexpr uint32_t kRecFirst = 0x00000100u;        // SEC: first section of its band
constexpr uint32_t kRecBitMask = 0x000000FFu;
constexpr uint32_t kRecNoBit = 0x000000FFu;        // SEC without a threshold behind it (the band's last section)
constexpr uint32_t kStreamEnd = 0xFFFFFFFFu;
constexpr int kMaxPixelRun = 0xFFFF;

#ifndef GUDNI_SLAB_CHUNKS
#define GUDNI_SLAB_CHUNKS 128
#endif
constexpr unsigned int kSlabChunks = GUDNI_SLAB_CHUNKS;   // chunks a warp draws from the pool at a time
constexpr unsigned int kSlabLow = 40;                      // refill when fewer are left at the top of a round

// ---- the slice kernel's view of a column-thread's threshold queue ------------------------------------------------------
// The reference keeps one sorted array per work-item and, every time a run of thresholds with the same top becomes
// active, cuts each of them at the next event below (sliceActive, K.cl:1053-1067) and inserts the lower parts back
// into the array behind the run (insertThreshold, K.cl:1105-1124) — every insertion moves the whole run.  The same
// queue here is three pieces that are never copied into one another:
//   * the ACTIVE RUN in shared memory (kActiveCap entries per lane, element-major / lane-minor).  An entry holds the
//     threshold as it was before the cut — (bottom, left, right), the top being the run's — plus the x of the cut,
//     from which both halves follow by selection: the upper part the band's sections are measured against, and the
//     lower part;
//   * the REMAINDERS: when the run ends its lower parts are written over it in place, in the order the reference's
//     insertions would have left them (they all start at the cut: ordered by x there, then slope; a later one goes
//     before an earlier one it ties with);
//   * the rest of the sorted queue, read in order from the threshold store and never written.
// The next run is the remainders merged with the store's thresholds that start at the same y (a remainder goes before
// a stored threshold it ties with, as an insertion from the front would put it).  Numbers that break the order — a
// NaN coordinate, produced or inherited — send the thread to the replay kernel, which follows the reference's
// insertion sequence literally; so does a run of more than kActiveCap thresholds.
#ifndef GUDNI_ACTIVE_CAP
#define GUDNI_ACTIVE_CAP 12
#endif
constexpr int kActiveCap = GUDNI_ACTIVE_CAP;

template <int CAP>
struct SliceScratchT {
    float4 aThr[CAP * 32];           // (bottom, left, right, x of the cut)
    uint32_t aHdr[CAP * 32];
    unsigned int slabNext, slabEnd;  // the warp's private range of pool chunks
    unsigned int pad[2];
};
using SliceScratch = SliceScratchT<kActiveCap>;
// A run that outgrows kActiveCap is not lost to the lane-private replay at once: the slice kernel flags the thread and
// lists its unit, and raster_slice_wide_kernel — one warp per CTA, a run of up to kActiveCapWide entries per lane —
// slices the flagged threads of the listed units again before the colour passes start.  Only a run beyond that (or a
// NaN) is replayed.
#ifndef GUDNI_ACTIVE_CAP_WIDE
#define GUDNI_ACTIVE_CAP_WIDE 72
#endif
constexpr int kActiveCapWide = GUDNI_ACTIVE_CAP_WIDE;   // 72 x 32 x 20 B = 46 KB of static shared memory

constexpr int kBadOrder = 1, kBadCapacity = 2;
struct ActiveRun {
    float4* thr;            // this lane's column of SliceScratch::aThr: entry e at thr[e * 32]
    uint32_t* hdr;
    const float4* sThr;     // the thread's sorted queue in the store
    const uint32_t* sHdr;
    int sNext, sCount;      // stored thresholds not consumed yet: [sNext, sCount); element sNext is in `head`
    float4 head;            // (top, bottom, left, right)
    uint32_t headH;
    float runTop, cutY;     // the active run's common top; where it was cut (the uppers' bottom)
    int front;              // first entry of the run still active (zero-height ones in front are dropped, K.cl:1794-1799)
    int rem;                // remainders waiting in thr[0 .. rem), all starting at remTop
    float remTop;
    int bad;                // kBadOrder: a NaN, replay the thread; kBadCapacity: a run longer than the scratch holds
    __device__ __forceinline__ void attach(const float4* t, const uint32_t* h, unsigned int offset, int count) {
        sThr = t + offset; sHdr = h + offset; sNext = 0; sCount = count;
        front = 0; rem = 0; remTop = FLT_MAX; runTop = 0.0f; cutY = 0.0f; bad = 0;
        head = make_float4(FLT_MAX, FLT_MAX, 0.f, 0.f); headH = 0u;
        if (count > 0) { head = sThr[0]; headH = sHdr[0]; }
    }
    __device__ __forceinline__ bool haveHead() const { return sNext < sCount; }
    __device__ __forceinline__ void advance() {
        sNext++;
        if (sNext < sCount) { head = sThr[sNext]; headH = sHdr[sNext]; }
        else head.x = FLT_MAX;
    }
    // the upper part of entry e: what the reference's array holds at that index while the run is active
    __device__ __forceinline__ void upper(int e, Thr& t, uint32_t& h) const {
        const float4 v = thr[e * 32];
        const uint32_t eh = hdr[e * 32];
        const bool cut = (runTop < cutY) && (cutY < v.x);
        const bool pos = hPositive(eh);
        t.top = runTop;
        t.bottom = cut ? cutY : v.x;
        t.left = (cut && !pos) ? v.w : v.y;
        t.right = (cut && pos) ? v.w : v.z;
        h = (cut && !pos) ? (eh & ~kPersistBit) : eh;
    }
    __device__ __forceinline__ uint32_t upperHeader(int e) const {
        const float bottom = thr[e * 32].x;
        const uint32_t eh = hdr[e * 32];
        const bool cut = (runTop < cutY) && (cutY < bottom);
        return (cut && !hPositive(eh)) ? (eh & ~kPersistBit) : eh;
    }
    __device__ __forceinline__ float upperBottom(int e) const {
        const float bottom = thr[e * 32].x;
        return ((runTop < cutY) && (cutY < bottom)) ? cutY : bottom;
    }
};

// Warp-converged: make sure the warp's slab holds at least `want` chunks (a lane that still runs dry inside a
// round falls back to the global cursor).  What is left of the old slab is abandoned: address space, not traffic.
template <class WS>
__device__ __forceinline__ void ensureSlab(const FrameParams& P, WS& W, unsigned int want) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane == 0) {
        const unsigned int next = W.slabNext, end = W.slabEnd;
        if (next >= end || end - next < want) {
            const unsigned int b = (unsigned int)atomicAdd(&P.counters[kCntStreamCursor], (unsigned long long)kSlabChunks);
            W.slabNext = b;
            W.slabEnd = b + kSlabChunks;
        }
    }
    __syncwarp();
}

struct StreamWriter {
    uint2* chunk;
    int pos;
    bool failed;           // the pool ran out
    unsigned int first;    // first chunk of the stream
    template <class WS>
    __device__ __forceinline__ unsigned int alloc(const FrameParams& P, WS& W) {
        unsigned int c = atomicAdd(&W.slabNext, 1u);
        if (c >= W.slabEnd) c = (unsigned int)atomicAdd(&P.counters[kCntStreamCursor], 1ull);
        return c;
    }
    template <class WS>
    __device__ __forceinline__ void open(const FrameParams& P, WS& W) {
        failed = false;
        pos = 0;
        first = alloc(P, W);
        if (first >= P.streamCapChunks) { failed = true; first = 0u; }
        chunk = P.streamPool + (size_t)first * kChunkRecs;
    }
    template <class WS>
    __device__ __forceinline__ void put(const FrameParams& P, WS& W, uint32_t tag, uint32_t payload) {
        if (failed) return;
        if (pos == kChunkRecs - 1) {   // the last slot of a chunk links to the next one
            const unsigned int c = alloc(P, W);
            if (c >= P.streamCapChunks) { failed = true; return; }
            chunk[pos] = make_uint2(kRecLink, c);
            chunk = P.streamPool + (size_t)c * kChunkRecs;
            pos = 0;
        }
        chunk[pos++] = make_uint2(tag, payload);
    }
    __device__ __forceinline__ void close() {
        if (!failed) chunk[pos] = make_uint2(kRecLink, kStreamEnd);   // pos <= kChunkRecs - 1 always
    }
};

// The run has ended (its bottom is the band's): its lower parts replace it, in queue order (see ActiveRun).
__device__ __forceinline__ void runToRemainders(ActiveRun& q, int numActive) {
    const float top = q.runTop, cutY = q.cutY;
    int m = 0;
    for (int i = 0; i < numActive; i++) {
        const float4 v = q.thr[(q.front + i) * 32];
        const uint32_t h = q.hdr[(q.front + i) * 32];
        if (!((top < cutY) && (cutY < v.x))) continue;                // not cut: it ends with the run
        // splitThreshold / divideThreshold, K.cl:928-979: the lower part
        const bool pos = hPositive(h);
        const Thr lower = pos ? Thr{cutY, v.x, v.w, v.z} : Thr{cutY, v.x, v.y, v.w};
        const uint32_t lh = pos ? (h & ~kPersistBit) : h;
        if (!tKeep(lh, lower)) continue;
        // insertThreshold (K.cl:1105-1124) walks from the front past everything the new one is strictly below and
        // stops; in a sorted list that is: behind it stay exactly the entries it is not strictly below
        int j = m;
        while (j > 0) {
            const float4 u = q.thr[(j - 1) * 32];
            const uint32_t uh = q.hdr[(j - 1) * 32];
            if (isBelow(lh, lower, uh, Thr{cutY, u.x, u.y, u.z})) break;
            q.thr[j * 32] = u;
            q.hdr[j * 32] = uh;
            j--;
        }
        q.thr[j * 32] = make_float4(lower.bottom, lower.left, lower.right, 0.0f);
        q.hdr[j * 32] = lh;
        m++;
    }
    q.rem = m;
    q.remTop = cutY;
    q.front = 0;
}

// splitNext = countActive + nextSlicePoint + sliceActive (K.cl:1007-1077): the thresholds that start at the queue's
// smallest top become the run and are cut at the next event.  Returns the slice point; numActive = 0 and q.bad set
// if the run does not fit.
template <int CAP>
__device__ __forceinline__ float formRun(ActiveRun& q, int& numActive) {
    const float top = fminf(q.rem > 0 ? q.remTop : FLT_MAX, q.head.x);
    int n = q.rem;
    q.rem = 0;
    q.runTop = top;
    q.front = 0;
    // stored thresholds with the same top join the remainders, each behind every entry that is not strictly below it
    while (q.haveHead() && !(q.head.x > top)) {
        if (n == CAP) { q.bad = kBadCapacity; numActive = 0; return top; }
        const Thr t{q.head.x, q.head.y, q.head.z, q.head.w};
        const uint32_t h = q.headH;
        int j = n;
        while (j > 0) {
            const float4 u = q.thr[(j - 1) * 32];
            const uint32_t uh = q.hdr[(j - 1) * 32];
            if (!isBelow(uh, Thr{top, u.x, u.y, u.z}, h, t)) break;
            q.thr[j * 32] = u;
            q.hdr[j * 32] = uh;
            j--;
        }
        q.thr[j * 32] = make_float4(t.bottom, t.left, t.right, 0.0f);
        q.hdr[j * 32] = h;
        n++;
        q.advance();
    }
    // countActive + nextSlicePoint: the next top below the run, the run's smallest bottom
    float slicePoint = q.head.x;    // FLT_MAX when the store is exhausted
    for (int i = 0; i < n; i++) {
        const float bottom = q.thr[i * 32].x;
        if (top < bottom) slicePoint = fminf(slicePoint, bottom);
    }
    q.cutY = slicePoint;
    // sliceActive: where each threshold that reaches below the slice point crosses it
    bool nan = false;
    for (int i = 0; i < n; i++) {
        const float4 v = q.thr[i * 32];
        if ((top < slicePoint) && (slicePoint < v.x)) {
            const float splitX = intersectX(q.hdr[i * 32], Thr{top, v.x, v.y, v.z}, slicePoint);
            q.thr[i * 32].w = splitX;
            nan = nan || (splitX != splitX);
        }
    }
    if (nan) { q.bad = kBadOrder; numActive = 0; return top; }
    numActive = n;
    return slicePoint;
}

// verticalAdvance, K.cl:1744-1824, without the shape stack: what it does to the stack (toggling the persistent
// thresholds that cross the band border) goes into the stream as FLIP records; the un-toggling of the band that
// ended (K.cl:1756-1759) is the resolve kernel resetting `cur` at the next first-of-band section.
template <int CAP, class WS>
__device__ __forceinline__ void sliceVertical(const FrameParams& P, WS& W, StreamWriter& out, ActiveRun& q, SweepState& st,
                                              float floatHeight) {
    st.gapTop = 0.0f;
    const float nextBreak = fminf(floatHeight, st.pixelY);
    const float activeBottom = st.numActive > 0 ? q.upperBottom(q.front) : FLT_MAX;
    if (st.numActive > 0 && activeBottom == st.ey) {   // the active run ends here: its persistent bottoms toggle, then it is popped
        for (int i = 0; i < st.numActive; i++) {
            const uint32_t h = q.upperHeader(q.front + i);
            if (hPersistBottom(h)) out.put(P, W, kRecFlip | (h & kRecBitMask), 0u);
        }
        runToRemainders(q, st.numActive);
        st.numActive = 0;
    }
    float nextBottom;
    if (st.numActive > 0) {
        nextBottom = fminf(activeBottom, nextBreak);
    } else {
        const float nextTop = fminf(q.rem > 0 ? q.remTop : FLT_MAX, q.head.x);
        if (nextTop > st.ey) {
            nextBottom = fminf(nextBreak, nextTop);
            st.gapTop = nextTop;   // nothing crosses the column above this y
        } else {
            nextBottom = fminf(nextBreak, formRun<CAP>(q, st.numActive));
            while (st.numActive > 0) {   // zero-height thresholds in front of the run: K.cl:1794-1799
                if (q.runTop != q.upperBottom(q.front)) break;
                const uint32_t h = q.upperHeader(q.front);
                if (hPersistTop(h)) out.put(P, W, kRecFlip | (h & kRecBitMask), 0u);
                q.front++;
                st.numActive--;
            }
            // K.cl:1803-1808 tests tTop(threshold i) > RENDERSTART per active; the actives share their top
            if (q.runTop > 0.0f) {
                for (int i = 0; i < st.numActive; i++) {
                    const uint32_t h = q.upperHeader(q.front + i);
                    if (hPersistTop(h)) out.put(P, W, kRecFlip | (h & kRecBitMask), 0u);
                }
            }
        }
    }
    st.sy = st.ey;
    st.ey = nextBottom;
    st.sx = st.ex = 0.0f;
    st.cur = 0;
}

// replayed by raster_spill_kernel against an HBM queue of MAXTHRESHOLDS entries
__device__ __forceinline__ void registerSpill(const FrameParams& P, int tileIndex, int column) {
    const unsigned long long slot = atomicAdd(&P.counters[kCntSpilled], 1ull);
    if (slot < (unsigned long long)P.spillCapacity)
        P.spillList[slot] = ((unsigned long long)tileIndex << 32) | (unsigned long long)column;
}


// ---- slice kernel body -------------------------------------------------------------------------------------------
// One warp, one (tile, 32-column group) of a dense tile.  Returns per lane 1 if the thread has to be replayed
// (a run outgrew the on-chip capacity, a NaN turned up, or the stream pool ran out: `exhausted`).
// `rec`: the lane's thread record, null for a lane that has no column-thread in this unit (units narrower than a warp)
// `wide` (per lane): the thread only ran out of run capacity — raster_slice_wide_kernel takes it (see kActiveCapWide).
template <int CAP, class WS>
__device__ __forceinline__ int sliceWarp(const FrameParams& P, WS& W, ActiveRun& q, const gudni_tile& tile,
                                         ThreadRec* recp, int column, bool& exhausted, bool& wide) {
    const unsigned full = 0xffffffffu;
    const ThreadGeom g = threadGeom(P, tile, column);
    const unsigned int recOffset = recp ? recp->offset : 0u, recCount = recp ? recp->count : kRecInactive;
    // a picture's colour depends on the row, so runs of untouched rows are only folded in tiles without pictures
    // (the generate kernel left the tile's answer in every thread record)
    const bool foldable = !(recp && (recp->pad1 & kRecTilePictures));
    const bool mine = recCount != kRecInactive;
    const float floatHeight = (float)g.intHeight;
    SweepState st;
    st.alive = false;
    ensureSlab(P, W, 32u + kSlabLow);
    StreamWriter out;
    out.failed = false; out.pos = 0; out.first = 0u; out.chunk = P.streamPool;
    q.attach(P.thrStore, P.hdrStore, recOffset, mine ? (int)recCount : 0);
    bool spilled = false;
    if (mine) {
        st.init(floatHeight);
        out.open(P, W);
        if (recp->pad1 & kRecUnordered) { spilled = true; st.alive = false; }   // a NaN among its thresholds (raster_sort_kernel)
    }
    bool first = false;    // the next section record opens a band
    int blankRun = 1;      // pixels the band being swept stands for
    for (;;) {
        if (!__any_sync(full, st.alive)) break;
        ensureSlab(P, W, kSlabLow);
        // ---- band boundary: close the pixel, open the next band (see sweepWarp (A) for the blank-run shortcut)
        if (st.alive && st.ex == 1.0f) {
            if (st.ey >= st.pixelY) {   // calculatePixel's loop condition failed: the pixel is complete
                out.put(P, W, kRecPixelEnd | (uint32_t)blankRun, 0u);
                nextPixel(st, floatHeight);
                if (blankRun > 1) {
                    const float skipped = (float)(blankRun - 1);
                    st.sy += skipped;
                    st.ey = st.sy;
                    st.pixelY += skipped;
                    st.row += blankRun - 1;
                    st.alive = st.pixelY <= floatHeight;
                }
            }
            if (st.alive) {
                sliceVertical<CAP>(P, W, out, q, st, floatHeight);
                if (q.bad) { spilled = true; st.alive = false; }
                first = true;
                blankRun = 1;
                if (foldable && st.sy == st.pixelY - 1.0f) {
                    const float limit = fminf(st.gapTop, floatHeight);
                    if (limit >= st.pixelY + 1.0f) blankRun = min((int)(limit - st.pixelY) + 1, kMaxPixelRun);
                }
            }
        }
        // ---- the sections of the band: horizontalAdvance (K.cl:1826-1851) + the bookkeeping of calculatePixel
        while (st.alive) {
            float nextX = 1.0f;
            uint32_t bit = kRecNoBit;
            if (st.cur < st.numActive) {
                Thr t;
                uint32_t h;
                q.upper(q.front + st.cur, t, h);
                const float yMid = st.sy + ((st.ey - st.sy) * 0.5f);
                const float x = intersectX(h, t, yMid);
                nextX = (x >= 1.0f) ? 0.0f : fmaxf(0.0f, x);
                bit = h & kRecBitMask;
            }
            st.sx = st.ex;
            st.ex = nextX;
            const float area = (st.ex - st.sx) * (st.ey - st.sy);
            st.cur++;
            // a zero-area section adds colour * 0 = 0 to every accumulator: it only matters for its toggle
            if (area != 0.0f || bit != kRecNoBit || first) {
                out.put(P, W, kRecSection | (first ? kRecFirst : 0u) | bit, __float_as_uint(area));
                first = false;
            }
            if (st.ex == 1.0f) break;
        }
        if (out.failed) st.alive = false;
    }
    exhausted = false;
    wide = false;
    if (mine) {
        out.close();
        if (out.failed) { spilled = true; exhausted = true; }
        recp->chunk = out.first;
        wide = spilled && q.bad == kBadCapacity && !out.failed && CAP < kActiveCapWide;
        if (wide) { recp->pad1 |= kRecWide; return 0; }   // its thresholds stay where they are: sliced again with a longer run
        if (spilled) recp->count = kRecInactive;   // the later passes skip it; the replay renders the whole thread
    }
    return spilled ? 1 : 0;
}

// Reads a stream two records (16 bytes) at a time: chunks are 128-byte aligned, records 8 bytes.
struct StreamReader {
    const uint2* chunk;
    int pos;
    uint4 pair;   // records (pos & ~1) and (pos | 1)
    __device__ __forceinline__ void open(const uint2* pool, unsigned int first) {
        chunk = pool + (size_t)first * kChunkRecs;
        pos = 0;
        pair = *reinterpret_cast<const uint4*>(chunk);
    }
    __device__ __forceinline__ uint2 get() const { return (pos & 1) ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y); }
    __device__ __forceinline__ void next() {
        pos++;
        if (!(pos & 1)) pair = *reinterpret_cast<const uint4*>(chunk + pos);
    }
    __device__ __forceinline__ void jump(const uint2* pool, unsigned int c) { open(pool, c); }
};

// ---- colours: resolve -> composite -> accumulate ------------------------------------------------------------
// The section streams are replayed twice.  raster_resolve_kernel rebuilds the shape stack of every section from
// the toggles and gives every *distinct* stack a number (a per-warp cache in shared memory keyed by the 128-bit
// stack does the deduplication: neighbouring columns and rows keep meeting the same few stacks); the stack goes
// into a global table, its number into the section's record.  raster_composite_kernel then composites every
// stack of the table once — one stack per lane, every lane of every warp busy with the same loop, operation for
// operation determineColor (K.cl:1447-1513).  raster_accumulate_kernel walks the streams again and adds
// colour * area into the pixel's accumulators in section order (K.cl:1904) and stores the pixels.
// Stack numbers are handed out in slabs of kRefSlab owned by one warp while it works on one tile, so all
// stacks of a slab share the tile's substance table and the composite kernel builds it once per slab.
// Tiles with picture substances (the colour depends on the pixel, K.cl:1420-1441) take raster_picture_kernel
// instead: one pass, every lane composites its own sections.
#ifndef GUDNI_COLOR_LINES
#define GUDNI_COLOR_LINES 256
#endif
constexpr int kColorLines = GUDNI_COLOR_LINES;     // per-warp stack cache of the resolve kernel (power of two)
#ifndef GUDNI_REF_SLAB
#define GUDNI_REF_SLAB 128
#endif
constexpr unsigned int kRefSlab = GUDNI_REF_SLAB;  // stack numbers a warp draws at a time
constexpr uint32_t kRefNone = 0xFFFFFFFFu;
constexpr uint32_t kRefMask = 0x3FFFFFFFu;         // a SEC record's first word once resolved (kind bits stay 00)

struct ResolveScratch {
    ulonglong2 key[kColorLines];     // stack (lo, hi)
    uint32_t ref[kColorLines];       // its number; kRefNone: empty line
    uint8_t claim[kColorLines];
};

__device__ __forceinline__ void colorLines(uint64_t hi, uint64_t lo, uint32_t& line1, uint32_t& line2) {
    uint32_t t = (uint32_t)lo ^ ((uint32_t)(lo >> 32) * 0x85EBCA6Bu) ^ ((uint32_t)hi * 0xC2B2AE35u) ^
                 ((uint32_t)(hi >> 32) * 0x27D4EB2Fu);
    t ^= t >> 15;
    t *= 0x2C1B3C6Du;
    line1 = (t >> 20) & (uint32_t)(kColorLines - 1);
    line2 = line1 ^ (((t >> 9) & (uint32_t)(kColorLines - 1)) | 1u);
}

// does the unit's tile list a picture substance?  The generate kernel left the answer in every thread record of
// the tile (kRecTilePictures); warp-uniform.
__device__ __forceinline__ bool unitHasPictures(const ThreadRec* recp) {
    return __any_sync(0xffffffffu, recp != nullptr && (recp->pad1 & kRecTilePictures) != 0u);
}

// Warp-uniform state of the resolve kernel's stack numbering.
struct RefSlab {
    unsigned int base;    // first number of the warp's current slab; kRefNone: none drawn yet
    unsigned int used;
    int tileIndex;        // the tile the slab's stacks belong to
};
__device__ __forceinline__ void closeSlab(const FrameParams& P, RefSlab& slab) {
    if ((threadIdx.x & 31) == 0 && slab.base != kRefNone)
        P.refSlabs[slab.base / kRefSlab] = make_uint2((unsigned int)slab.tileIndex, slab.used);
    slab.base = kRefNone;
    slab.used = 0;
}

// One warp, one (tile, 32-column group) of a dense tile without pictures.  Returns per lane 1 if the stack table
// ran out under this thread (it is handed to the replay).
__device__ __forceinline__ int resolveWarp(const FrameParams& P, ResolveScratch& W, RefSlab& slab, int tileIndex, ThreadRec* recp) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (slab.tileIndex != tileIndex) {   // numbers are only shared inside a tile: new slab, empty cache
        closeSlab(P, slab);
        slab.tileIndex = tileIndex;
        for (int i = lane; i < kColorLines; i += 32) W.ref[i] = kRefNone;
    }
    ThreadRec rec{};
    rec.count = kRecInactive;
    if (recp) rec = *recp;
    ShapeStack base{rec.lo, rec.hi}, cur{rec.lo, rec.hi};
    bool done = rec.count == kRecInactive;
    bool failed = false;
    StreamReader in;
    in.chunk = P.streamPool; in.pos = 0; in.pair = make_uint4(0u, 0u, 0u, 0u);
    if (!done) in.open(P.streamPool, rec.chunk);
    __syncwarp();
    for (;;) {
        if (!__any_sync(full, !done)) break;
        const uint2 r = done ? make_uint2(kRecFlip | kRecNoBit, 0u) : in.get();
        const uint32_t kind = r.x & kRecKindMask;
        const uint32_t bit = r.x & kRecBitMask;
        bool need = false;
        if (!done) {
            if (kind == kRecSection) {
                if (r.x & kRecFirst) cur = base;   // K.cl:1756-1759: the band that ended is un-toggled
                need = __uint_as_float(r.y) != 0.0f;
            } else if (kind == kRecFlip) {
                if (bit != kRecNoBit) base.flip(bit);
            }
        }
        // ---- the stack's number: from the cache, or a new one ------------------------------------------
        uint32_t ref = kRefNone, line = 0;
        bool claimed = false;
        if (need) {
            uint32_t line2;
            colorLines(cur.hi, cur.lo, line, line2);
            const uint32_t r1 = W.ref[line], r2 = W.ref[line2];
            const ulonglong2 k1 = W.key[line];
            if (r1 != kRefNone && k1.x == cur.lo && k1.y == cur.hi) {
                ref = r1;
            } else {
                const ulonglong2 k2 = W.key[line2];
                if (r2 != kRefNone && k2.x == cur.lo && k2.y == cur.hi) {
                    ref = r2;
                } else {   // a new stack: an empty line if there is one, else the first line's stack is evicted
                    if (r1 != kRefNone && r2 == kRefNone) line = line2;
                    claimed = true;
                    W.claim[line] = (uint8_t)lane;
                }
            }
        }
        if (__any_sync(full, claimed)) {
            __syncwarp();
            const bool winner = claimed && W.claim[line] == (uint8_t)lane;   // one per line
            const unsigned winners = __ballot_sync(full, winner);
            const unsigned int n = (unsigned int)__popc(winners);
            if (slab.base == kRefNone || slab.used + n > kRefSlab) {
                closeSlab(P, slab);
                unsigned int s = 0;
                if (lane == 0) s = atomicAdd(P.work + kWorkRefSlabs, 1u);
                s = __shfl_sync(full, s, 0);
                if (s >= P.refCapSlabs) {   // the batch's region of the table is full: the warp's remaining threads go to the replay
                    failed = failed || !done;
                    done = true;
                    continue;
                }
                slab.base = (P.refSlabBase + s) * kRefSlab;
            }
            if (winner) {
                ref = slab.base + slab.used + (unsigned int)__popc(winners & ((1u << lane) - 1u));
                P.stackKeys[ref] = make_ulonglong2(cur.lo, cur.hi);
                W.key[line] = make_ulonglong2(cur.lo, cur.hi);
                W.ref[line] = ref;
            }
            slab.used += n;
            __syncwarp();
            if (claimed && !winner) {   // neighbouring columns tend to meet a new stack in the same round: share the winner's number
                const ulonglong2 k = W.key[line];
                if (k.x == cur.lo && k.y == cur.hi) ref = W.ref[line];   // (else: look again next round)
            }
        }
        if (!done && (!need || ref != kRefNone)) {
            if (need) const_cast<uint2*>(in.chunk)[in.pos].x = ref;   // kind bits 00: still a SEC record
            if (kind == kRecSection && bit != kRecNoBit) cur.flip(bit);   // K.cl:1907-1910
            if (kind == kRecLink) {
                if (r.y == kStreamEnd) done = true;
                else in.jump(P.streamPool, r.y);
            } else {
                in.next();
            }
        }
    }
    if (failed && recp) recp->count = kRecInactive;
    return failed ? 1 : 0;
}

// determineColor (K.cl:1447-1513) against a table of premultiplied colours and meta words indexed by stack bit
struct TileTable {
    float4 premul[kWarpTableCap];        // 2,048 B
    uint32_t meta[kWarpTableCap];        //   512 B
};
// fills the table (warp-cooperative); returns through the flags what kind of substances the tile lists
// `plain` (may be null): every shape of the tile blends (add / continue tag) and their substance ids strictly rise or strictly
// fall down the list — the usual case, a substance per shape handed out in scene order — so no two layers of a stack share a
// substance and determineColor's "same substance as the layer above" test (K.cl:1476-1490) can never fire.
__device__ __forceinline__ void buildTileTable(const FrameParams& P, TileTable& T, const gudni_tile& tile, bool& anyPicture, bool& anyWild,
                                               bool* plain = nullptr) {
    const int lane = threadIdx.x & 31;
    bool pic = false, wild = false, allSet = true;
    for (uint32_t i = lane; i < tile.shape_count; i += 32) {
        const uint32_t meta = tagMeta(__ldg(&P.shapes[tile.shape_start + i].tag));
        T.meta[i] = meta;
        allSet = allSet && (meta & kMetaSet) != 0u;
        pic = pic || (meta & kMetaPicture);
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!(meta & kMetaPicture)) {
            c = __ldg(P.substances + (meta & kMetaIdMask));
            wild = wild || !substanceIsTame(c);
            c = premultiply(c);
        }
        T.premul[i] = c;
    }
    anyPicture = __any_sync(0xffffffffu, pic);
    anyWild = __any_sync(0xffffffffu, wild);
    if (plain) {
        __syncwarp();
        bool rising = true, falling = true;
        for (uint32_t i = lane; i + 1 < tile.shape_count; i += 32) {
            const uint32_t a = T.meta[i] & kMetaIdMask, b = T.meta[i + 1] & kMetaIdMask;
            rising = rising && a < b;
            falling = falling && a > b;
        }
        *plain = __all_sync(0xffffffffu, allSet) && (__all_sync(0xffffffffu, rising) || __all_sync(0xffffffffu, falling));
    }
}

// index of the highest set bit of a non-zero word (one instruction; 31 - __clz() is three)
__device__ __forceinline__ int topBit(uint32_t word) {
#ifdef GUDNI_HOST_EMULATION
    return 31 - __clz((int)word);
#else
    int b;
    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(word));
    return b;
#endif
}
// The tame walk is the composite kernel's inner loop (15.7 M stacks of ~38 layers per S4 frame, 95 % issue-active), so it
// is written without branches inside a layer: the four words of the stack are walked by four copies of one loop, a layer
// that does not blend (same substance as the one above, or not an add / continue tag) computes and discards, and
// composite's `alphaOut > 0` test (K.cl:883) becomes a divisor of 1 — alphaOut is 0 only when the colour so far and the
// layer both have alpha 0, and then every numerator is exactly 0, so the quotients are the 0 the reference returns.
#ifdef GUDNI_TAME_BRANCHY
__device__ __forceinline__ float4 stackColorTame(const TileTable& T, uint64_t hi, uint64_t lo, float4 bgPremul) {
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
        const uint32_t meta = T.meta[bit];
        const uint32_t id = meta & kMetaIdMask;
        if (id != lastId && (meta & kMetaSet)) {
            base = compositeOverPremulT<false>(base, T.premul[bit]);
            if (base.w == 1.0f) return base;
        }
        lastId = id;
    }
}
#else
__device__ __forceinline__ void tameLayer(float& bx, float& by, float& bz, float& bw, const float4 pm, bool take) {
    const float oneMinus = 1.0f - bw;
    const float alphaOut = bw + pm.w * oneMinus;
    const float d = alphaOut > 0.0f ? alphaOut : 1.0f;
    float qx, qy, qz;
    div3<false>((bx * bw) + (pm.x * oneMinus), (by * bw) + (pm.y * oneMinus), (bz * bw) + (pm.z * oneMinus), d, qx, qy, qz);
    bx = take ? qx : bx;
    by = take ? qy : by;
    bz = take ? qz : bz;
    bw = take ? alphaOut : bw;
}
#ifdef GUDNI_TAME_UNROLLED
__device__ __forceinline__ float4 stackColorTame(const TileTable& T, uint64_t hi, uint64_t lo, float4 bgPremul) {
    float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f;
    uint32_t lastId = 0xFFFFFFFFu;
#pragma unroll
    for (int w = 3; w >= 0; w--) {
        uint32_t word = (w == 3) ? (uint32_t)(hi >> 32) : (w == 2) ? (uint32_t)hi : (w == 1) ? (uint32_t)(lo >> 32) : (uint32_t)lo;
        const uint32_t* meta = T.meta + 32 * w;
        const float4* premul = T.premul + 32 * w;
        while (word != 0u) {
            const int b = 31 - __clz((int)word);
            word ^= (1u << b);
            const uint32_t m = meta[b];
            const uint32_t id = m & kMetaIdMask;
            const bool take = id != lastId && (m & kMetaSet) != 0u;
            lastId = id;
            tameLayer(bx, by, bz, bw, premul[b], take);
            if (bw == 1.0f) return make_float4(bx, by, bz, bw);   // (only a layer that was taken can have made it 1)
        }
    }
    tameLayer(bx, by, bz, bw, bgPremul, true);
    return make_float4(bx, by, bz, bw);
}
#else
__device__ __forceinline__ float4 stackColorTame(const TileTable& T, uint64_t hi, uint64_t lo, float4 bgPremul) {
    float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f;
    uint32_t lastId = 0xFFFFFFFFu;
    uint32_t word = (uint32_t)(hi >> 32);
    int wordBase = 96;
    // one loop for all four words: the lanes of a warp hold different stacks and stay together as long as any has a layer left
    for (;;) {
        while (word == 0u) {
            if (wordBase == 0) {
                tameLayer(bx, by, bz, bw, bgPremul, true);
                return make_float4(bx, by, bz, bw);
            }
            wordBase -= 32;
            word = (wordBase == 64) ? (uint32_t)hi : (wordBase == 32) ? (uint32_t)(lo >> 32) : (uint32_t)lo;
        }
        const int b = topBit(word);
        word ^= (1u << b);
        const int bit = wordBase + b;
        const uint32_t m = T.meta[bit];
        const uint32_t id = m & kMetaIdMask;
        const bool take = id != lastId && (m & kMetaSet) != 0u;
        lastId = id;
        tameLayer(bx, by, bz, bw, T.premul[bit], take);
        if (bw == 1.0f) return make_float4(bx, by, bz, bw);   // (only a layer that was taken can have made it 1)
    }
}
#endif
#endif
// ... and for a plain tile (buildTileTable): every layer blends, nothing to remember from the layer above
__device__ __forceinline__ float4 stackColorPlain(const TileTable& T, uint64_t hi, uint64_t lo, float4 bgPremul) {
    float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f;
    uint32_t word = (uint32_t)(hi >> 32);
    int wordBase = 96;
    const float4* premul = T.premul + 96;   // the word's 32 entries
    for (;;) {
        while (word == 0u) {
            if (wordBase == 0) {
                tameLayer(bx, by, bz, bw, bgPremul, true);
                return make_float4(bx, by, bz, bw);
            }
            wordBase -= 32;
            premul -= 32;
            word = (wordBase == 64) ? (uint32_t)hi : (wordBase == 32) ? (uint32_t)(lo >> 32) : (uint32_t)lo;
        }
        const int b = topBit(word);
        word ^= (1u << b);
        tameLayer(bx, by, bz, bw, premul[b], true);
        if (bw == 1.0f) return make_float4(bx, by, bz, bw);
    }
}
__device__ __forceinline__ float4 stackColorAny(const FrameParams& P, const TileTable& T, uint64_t hi, uint64_t lo, float4 bgPremul,
                                                int absX, int absY) {
    float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t lastId = 0xFFFFFFFFu;
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
        const uint32_t meta = T.meta[bit];
        const uint32_t id = meta & kMetaIdMask;
        if (id != lastId && (meta & kMetaSet)) {
            float4 pm = T.premul[bit];
            if (meta & kMetaPicture) pm = premultiply(readPicture(P, id, absX, absY));
            base = compositeOverPremul(base, pm);
            if (base.w == 1.0f) return base;
        }
        lastId = id;
    }
}

// One warp, one slab of the stack table.
// `mode` (warp-uniform, kept with the table): kWalkAny, kWalkTame or kWalkPlain
enum { kWalkAny = 0, kWalkTame = 1, kWalkPlain = 2 };
__device__ __forceinline__ void compositeSlab(const FrameParams& P, TileTable& T, int& tableTile, int& mode, unsigned int s) {
    const int lane = threadIdx.x & 31;
    const uint2 info = P.refSlabs[s];
    if ((int)info.x != tableTile) {
        __syncwarp();
        bool anyPicture, anyWild, plain;
        buildTileTable(P, T, P.tiles[info.x], anyPicture, anyWild, &plain);
        const bool tame = !anyWild && substanceIsTame(P.background);
        mode = tame ? (plain ? kWalkPlain : kWalkTame) : kWalkAny;
#ifdef GUDNI_NO_PLAIN_WALK
        if (mode == kWalkPlain) mode = kWalkTame;
#endif
        tableTile = (int)info.x;
        __syncwarp();
    }
    const float4 bgPremul = premultiply(P.background);
    for (unsigned int i = lane; i < info.y; i += 32) {
        const unsigned int ref = s * kRefSlab + i;
        const ulonglong2 key = P.stackKeys[ref];
        float4 c;
        if (mode == kWalkPlain) c = stackColorPlain(T, key.y, key.x, bgPremul);
        else if (mode == kWalkTame) c = stackColorTame(T, key.y, key.x, bgPremul);
        else c = stackColorAny(P, T, key.y, key.x, bgPremul, 0, 0);
        P.stackColors[ref] = c;
    }
}

// writePixelGlobal (K.cl:1853-1862) for a finished pixel and the n - 1 untouched rows below it that equal it
__device__ __forceinline__ void storePixels(const FrameParams& P, uint32_t* outp, int& wrow, int rep, float accR, float accG,
                                            float accB, float accArea) {
    const uint32_t word = pixelWord(accR, accG, accB, accArea);
    for (int k = 0; k < rep; k++) outp[(size_t)(wrow + k) * P.width] = word;
    wrow += rep;
}

// One warp, one (tile, 32-column group) of a dense tile without pictures: the resolved streams again, now with
// the colours at hand.
//
// Pixel stores.  A lane finishes the pixels of its column at its own pace (a column the shapes leave alone is a
// few PEND records, each standing for a run of rows; its neighbour may cross an outline and need dozens of
// sections for the same rows), and 32 lanes storing 4 bytes each to 32 different rows is 32 partial sectors.  So
// finished pixels first go into a window of kRowWindow rows x 32 columns in shared memory — every lane only ever
// touches its own column of it — and a row leaves for memory when every lane has produced it: one 128-byte
// store per row.  A lane that gets kRowWindow rows ahead of the slowest one waits.  (Letting every lane run on until
// its window is full between two flushes measured slower — S5 13.5 -> 14.7 ms: one record per lane per round keeps
// the lanes of neighbouring columns on the same branch.)
#ifndef GUDNI_ROW_WINDOW
#define GUDNI_ROW_WINDOW 16
#endif
constexpr int kRowWindow = GUDNI_ROW_WINDOW;   // power of two
struct AccumScratch {
    uint32_t rows[kRowWindow][32];
};
__device__ __forceinline__ void accumulateWarp(const FrameParams& P, AccumScratch& W, const gudni_tile& tile, const ThreadRec* recp, int column) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const ThreadGeom g = threadGeom(P, tile, column);
    ThreadRec rec{};
    rec.count = kRecInactive;
    if (recp) rec = *recp;
    uint32_t* outp = P.out + (size_t)(g.originY - P.rowOrigin) * P.width + g.originX;   // only dereferenced when active
    const bool mine = rec.count != kRecInactive;
    bool done = !mine;
    StreamReader in;
    in.chunk = P.streamPool; in.pos = 0; in.pair = make_uint4(0u, 0u, 0u, 0u);
    if (mine) in.open(P.streamPool, rec.chunk);
    float accR = 0.f, accG = 0.f, accB = 0.f, accArea = 0.f;
    int wrow = 0;             // rows this lane has produced (into the window)
    int flushed = 0;          // rows already stored (warp-uniform)
    uint32_t pendWord = 0u;   // a finished pixel and the rows it still has to fill
    int pendRep = 0;
    for (;;) {
        // (1) every lane adds up the sections of its current pixel — a few instructions per section
        int rep = 0;
        while (!done && pendRep == 0 && rep == 0) {
            const uint2 r = in.get();
            const uint32_t kind = r.x & kRecKindMask;
            if (kind == kRecLink) {
                if (r.y == kStreamEnd) done = true;
                else in.jump(P.streamPool, r.y);
                continue;
            }
            in.next();
            if (kind == kRecSection) {
                const float area = __uint_as_float(r.y);
                if (area != 0.0f) {
                    const float4 c = P.stackColors[r.x & kRefMask];
                    accR += c.x * area;   // K.cl:1904
                    accG += c.y * area;
                    accB += c.z * area;
                    accArea += area;
                }
            } else if (kind == kRecPixelEnd) {
                rep = (int)(r.x & 0xFFFFu);
            }
        }
        __syncwarp();
        // (2) the lanes that finished a pixel convert it together (three divisions: K.cl:1853-1862)
        if (rep) {
            pendWord = pixelWord(accR, accG, accB, accArea);
            pendRep = rep;
            accR = accG = accB = accArea = 0.f;
        }
        // (3) ... and put it into their column of the window, as far as the window reaches
        if (pendRep) {
            const int n = min(pendRep, flushed + kRowWindow - wrow);
            for (int k = 0; k < n; k++) W.rows[(wrow + k) & (kRowWindow - 1)][lane] = pendWord;
            wrow += n;
            pendRep -= n;
        }
        // (4) rows every unfinished lane has produced leave as whole rows
        const bool finished = done && pendRep == 0;
        int lo = finished ? 0x7FFFFFFF : wrow, hi = mine ? wrow : 0;
        for (int d = 16; d > 0; d >>= 1) {
            lo = min(lo, __shfl_xor_sync(full, lo, d));
            hi = max(hi, __shfl_xor_sync(full, hi, d));
        }
        const int upTo = min(lo, hi);
        for (int r = flushed; r < upTo; r++)
            if (mine && r < wrow) outp[(size_t)r * P.width] = W.rows[r & (kRowWindow - 1)][lane];
        flushed = upTo;
        if (lo == 0x7FFFFFFF) break;   // every lane finished, and what they produced has just been stored
    }
}

// One warp, one (tile, 32-column group) of a dense tile WITH pictures: one pass over the unresolved streams,
// every lane composites the sections of its own pixels.
__device__ __forceinline__ void pictureWarp(const FrameParams& P, const TileTable& T, const gudni_tile& tile, const ThreadRec* recp, int column) {
    const unsigned full = 0xffffffffu;
    const ThreadGeom g = threadGeom(P, tile, column);
    ThreadRec rec{};
    rec.count = kRecInactive;
    if (recp) rec = *recp;
    const float4 bgPremul = premultiply(P.background);
    uint32_t* outp = P.out + (size_t)(g.originY - P.rowOrigin) * P.width + g.originX;
    ShapeStack base{rec.lo, rec.hi}, cur{rec.lo, rec.hi};
    const uint2* chunk = P.streamPool + (size_t)rec.chunk * kChunkRecs;
    int pos = 0;
    bool done = rec.count == kRecInactive;
    float accR = 0.f, accG = 0.f, accB = 0.f, accArea = 0.f;
    int wrow = 0;
    for (;;) {
        if (!__any_sync(full, !done)) break;
        // everything up to the next section with an area
        float area = 0.0f;
        uint32_t bit = kRecNoBit;
        while (!done && area == 0.0f) {
            const uint2 r = chunk[pos++];
            const uint32_t kind = r.x & kRecKindMask;
            if (kind == kRecSection) {
                if (r.x & kRecFirst) cur = base;
                area = __uint_as_float(r.y);
                bit = r.x & kRecBitMask;
                if (area == 0.0f && bit != kRecNoBit) cur.flip(bit);
            } else if (kind == kRecFlip) {
                base.flip(r.x & kRecBitMask);
            } else if (kind == kRecPixelEnd) {
                storePixels(P, outp, wrow, (int)(r.x & 0xFFFFu), accR, accG, accB, accArea);
                accR = accG = accB = accArea = 0.f;
            } else if (r.y == kStreamEnd) {
                done = true;
            } else {
                chunk = P.streamPool + (size_t)r.y * kChunkRecs;
                pos = 0;
            }
        }
        __syncwarp();
        if (!done) {
            const float4 c = stackColorAny(P, T, cur.hi, cur.lo, bgPremul, g.originX, g.originY + wrow);
            accR += c.x * area;
            accG += c.y * area;
            accB += c.z * area;
            accArea += area;
            if (bit != kRecNoBit) cur.flip(bit);
        }
    }
}

}  // namespace gudni_dev
