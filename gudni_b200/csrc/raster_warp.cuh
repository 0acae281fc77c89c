// raster_warp.cuh — warp-cooperative rasterization of one (tile, 32-column group).
//
// The reference gives each work-item the whole job of its column: generate, sort, sweep, and for
// every section of the sweep a walk down the shape stack compositing every translucent layer
// (Kernels.cl:1881-1916, 1447-1513).  That walk is > 80 % of the arithmetic on deep scenes and its
// length per pixel differs from lane to lane, so a warp that runs it lane-private idles most of its
// lanes.  Here the sweep is split in two:
//
//   emit        each lane runs the (cheap, divergent) sweep bookkeeping and writes up to
//               kSectionsPerRound section records {shape stack, area} into the warp's shared memory;
//   evaluate    the records of all lanes are numbered with a warp prefix sum and dealt out round
//               robin, so every lane composites one record at a time whoever produced it;
//   accumulate  each lane reads back the colours of its own records IN SECTION ORDER and adds
//               colour * area into its pixel accumulators — the same additions in the same order
//               as the reference, hence bit-identical pixels.
//
// Works for "dense" tiles (no more shapes than MAXSHAPE: every tile above the 8-pixel floor), where
// a shape's stack bit is its position in the tile's list and the tile's substance table in shared
// memory is indexed by the bit directly.
#pragma once
#include "raster_device.cuh"

namespace gudni_dev {

constexpr int kSectionsPerRound = 2;
constexpr int kWarpTableCap = 128;
constexpr int kQueueCap = 64;      // thresholds per column-thread before the HBM replay takes over
constexpr int kQueueHot = 10;      // of which in shared memory
constexpr uint32_t kRecPixelEnd = 1u;

struct SectionRec {   // 32 bytes
    uint64_t hi, lo;  // shape stack the section is coloured with; overwritten by its colour (float4)
    float area;
    uint32_t xy;      // absolute pixel x | y << 16 (picture substances only)
    uint32_t flags;
    uint32_t pad;
};

struct WarpScratch {
    float4 premul[kWarpTableCap];                 //  2,048 B  tile substance table
    uint32_t meta[kWarpTableCap];                 //    512 B
    SectionRec rec[32 * kSectionsPerRound];       //  2,048 B  section records / their colours
    float4 qThr[kQueueHot * 32];                  //  8,192 B  hot part of the 32 threshold queues
    uint32_t qHdr[kQueueHot * 32];                //  2,048 B
};
typedef WarpQueue<kQueueCap, kQueueHot> LaneQueue;

// determineColor (K.cl:1447-1513) for a dense tile: table index = stack bit.
__device__ __forceinline__ float4 denseColor(const FrameParams& P, const WarpScratch& W, uint64_t hi, uint64_t lo,
                                             float4 bgPremul, uint32_t xy) {
    float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t lastId = 0xFFFFFFFFu;
    for (;;) {
        int bit;
        if (hi) { const int b = 63 - __clzll((long long)hi); hi ^= (1ull << b); bit = 64 + b; }
        else if (lo) { const int b = 63 - __clzll((long long)lo); lo ^= (1ull << b); bit = b; }
        else return compositeOverPremul(base, bgPremul);
        const uint32_t meta = W.meta[bit];
        const uint32_t id = meta & kMetaIdMask;
        if (id != lastId && (meta & kMetaSet)) {
            float4 pm = W.premul[bit];
            if (meta & kMetaPicture) pm = premultiply(readPicture(P, id, (int)(xy & 0xFFFFu), (int)(xy >> 16)));
            base = compositeOverPremul(base, pm);
            if (base.w == 1.0f) return base;
        }
        lastId = id;
    }
}

// One warp, one (tile, 32-column group) of a dense tile.  Returns per lane: 0 = done or inactive,
// 1 = the lane's threshold queue outgrew the on-chip capacity (caller hands it to the spill list).
// `generated` receives the lane's qSlice.sLength after generation (-1 if inactive or spilled early).
__device__ __forceinline__ int rasterWarpDense(const FrameParams& P, WarpScratch& W, LaneQueue& q, const gudni_tile& tile,
                                               int tileIndex, int column, int& generated) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const ThreadGeom g = threadGeom(P, tile, column);
    // ---- tile substance table (warp-cooperative) -------------------------------------------------
    for (uint32_t i = lane; i < tile.shape_count; i += 32) {
        const uint32_t meta = tagMeta(__ldg(&P.shapes[tile.shape_start + i].tag));
        W.meta[i] = meta;
        W.premul[i] = (meta & kMetaPicture) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                            : premultiply(__ldg(P.substances + (meta & kMetaIdMask)));
    }
    __syncwarp();
    // ---- generate + sort, lane-private -------------------------------------------------------------
    ShapeStack stack{0ull, 0ull};
    SweepState st;
    const float floatHeight = (float)g.intHeight;
    bool spilled = false;
    generated = -1;
    st.alive = false;
    if (g.active) {
        q.init();
        const uint32_t bits = buildThresholds<true>(P, g, q, stack, nullptr);
        if (q.failed()) {
            spilled = true;
        } else {
            generated = q.len;
            const int threadId = P.tileThreadBase[tileIndex] + column;
            if (P.dbgThresholds) P.dbgThresholds[threadId] = q.len;
            if (P.dbgShapeBits) P.dbgShapeBits[threadId] = (int32_t)bits;
            sortQueue(q);
            st.init(floatHeight);
        }
    }
    // ---- sweep in rounds -----------------------------------------------------------------------------
    const float4 bgPremul = premultiply(P.background);
    uint32_t* outp = P.out + (size_t)(g.originY - P.rowOrigin) * P.width + g.originX;   // only dereferenced when active
    SectionRec* myRec = W.rec + lane * kSectionsPerRound;
    int wrow = 0;   // pixels of the slab stored so far (rows complete in order)
    while (__any_sync(full, st.alive)) {
        // ---- emit ---------------------------------------------------------------------------------
        int count = 0;
        while (st.alive && count < kSectionsPerRound) {
            float area;
            uint64_t hi, lo;
            if (sweepStep(q, stack, st, floatHeight, area, hi, lo) == kSweepPixelDone) {
                if (count > 0) {
                    myRec[count - 1].flags = kRecPixelEnd;   // stored when the colour of its last section is back
                } else {                                     // every section of the pixel is already accumulated
                    outp[(size_t)wrow * P.width] = pixelWord(st.accR, st.accG, st.accB, st.accArea);
                    st.accR = st.accG = st.accB = st.accArea = 0.f;
                    wrow++;
                }
                nextPixel(st, floatHeight);
                continue;
            }
            if (q.failed()) { spilled = true; st.alive = false; break; }
            if (area != 0.0f) {   // a zero-area section adds colour * 0 = 0 to every accumulator
                SectionRec r;
                r.hi = hi; r.lo = lo; r.area = area;
                r.xy = (uint32_t)g.originX | ((uint32_t)(g.originY + st.row) << 16);
                r.flags = 0u; r.pad = 0u;
                myRec[count++] = r;
            }
        }
        if (spilled) count = 0;
        // ---- number the records of the warp -------------------------------------------------------
        int incl = count;
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += t;
        }
        const int excl = incl - count;
        const int total = __shfl_sync(full, incl, 31);
        __syncwarp();
        // ---- evaluate: record f belongs to the first lane o with incl(o) > f ----------------------
        for (int f0 = 0; f0 < total; f0 += 32) {
            const int f = f0 + lane;
            int lo_ = 0, hi_ = 31;   // 32 candidates, 5 halvings
            for (int step = 0; step < 5; step++) {
                const int mid = (lo_ + hi_) >> 1;
                const int v = __shfl_sync(full, incl, mid);
                if (v > f) hi_ = mid; else lo_ = mid + 1;
            }
            const int owner = lo_ & 31;
            const int ownerExcl = __shfl_sync(full, excl, owner);
            if (f < total) {
                const int slot = owner * kSectionsPerRound + (f - ownerExcl);
                const SectionRec r = W.rec[slot];
                *reinterpret_cast<float4*>(&W.rec[slot]) = denseColor(P, W, r.hi, r.lo, bgPremul, r.xy);
            }
        }
        __syncwarp();
        // ---- accumulate in section order (K.cl:1904) ----------------------------------------------
        for (int j = 0; j < count; j++) {
            const float4 color = *reinterpret_cast<const float4*>(&myRec[j]);
            const float area = myRec[j].area;
            st.accR += color.x * area;
            st.accG += color.y * area;
            st.accB += color.z * area;
            st.accArea += area;
            if (myRec[j].flags & kRecPixelEnd) {
                outp[(size_t)wrow * P.width] = pixelWord(st.accR, st.accG, st.accB, st.accArea);
                st.accR = st.accG = st.accB = st.accArea = 0.f;
                wrow++;
            }
        }
        __syncwarp();
    }
    return spilled ? 1 : 0;
}

}  // namespace gudni_dev
