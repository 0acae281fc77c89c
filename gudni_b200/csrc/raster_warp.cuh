// raster_warp.cuh — warp-cooperative rasterization of one (tile, 32-column group).
//
// The reference gives each work-item the whole job of its column: generate, sort, sweep, and for
// every section of the sweep a walk down the shape stack compositing every translucent layer
// (Kernels.cl:1881-1916, 1447-1513).  On deep scenes that walk is > 80 % of the arithmetic, and it
// is massively redundant: the colour of a section is a pure function of the set of shapes present
// (for solid substances), and the sections of a pixel, of the pixel below and of the neighbouring
// columns keep meeting the same few sets.
//
// For "dense" tiles (no more shapes than MAXSHAPE: every tile above the 8-pixel floor) a shape's
// stack bit is its position in the tile's list, so the 128-bit stack is the same key in every lane
// of every warp working on the tile.  Each warp keeps
//   * the tile's substance table in shared memory (premultiplied colour + meta word per bit), and
//   * a direct-mapped colour cache in shared memory keyed by the stack.
// A lane sweeps lane-privately (cheap bookkeeping) and looks every non-empty section up in the
// cache; on a miss it parks.  When every lane is parked or finished, the parked lanes elect one
// owner per cache line (a 4-byte claim word), the owners composite their stack once
// (determineColor, operation for operation as the reference), publish it, and everyone resumes.
// Each lane still adds colour * area into its own accumulators in section order, so pixels are
// bit-identical to the reference; only the number of times a colour is recomputed changes.
#pragma once
#include "raster_device.cuh"

namespace gudni_dev {

constexpr int kWarpTableCap = 128;
constexpr int kQueueCap = 64;          // thresholds per column-thread before the HBM replay takes over
constexpr int kQueueHot = 16;          // of which in shared memory
constexpr int kColorCacheLines = 64;   // direct mapped
constexpr int kSectionsPerRound = 4;   // section records a lane may park per round

struct WarpScratch {
    float4 premul[kWarpTableCap];                    // 2,048 B  tile substance table
    uint32_t meta[kWarpTableCap];                    //   512 B
    ulonglong2 cacheKey[kColorCacheLines];           // 1,024 B  colour cache: stack (lo, hi)
    float4 cacheColor[kColorCacheLines];             // 1,024 B  colour; w < 0 marks an empty line
    uint32_t cacheClaim[kColorCacheLines];           //   256 B  lane that owns the line this pass
    ulonglong2 recKey[32 * kSectionsPerRound];       // 2,048 B  parked sections: stack, then (aliased) colour
    float recArea[32 * kSectionsPerRound];           //   512 B
    float4 qThr[kQueueHot * 32];                     // 8,192 B  hot part of the 32 threshold queues
    uint32_t qHdr[kQueueHot * 32];                   // 2,048 B
};
typedef WarpQueue<kQueueCap, kQueueHot> LaneQueue;

// determineColor (K.cl:1447-1513) for a dense tile: table index = stack bit.
__device__ __forceinline__ float4 denseColor(const FrameParams& P, const WarpScratch& W, uint64_t hi, uint64_t lo,
                                             float4 bgPremul, int absX, int absY) {
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
            if (meta & kMetaPicture) pm = premultiply(readPicture(P, id, absX, absY));
            base = compositeOverPremul(base, pm);
            if (base.w == 1.0f) return base;
        }
        lastId = id;
    }
}

__device__ __forceinline__ uint32_t stackHash(uint64_t hi, uint64_t lo) {
    uint32_t t = (uint32_t)lo ^ ((uint32_t)(lo >> 32) * 0x85EBCA6Bu) ^ ((uint32_t)hi * 0xC2B2AE35u) ^
                 ((uint32_t)(hi >> 32) * 0x27D4EB2Fu);
    t ^= t >> 15;
    t *= 0x2C1B3C6Du;
    return (t >> 20) & (uint32_t)(kColorCacheLines - 1);
}

// One warp, one (tile, 32-column group) of a dense tile.  Returns per lane: 0 = done or inactive,
// 1 = the lane's threshold queue outgrew the on-chip capacity (caller hands it to the spill list).
// `generated` receives the lane's qSlice.sLength after generation (-1 if inactive or spilled early).
__device__ __forceinline__ int rasterWarpDense(const FrameParams& P, WarpScratch& W, LaneQueue& q, const gudni_tile& tile,
                                               int tileIndex, int column, int& generated) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const ThreadGeom g = threadGeom(P, tile, column);
    // ---- tile substance table + empty colour cache (warp-cooperative) -----------------------------
    bool anyPicture = false;
    for (uint32_t i = lane; i < tile.shape_count; i += 32) {
        const uint32_t meta = tagMeta(__ldg(&P.shapes[tile.shape_start + i].tag));
        W.meta[i] = meta;
        anyPicture = anyPicture || (meta & kMetaPicture);
        W.premul[i] = (meta & kMetaPicture) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                            : premultiply(__ldg(P.substances + (meta & kMetaIdMask)));
    }
    // a picture's colour depends on the pixel, so stacks are only a valid key without pictures
    const bool cacheable = !__any_sync(full, anyPicture);
    for (int i = lane; i < kColorCacheLines; i += 32) W.cacheColor[i] = make_float4(0.f, 0.f, 0.f, -1.f);
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
    // ---- sweep, band-synchronous rounds ---------------------------------------------------------------
    // Every round a lane (A) closes the pixel and opens the next band if it stands at a band boundary
    // — the expensive, branchy part, which the lanes therefore reach together — then (B) walks up to
    // kSectionsPerRound sections of the band, parking {stack, area} records; (C) the warp resolves
    // the colours of all parked records through the colour cache, compositing each missing stack
    // once; (D) each lane adds colour * area of its records in section order (K.cl:1904).
    const float4 bgPremul = premultiply(P.background);
    uint32_t* outp = P.out + (size_t)(g.originY - P.rowOrigin) * P.width + g.originX;   // only dereferenced when active
    ulonglong2* myKey = W.recKey + lane * kSectionsPerRound;
    float* myArea = W.recArea + lane * kSectionsPerRound;
    while (__any_sync(full, st.alive)) {
        // ---- (A) band boundary --------------------------------------------------------------------
        if (st.alive && st.ex == 1.0f) {
            if (st.ey >= st.pixelY) {   // calculatePixel's loop condition failed: the pixel is complete
                outp[(size_t)st.row * P.width] = pixelWord(st.accR, st.accG, st.accB, st.accArea);
                st.accR = st.accG = st.accB = st.accArea = 0.f;
                nextPixel(st, floatHeight);
            }
            if (st.alive) {
                sweepVertical(q, stack, st, floatHeight);
                if (q.failed()) { spilled = true; st.alive = false; }
            }
        }
        // ---- (B) sections of the band ---------------------------------------------------------------
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
        // ---- (C) resolve colours --------------------------------------------------------------------
        if (!cacheable) {
            // picture substances: the colour depends on the pixel, every lane composites its own records
            // (all of them lie in the pixel row the lane is sweeping)
            for (int j = 0; j < count; j++) {
                const ulonglong2 key = myKey[j];
                const float4 color = denseColor(P, W, key.y, key.x, bgPremul, g.originX, g.originY + st.row);
                *reinterpret_cast<float4*>(&myKey[j]) = color;
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
                ulonglong2 key = make_ulonglong2(0ull, 0ull);
                uint32_t line = 0;
                bool need = false;
                float4 color = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    key = W.recKey[slot];
                    line = stackHash(key.y, key.x);
                    const float4 c = W.cacheColor[line];
                    const ulonglong2 k = W.cacheKey[line];
                    if (c.w >= 0.f && k.x == key.x && k.y == key.y) color = c;
                    else need = true;
                }
                if (__any_sync(full, need)) {
                    // one owner per cache line composites; lanes holding the same stack share the result
                    if (need) W.cacheClaim[line] = (uint32_t)lane;
                    __syncwarp();
                    const bool owner = need && W.cacheClaim[line] == (uint32_t)lane;
                    if (owner) {
                        color = denseColor(P, W, key.y, key.x, bgPremul, 0, 0);
                        W.cacheKey[line] = key;
                        W.cacheColor[line] = make_float4(color.x, color.y, color.z, 1.f);
                    }
                    __syncwarp();
                    if (need && !owner) {
                        const ulonglong2 k = W.cacheKey[line];
                        if (k.x == key.x && k.y == key.y) color = W.cacheColor[line];
                        else color = denseColor(P, W, key.y, key.x, bgPremul, 0, 0);   // lost the line to another stack
                    }
                }
                if (valid) *reinterpret_cast<float4*>(&W.recKey[slot]) = color;
            }
        }
        __syncwarp();
        // ---- (D) accumulate in section order ----------------------------------------------------------
        for (int j = 0; j < count; j++) {
            const float4 color = *reinterpret_cast<const float4*>(&myKey[j]);
            const float area = myArea[j];
            st.accR += color.x * area;
            st.accG += color.y * area;
            st.accB += color.z * area;
            st.accArea += area;
        }
        __syncwarp();
    }
    return spilled ? 1 : 0;
}

}  // namespace gudni_dev
