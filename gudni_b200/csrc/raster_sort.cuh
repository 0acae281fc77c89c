// raster_sort.cuh — sortThresholds (Kernels.cl:2084-2115) as a warp-cooperative rank sort.
//
// The reference bubble-sorts every column-thread's threshold array in place in global memory (K.cl:1932-1976), one
// work-item per array.  Round 1 and the first half of round 2 did the same thing lane-privately inside the generate
// kernel (insertion sort where the queue was built: four slots in shared memory, the rest in local memory), which is
// a chain of dependent local-memory round trips per element moved — on S5 (256-row column-threads, queues of a
// hundred thresholds) more than half of the generate kernel's samples.
//
// Here the generate kernel packs the queues UNSORTED, in the order they were built, and a warp sorts the 32 queues
// of a (tile, 32-column) unit together: the queues are staged in shared memory, every lane takes elements of the
// flattened list and counts, for its element, the elements of the same queue that sort before it — the element's
// rank — and the element is written straight to its final place in the store.  No element is moved twice, no lane
// waits on another lane's memory, and the comparison count (sum of n^2) is spread over 32 lanes.
//
// The order is the reference's: `isBelow` (thresholdIsBelow, K.cl:1079-1094: top, then x at the top, then inverse
// slope) with ties keeping the order in which the queue was built — what a stable insertion sort over the same
// predicate gives (sortQueue), which is what the bubble sort of K.cl:1962-1976 gives.  For keys without NaN the
// predicate is a strict weak order, so ranks are a permutation.  A queue that holds a NaN is sorted by its own lane
// with the sequential insertion sort instead (the order then depends on the comparison sequence) and its thread is
// flagged (kRecUnordered): the slice kernel hands it to the replay.
#pragma once
#include "raster_warp.cuh"

namespace gudni_dev {

#ifndef GUDNI_SORT_CAP
#define GUDNI_SORT_CAP 256
#endif
constexpr int kSortCap = GUDNI_SORT_CAP;   // thresholds staged at a time; at least kQueueCap so that any one queue fits
static_assert(kSortCap >= kQueueCap, "a whole queue must fit the staging area");

static_assert(kSortCap < 2048, "staged indices are packed into 11 bits");
#ifndef GUDNI_SORT_LONG
#define GUDNI_SORT_LONG 64
#endif
// Ranking costs n^2 comparisons per queue; a queue longer than this is sorted on its own by the whole warp with a
// bitonic network over a permutation in shared memory instead (n log^2 n; the keys stay where they were staged and
// the element index breaks ties, which makes the order total and therefore the same stable order).
constexpr unsigned int kSortLong = GUDNI_SORT_LONG;

struct SortScratch {
    float4 key[kSortCap];       // (top, x at the top, inverse slope, header bits): what isBelow compares, one 16-byte load
    float4 thr[kSortCap];       // (top, bottom, left, right)
    uint32_t info[kSortCap];    // lane | first staged index of the element's queue << 5 | queue length << 16
    uint32_t hdrRaw[kSortCap + 8];   // headers as the bulk copy lands them: from the 16-byte boundary below the batch's first
    unsigned long long mbar;    // transaction barrier the bulk copies complete on
    unsigned int first[33];     // staged index of each lane's first element (exclusive scan of the lengths)
    unsigned int offset[32];    // each lane's place in the store
    unsigned int nanMask;       // lanes whose queue holds a NaN key
};

// ---- staging by bulk copy ---------------------------------------------------------------------------------------
// The queues of a unit lie end to end in the store (the generate kernel packs a warp's queues with one prefix sum), so
// a batch is one contiguous run of float4 thresholds and one of header words: two 1-D bulk copies (TMA,
// cp.async.bulk global -> shared, completion on an mbarrier) issued by one lane, instead of a gather per element.
// The header run starts on a 4-byte boundary; the copy starts at the 16-byte boundary below it.
struct BulkStage {
    uint32_t phase;   // parity of the barrier's current phase (warp-uniform)
};
__device__ __forceinline__ void bulkInit(SortScratch& W, BulkStage& B) {
    B.phase = 0u;
#ifndef GUDNI_HOST_EMULATION
    if ((threadIdx.x & 31) == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&W.mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncwarp();
}
// Whole warp.  Afterwards W.thr[0 .. n) and W.hdrRaw[pad .. pad + n) hold store[first .. first + n); returns pad.
__device__ __forceinline__ unsigned int bulkStage(const FrameParams& P, SortScratch& W, BulkStage& B, unsigned int first, unsigned int n) {
    const unsigned int first4 = first & ~3u, pad = first - first4;
    const unsigned int hdrBytes = ((pad + n) * 4u + 15u) & ~15u;
    __syncwarp();
#ifdef GUDNI_HOST_EMULATION
    if ((threadIdx.x & 31) == 0) {
        for (unsigned int i = 0; i < n; i++) W.thr[i] = P.thrStore[first + i];
        for (unsigned int i = 0; i < hdrBytes / 4u; i++) W.hdrRaw[i] = P.hdrStore[first4 + i];
    }
    __syncwarp();
#else
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&W.mbar);
    if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n * 16u + hdrBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(W.thr)), "l"(P.thrStore + first), "r"(n * 16u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(W.hdrRaw)), "l"(P.hdrStore + first4), "r"(hdrBytes), "r"(bar) : "memory");
    }
    uint32_t landed = 0u;
    while (!landed) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(landed) : "r"(bar), "r"(B.phase) : "memory");
    }
    B.phase ^= 1u;
#endif
    return pad;
}

// the queue as it lies in the store, for the sequential fallback
struct StoreQueue {
    float4* thr;
    uint32_t* hdr;
    int len;
    __device__ __forceinline__ Thr getT(int i) const { const float4 v = thr[i]; return Thr{v.x, v.y, v.z, v.w}; }
    __device__ __forceinline__ uint32_t getH(int i) const { return hdr[i]; }
    __device__ __forceinline__ void set(int i, uint32_t h, const Thr& t) { thr[i] = make_float4(t.top, t.bottom, t.left, t.right); hdr[i] = h; }
};

// isBelow on staged keys: a sorts strictly after b
__device__ __forceinline__ bool keyBelow(const float4& a, const float4& b) {
    return (bool)((int)(a.x > b.x) | ((int)(a.x == b.x) & ((int)(a.y > b.y) | ((int)(a.y == b.y) & (int)(a.z > b.z)))));
}

// One warp, one long queue (n > kSortLong) at store[offset ...].  Returns false (nothing written) if a key is NaN.
__device__ __forceinline__ bool sortLongQueue(const FrameParams& P, SortScratch& W, BulkStage& B, unsigned int offset, unsigned int n) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned int size = 128;
    while (size < n) size <<= 1;                 // n <= kQueueCap <= kSortCap
    uint32_t* perm = W.info;
    bool nan = false;
    const unsigned int pad = bulkStage(P, W, B, offset, n);
    for (unsigned int f = lane; f < size; f += 32) {
        perm[f] = f;
        if (f < n) {
            const float4 t = W.thr[f];
            const uint32_t h = W.hdrRaw[pad + f];
            const Thr th{t.x, t.y, t.z, t.w};
            const float4 k = make_float4(t.x, tTopX(h, th), invSlope(h, th), __uint_as_float(h));
            W.key[f] = k;
            nan = nan || (t.x != t.x) || (t.y != t.y) || (t.z != t.z) || (t.w != t.w) || (k.z != k.z);
        }
    }
    if (__any_sync(full, nan)) return false;
    __syncwarp();
    // a precedes b: indices >= n are padding and sort last; equal keys keep the order of their indices
    auto precedes = [&](unsigned int a, unsigned int b) -> bool {
        if (a >= n || b >= n) return a < b;
        const float4 ka = W.key[a], kb = W.key[b];
        const bool bAfterA = keyBelow(kb, ka), aAfterB = keyBelow(ka, kb);
        return bAfterA || (!aAfterB && (a < b));
    };
    for (unsigned int k = 2; k <= size; k <<= 1) {
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            for (unsigned int p = lane; p < (size >> 1); p += 32) {
                const unsigned int i = ((p & ~(j - 1u)) << 1) | (p & (j - 1u)), l = i | j;
                const unsigned int a = perm[i], b = perm[l];
                const bool ascending = (i & k) == 0u;
                if (ascending ? precedes(b, a) : precedes(a, b)) { perm[i] = b; perm[l] = a; }
            }
            __syncwarp();
        }
    }
    for (unsigned int r = lane; r < n; r += 32) {
        const unsigned int e = perm[r];
        P.thrStore[offset + r] = W.thr[e];
        P.hdrStore[offset + r] = __float_as_uint(W.key[e].w);
    }
    __syncwarp();
    return true;
}

// One warp, the queues of one unit.  `recp`: the lane's thread record (null: no column-thread in this lane).
__device__ __forceinline__ void sortWarp(const FrameParams& P, SortScratch& W, BulkStage& B, ThreadRec* recp) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned int count = 0u, offset = 0u;
    if (recp) {
        const unsigned int c = recp->count;
        if (c != kRecInactive) { count = c; offset = recp->offset; }
    }
    if (!__any_sync(full, count > 0u)) return;
    bool sequential = false;   // this lane's queue holds a NaN key
    // long queues first, one at a time
    const bool isLong = count > kSortLong;
    for (unsigned longLanes = __ballot_sync(full, isLong); longLanes; longLanes &= longLanes - 1u) {
        const int l = __ffs((int)longLanes) - 1;
        const bool sorted = sortLongQueue(P, W, B, __shfl_sync(full, offset, l), __shfl_sync(full, count, l));
        if (!sorted && lane == l) sequential = true;
    }
    const unsigned longMask = __ballot_sync(full, isLong);
    if (isLong) count = 0u;    // (done, or left to the sequential sort below)
    const unsigned int ownCount = isLong ? recp->count : count;
    int laneBegin = 0;
    while (laneBegin < 32) {
        // ---- the batch: consecutive lanes from laneBegin whose queues fit the staging area together ----------
        unsigned int incl = lane >= laneBegin ? count : 0u;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int v = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += v;
        }
        // (a long queue lies between its neighbours in the store: a batch, one contiguous run, stops in front of it)
        const bool longBefore = lane > laneBegin && ((longMask >> laneBegin) & ((2u << (lane - laneBegin)) - 1u)) != 0u;
        const unsigned fits = __ballot_sync(full, lane >= laneBegin && incl <= (unsigned int)kSortCap && !longBefore) >> laneBegin;
        // lanes laneBegin .. laneEnd-1 (at least one: a single queue always fits)
        const int laneEnd = fits == (full >> laneBegin) ? 32 : laneBegin + __ffs((int)~fits) - 1;
        const bool inBatch = lane >= laneBegin && lane < laneEnd;
        const unsigned int total = __shfl_sync(full, incl, laneEnd - 1);
        __syncwarp();
        W.first[lane] = inBatch ? incl - count : total;
        W.offset[lane] = offset;
        if (lane == 0) { W.first[32] = total; W.nanMask = 0u; }
        __syncwarp();
        // ---- stage: the batch's thresholds are one contiguous run of the store ------------------------------------------
        // (its first element: the first non-empty queue of the batch)
        const unsigned nonEmpty = __ballot_sync(full, inBatch && count > 0u);
        if (nonEmpty == 0u) { laneBegin = laneEnd; continue; }
        const unsigned int runFirst = __shfl_sync(full, offset, __ffs((int)nonEmpty) - 1);
        const unsigned int pad = bulkStage(P, W, B, runFirst, total);
        for (unsigned int f = lane; f < total; f += 32) {
            // the queue element f belongs to: the last lane of the batch whose first index is <= f
            int lo = laneBegin, hi = laneEnd - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (W.first[mid] <= f) lo = mid; else hi = mid - 1;
            }
            const unsigned int qs = W.first[lo], n = W.first[lo + 1 < laneEnd ? lo + 1 : 32] - qs;
            const float4 t = W.thr[f];
            const uint32_t h = W.hdrRaw[pad + f];
            const Thr th{t.x, t.y, t.z, t.w};
            const float4 k = make_float4(t.x, tTopX(h, th), invSlope(h, th), __uint_as_float(h));
            W.key[f] = k;
            W.info[f] = (uint32_t)lo | (qs << 5) | (n << 16);
            // a NaN anywhere in a queue: that queue's lane sorts it sequentially (the order then depends on the comparison sequence)
            if ((t.x != t.x) || (t.y != t.y) || (t.z != t.z) || (t.w != t.w) || (k.z != k.z)) atomicOr(&W.nanMask, 1u << lo);
        }
        __syncwarp();
        const unsigned int owners = W.nanMask;
        sequential = sequential || ((owners >> lane) & 1u);
        // ---- rank + place ------------------------------------------------------------------------------------------------
        for (unsigned int f = lane; f < total; f += 32) {
            const uint32_t inf = W.info[f];
            const unsigned int q = inf & 31u, qs = (inf >> 5) & 0x7FFu, n = inf >> 16;
            if ((owners >> q) & 1u) continue;
            const float4 me = W.key[f];
            const unsigned int e = f - qs;
            unsigned int rank = 0u;
            for (unsigned int j = 0; j < e; j++) rank += keyBelow(W.key[qs + j], me) ? 0u : 1u;      // not after me and built before me
            for (unsigned int j = e + 1; j < n; j++) rank += keyBelow(me, W.key[qs + j]) ? 1u : 0u;  // I am strictly after it
            const unsigned int dst = W.offset[q] + rank;
            P.thrStore[dst] = W.thr[f];
            P.hdrStore[dst] = __float_as_uint(me.w);
        }
        __syncwarp();
        laneBegin = laneEnd;
    }
    if (sequential) {
        // the slice kernel hands such a thread to the replay, which follows the reference's insertion sequence
        recp->pad1 |= kRecUnordered;
        StoreQueue q{P.thrStore + offset, P.hdrStore + offset, (int)ownCount};
        sortQueue(q);
    }
}

}  // namespace gudni_dev
