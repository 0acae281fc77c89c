// binning.cu — tile binning on the GPU (level 2 of the C ABI).
//
// Reproduces the OUTCOME of the reference's CPU tile tree
//   buildTileTree / addShapeToTree / hSplit / vSplit / traverseTileTree
//       /root/reference/src/Graphics/Gudni/Raster/TileTree.hs:81-204
//   accumulateRasterJobs / addTileToRasterJob      Raster/Job.hs:132-178
// without replaying its one-shape-at-a-time insertion.  Two facts make a parallel restatement exact:
//   (1) a shape reaches a node with box [L,R)x[T,B) iff  left < R && right > L && top < B &&
//       bottom > T  (float compares against the integer cuts; every went-left cut on the path is
//       >= R and every went-right cut is <= L; the outermost bounds are never tested by the tree —
//       canvas culling supplies them — and forEachRoot leaves them untested too);
//   (2) a leaf splits iff the shapes that reach it number >= 127 or their strands sum to
//       >= maxStrandsPerTile (checkTileSpace :161-166 fails for some prefix iff it fails for the
//       whole set, counts being monotone) and the leaf is still larger than 8 px in its split axis.
// So the tree is a function of the shape SET.  Leaves come out in traverseTileTree order, which is
// y-major Morton order of the tile corner; per-tile lists are newest-first (:151) = descending
// scene index.
//
// Pipeline (no host round trip until the leaf totals are needed to size the output):
//   bin_root_count   thread per shape: atomics count shapes / strands per root tile
//   bin_root_scan    one CTA: exclusive scan of the root counts -> list offsets in the arena
//   bin_root_fill    thread per shape: scatter shape ids into the root lists
//   bin_subdivide    one CTA per root tile, one warp per node: top-down splitting with warp-ballot
//                    compaction of each node's list into its two children
//   bin_leaf_scan    one CTA: scan of per-root leaf / shape-reference totals
//   bin_emit         one CTA per root tile, one warp per leaf: TileInfo + rank-sorted Shape records
#include "binning.cuh"

#include <algorithm>

namespace {

constexpr int kMinTile = 8;        // mINtILEsIZE, Raster/Constants.hs:65
constexpr int kLeafShapeCap = 126; // a leaf accepts while tileShapeCount < mAXsHAPE - 1 (TileTree.hs:165)

struct BinNode {          // 32 bytes
    uint16_t x, y;        // corner relative to the root tile, pixels
    uint8_t hDepth, vDepth;
    uint8_t isV;          // VLeaf (splits top/bottom) or HLeaf (splits left/right)
    uint8_t closed;       // will not split again
    uint32_t start;       // list offset in the arena
    uint32_t count;
    uint32_t strands;
    uint32_t outStart;    // (emit) offset of the leaf's Shape records within the root tile's run
    uint32_t dest;        // (subdivide) index in the next frontier; bit 31 = node splits
    uint32_t slot;        // (subdivide) offset of the children's list space in this level's arena block
};
static_assert(sizeof(BinNode) == 32, "BinNode layout");

struct BinParams {
    const gudni_shape_entry* entries;
    int nEntries;
    int rootSize;          // pixels, power of two
    int rootDepth;         // log2(rootSize)
    int rootsPerSide;      // R
    int rowBegin, rowEnd;  // strip
    uint32_t maxStrands;
    int maxNodes;          // (rootSize / 8)^2
    int maxLevels;
    int threadsPerTile, tilesPerCall, columnsPerTile;
    // per root tile
    uint32_t *rootCount, *rootStrands, *rootStart, *rootCursor, *leafCount, *refCount, *tileOffset, *shapeOffset;
    BinNode* frontier;     // [roots][2][maxNodes]
    uint32_t* arena;
    unsigned long long arenaCap;
    unsigned long long* counters;  // [0] arena cursor, [1] overflow flag, [2] nTiles, [3] nShapeRefs
    // outputs
    gudni_tile* tiles;
    gudni_shape* shapes;
    int32_t* tileThreadBase;
};
enum { kArenaCursor = 0, kOverflow = 1, kTotalTiles = 2, kTotalRefs = 3, kTooManyShapes = 4 };

// y-major Morton rank of a root tile: traverseTileTree visits top before bottom, then left before right
__host__ __device__ inline uint32_t mortonYX(uint32_t tx, uint32_t ty) {
    uint32_t m = 0;
    for (int k = 0; k < 15; k++) m |= ((tx >> k) & 1u) << (2 * k) | ((ty >> k) & 1u) << (2 * k + 1);
    return m;
}
__device__ inline void unmortonYX(uint32_t m, uint32_t& tx, uint32_t& ty) {
    tx = ty = 0;
    for (int k = 0; k < 15; k++) {
        tx |= ((m >> (2 * k)) & 1u) << k;
        ty |= ((m >> (2 * k + 1)) & 1u) << k;
    }
}

__device__ __forceinline__ bool rootActive(const BinParams& P, uint32_t ty) {
    int top = (int)ty * P.rootSize;
    return top >= P.rowBegin && top < P.rowEnd;
}

// fact (1) above, with the reference's float compares (TileTree.hs:120-139).  The tree only ever tests a box
// against a cut, never against the outside edge of the square it covers, so a root tile in the first / last
// column or row takes everything on that side: for the culled, consistent boxes the ABI asks for that is no
// different from a plain overlap test, and for anything else (inverted or NaN boxes, shapes off the canvas)
// it is what addShapeToTree does with them.
template <class F>
__device__ __forceinline__ void forEachRoot(const BinParams& P, const float4 box, F&& f) {
    const int R = P.rootsPerSide, S = P.rootSize;
    // candidate range, one root to spare on each side; clamped as floats, because a box may be far outside
    // what an int holds (or infinite)
    const float last = (float)(R - 1);
    const int tx0 = (int)fminf(fmaxf(floorf(box.x / (float)S) - 1.0f, 0.0f), last);
    const int tx1 = (int)fmaxf(fminf(floorf(box.z / (float)S) + 1.0f, last), 0.0f);
    const int ty0 = (int)fminf(fmaxf(floorf(box.y / (float)S) - 1.0f, 0.0f), last);
    const int ty1 = (int)fmaxf(fminf(floorf(box.w / (float)S) + 1.0f, last), 0.0f);
    for (int ty = ty0; ty <= ty1; ty++) {
        if (!rootActive(P, ty)) continue;
        if (!((ty == R - 1 || box.y < (float)((ty + 1) * S)) && (ty == 0 || box.w > (float)(ty * S)))) continue;
        for (int tx = tx0; tx <= tx1; tx++)
            if ((tx == R - 1 || box.x < (float)((tx + 1) * S)) && (tx == 0 || box.z > (float)(tx * S))) f(mortonYX(tx, ty));
    }
}

__global__ void bin_root_count(const BinParams P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nEntries) return;
    const float4 box = __ldg(reinterpret_cast<const float4*>(P.entries + i) + 1);
    const uint32_t strands = __ldg(&P.entries[i].num_strands);
    forEachRoot(P, box, [&](uint32_t root) {
        atomicAdd(&P.rootCount[root], 1u);
        atomicAdd(&P.rootStrands[root], strands);
    });
}

// exclusive scan of `in` (n elements) by one CTA of 1024 threads; returns the total to all threads
__device__ uint32_t ctaExclusiveScan(const uint32_t* in, uint32_t* out, int n, uint32_t* smem /*33 words*/) {
    uint32_t running = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        uint32_t v = i < n ? in[i] : 0u;
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) smem[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < nWarps ? smem[lane] : 0u;
            uint32_t wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            smem[lane] = wi - w;          // exclusive warp offsets
            if (lane == 31) smem[32] = wi; // chunk total
        }
        __syncthreads();
        if (i < n) out[i] = running + smem[warp] + incl - v;
        running += smem[32];
        __syncthreads();
    }
    return running;
}

__global__ void __launch_bounds__(1024) bin_root_scan(const BinParams P) {
    __shared__ uint32_t smem[33];
    const int nRoots = P.rootsPerSide * P.rootsPerSide;
    uint32_t total = ctaExclusiveScan(P.rootCount, P.rootStart, nRoots, smem);
    for (int i = threadIdx.x; i < nRoots; i += blockDim.x) P.rootCursor[i] = 0u;
    if (threadIdx.x == 0) {
        P.counters[kArenaCursor] = total;
        if ((unsigned long long)total > P.arenaCap) P.counters[kOverflow] = 1ull;
    }
}

__global__ void bin_root_fill(const BinParams P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nEntries || P.counters[kOverflow]) return;
    const float4 box = __ldg(reinterpret_cast<const float4*>(P.entries + i) + 1);
    forEachRoot(P, box, [&](uint32_t root) {
        uint32_t slot = atomicAdd(&P.rootCursor[root], 1u);
        P.arena[P.rootStart[root] + slot] = (uint32_t)i;
    });
}

__device__ __forceinline__ int originXOf(const BinParams& P, uint32_t rtx) { return (int)rtx * P.rootSize; }
__device__ __forceinline__ int originYOf(const BinParams& P, uint32_t rty) { return (int)rty * P.rootSize; }

// One CTA per root tile.  The frontier of nodes is kept in tree order and ping-ponged between two
// arrays.  Per level: (A) thread per node decides whether it splits (fact 2) and a CTA-wide scan
// hands every node its place in the next frontier and its children's list space; (B) warp per
// node: each lane tests one shape of the node's list against the cut and the survivors of each
// half are compacted with ballot + popc prefix sums into the children's lists.
constexpr int kSubdivideThreads = 1024;   // 32 warps: a level's nodes are partitioned a warp each, and what a warp does is wait for loads
__global__ void __launch_bounds__(kSubdivideThreads) bin_subdivide(const BinParams P) {
    __shared__ uint32_t wN[32], wS[32];
    __shared__ uint32_t sNodes, sBase, sAnySplit, sAbort;
    const uint32_t root = blockIdx.x;
    uint32_t rtx, rty;
    unmortonYX(root, rtx, rty);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    BinNode* cur = P.frontier + (size_t)root * 2 * P.maxNodes;
    BinNode* nxt = cur + P.maxNodes;
    uint32_t side = 0;
    // Another CTA may raise the overflow flag at any moment (arena cursor check below): one thread reads it and the
    // CTA branches on the shared copy, so that its threads leave together or not at all.
    if (threadIdx.x == 0) sAbort = (!rootActive(P, rty) || P.counters[kOverflow]) ? 1u : 0u;
    __syncthreads();
    if (sAbort) {
        if (threadIdx.x == 0) { P.leafCount[root] = 0; P.refCount[root] = 0; P.rootCursor[root] = 0; }
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        BinNode n{};
        n.hDepth = n.vDepth = (uint8_t)P.rootDepth;
        n.isV = 1;
        n.start = P.rootStart[root];
        n.count = P.rootCount[root];
        n.strands = P.rootStrands[root];
        cur[0] = n;
        sNodes = 1;
        sAbort = 0;
    }
    __syncthreads();
    for (int level = 0; level < P.maxLevels; level++) {
        const uint32_t nNodes = sNodes;
        if (threadIdx.x == 0) sAnySplit = 0;
        __syncthreads();
        // ---- (A) decide + scan -----------------------------------------------------------------
        uint32_t runNodes = 0, runSlots = 0;
        for (uint32_t base = 0; base < nNodes; base += blockDim.x) {
            const uint32_t i = base + threadIdx.x;
            uint32_t outNodes = 0, outSlots = 0;
            bool split = false;
            if (i < nNodes) {
                const BinNode n = cur[i];
                if (!n.closed) {
                    const bool fits = n.count <= (uint32_t)kLeafShapeCap && n.strands < P.maxStrands;
                    const int extent = n.isV ? (1 << n.vDepth) : (1 << n.hDepth);
                    split = !fits && extent > kMinTile;
                    if (!split) cur[i].closed = 1;
                }
                outNodes = split ? 2u : 1u;
                outSlots = split ? 2u * n.count : 0u;
                if (split) sAnySplit = 1;
            }
            uint32_t inclN = outNodes, inclS = outSlots;
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t tn = __shfl_up_sync(0xffffffffu, inclN, d), ts = __shfl_up_sync(0xffffffffu, inclS, d);
                if (lane >= d) { inclN += tn; inclS += ts; }
            }
            if (lane == 31) { wN[warp] = inclN; wS[warp] = inclS; }
            __syncthreads();
            uint32_t offN = runNodes, offS = runSlots;
            for (int w = 0; w < nWarps; w++) {
                if (w < warp) { offN += wN[w]; offS += wS[w]; }
                runNodes += wN[w];
                runSlots += wS[w];
            }
            if (i < nNodes) {
                cur[i].dest = (offN + inclN - outNodes) | (split ? 0x80000000u : 0u);
                cur[i].slot = offS + inclS - outSlots;
            }
            __syncthreads();
        }
        if (!sAnySplit) break;
        if (threadIdx.x == 0) {
            const unsigned long long b = atomicAdd(&P.counters[kArenaCursor], (unsigned long long)runSlots);
            if (b + runSlots > P.arenaCap) { P.counters[kOverflow] = 1ull; sAbort = 1; }
            sBase = (uint32_t)b;
        }
        __syncthreads();
        if (sAbort) break;
        const uint32_t arenaBase = sBase;
        // ---- (B) warp per node: copy closed nodes, partition the lists of splitting ones ---------
        for (uint32_t i = warp; i < nNodes; i += nWarps) {
            const BinNode n = cur[i];
            const uint32_t dest = n.dest & 0x7FFFFFFFu;
            if (!(n.dest & 0x80000000u)) {
                if (lane == 0) nxt[dest] = n;
                continue;
            }
            BinNode c0 = n, c1 = n;
            float cut;
            if (n.isV) {   // vSplit, TileTree.hs:181-190: top / bottom HLeafs, vDepth - 1
                const int half = (1 << n.vDepth) >> 1;
                c0.vDepth = c1.vDepth = n.vDepth - 1;
                c1.y = n.y + half;
                cut = (float)(originYOf(P, rty) + n.y + half);
            } else {       // hSplit, :169-178: left / right VLeafs, hDepth - 1
                const int half = (1 << n.hDepth) >> 1;
                c0.hDepth = c1.hDepth = n.hDepth - 1;
                c1.x = n.x + half;
                cut = (float)(originXOf(P, rtx) + n.x + half);
            }
            c0.isV = c1.isV = n.isV ? 0 : 1;
            c0.closed = c1.closed = 0;
            c0.start = arenaBase + n.slot;
            c1.start = arenaBase + n.slot + n.count;
            uint32_t n0 = 0, n1 = 0, s0 = 0, s1 = 0;
            for (uint32_t j = 0; j < n.count; j += 32) {
                const uint32_t k = j + lane;
                bool f0 = false, f1 = false;
                uint32_t shape = 0, strands = 0;
                if (k < n.count) {
                    shape = P.arena[n.start + k];
                    const float4 box = __ldg(reinterpret_cast<const float4*>(P.entries + shape) + 1);
                    strands = __ldg(&P.entries[shape].num_strands);
                    // insertShapeV / insertShapeH on a branch, TileTree.hs:120-139
                    f0 = n.isV ? (box.y < cut) : (box.x < cut);
                    f1 = n.isV ? (box.w > cut) : (box.z > cut);
                }
                const uint32_t b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
                const uint32_t below = (1u << lane) - 1u;
                if (f0) { P.arena[c0.start + n0 + __popc(b0 & below)] = shape; s0 += strands; }
                if (f1) { P.arena[c1.start + n1 + __popc(b1 & below)] = shape; s1 += strands; }
                n0 += __popc(b0);
                n1 += __popc(b1);
            }
            for (int d = 16; d > 0; d >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, d);
                s1 += __shfl_xor_sync(0xffffffffu, s1, d);
            }
            if (lane == 0) {
                c0.count = n0; c0.strands = s0;
                c1.count = n1; c1.strands = s1;
                nxt[dest] = c0;
                nxt[dest + 1] = c1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) sNodes = runNodes;
        BinNode* t = cur; cur = nxt; nxt = t;
        side ^= 1u;
        __syncthreads();
    }
    // ---- leaves: offsets of their Shape records within this root tile's run ----------------------
    const uint32_t nLeaves = sNodes;
    uint32_t running = 0;
    for (uint32_t base = 0; base < nLeaves; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nLeaves ? cur[i].count : 0u;
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wN[warp] = incl;
        __syncthreads();
        uint32_t off = running;
        for (int w = 0; w < nWarps; w++) {
            if (w < warp) off += wN[w];
            running += wN[w];
        }
        if (i < nLeaves) cur[i].outStart = off + incl - v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        P.leafCount[root] = sAbort ? 0u : nLeaves;
        P.refCount[root] = sAbort ? 0u : running;
        P.rootCursor[root] = side;   // which frontier array holds the leaves
    }
}

__global__ void __launch_bounds__(1024) bin_leaf_scan(const BinParams P) {
    __shared__ uint32_t smem[33];
    const int nRoots = P.rootsPerSide * P.rootsPerSide;
    const uint32_t tiles = ctaExclusiveScan(P.leafCount, P.tileOffset, nRoots, smem);
    const uint32_t refs = ctaExclusiveScan(P.refCount, P.shapeOffset, nRoots, smem);
    if (threadIdx.x == 0) {
        P.counters[kTotalTiles] = tiles;
        P.counters[kTotalRefs] = refs;
    }
}

// One CTA per root tile, one warp per leaf: the TileInfo record (Raster/Types.hs:176-198) with the
// column allocation addTileToRasterJob would have given it (Raster/Job.hs:132-178), and the leaf's
// Shape records newest-first (descending scene index) by rank sort.
// (kEmitParts CTAs share a root tile's leaves: a frame has about as many root tiles as the chip has SMs, and one CTA per SM
// leaves the rank sort waiting on its own loads.)
constexpr unsigned kEmitParts = 4;
__global__ void __launch_bounds__(256) bin_emit(const BinParams P) {
    const uint32_t parts = gridDim.x / (uint32_t)(P.rootsPerSide * P.rootsPerSide);   // 1 or kEmitParts (binScene)
    const uint32_t root = blockIdx.x / parts, part = blockIdx.x % parts;
    const uint32_t nLeaves = P.leafCount[root];
    if (nLeaves == 0) return;
    uint32_t rtx, rty;
    unmortonYX(root, rtx, rty);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const BinNode* fin = P.frontier + (size_t)root * 2 * P.maxNodes + (size_t)P.rootCursor[root] * P.maxNodes;
    const uint32_t tileBase = P.tileOffset[root], shapeBase = P.shapeOffset[root];
    for (uint32_t i = part * nWarps + warp; i < nLeaves; i += nWarps * parts) {
        const BinNode n = fin[i];
        const uint32_t tileIdx = tileBase + i;
        const uint32_t shapeStart = shapeBase + n.outStart;
        if (lane == 0) {
            if (n.count > 65535u) P.counters[kTooManyShapes] = 1ull;   // the raster kernels index a tile's list with 16 bits
            gudni_tile t;
            t.left = originXOf(P, rtx) + n.x;
            t.top = originYOf(P, rty) + n.y;
            t.right = t.left + (1 << n.hDepth);
            t.bottom = t.top + (1 << n.vDepth);
            t.h_depth = n.hDepth;
            t.v_depth = n.vDepth;
            t.column_allocation = (int32_t)((tileIdx % (uint32_t)P.tilesPerCall) * (uint32_t)P.columnsPerTile);
            t.shape_start = shapeStart;
            t.shape_count = n.count;
            P.tiles[tileIdx] = t;
            P.tileThreadBase[tileIdx] = (int32_t)(tileIdx * (uint32_t)P.columnsPerTile);
        }
        for (uint32_t j = lane; j < n.count; j += 32) {
            const uint32_t e = P.arena[n.start + j];
            uint32_t rank = 0;
            for (uint32_t k = 0; k < n.count; k++) rank += (P.arena[n.start + k] > e) ? 1u : 0u;
            const uint4 rec = __ldg(reinterpret_cast<const uint4*>(P.entries + e));   // tag, geo_start, num_strands
            reinterpret_cast<uint4*>(P.shapes)[shapeStart + rank] = rec;
        }
    }
}

}  // namespace

#ifndef GUDNI_HOST_EMULATION   // the emulator drives the kernels itself (tests/native/raster_emu.cpp)
namespace gudni_bin {

static int log2ceil(int x) { int d = 0; while ((1 << d) < x) d++; return d; }

int binScene(gudni_ctx* ctx, const gudni_shape_entry* devEntries, int n, int (*whileBinning)(gudni_ctx*)) {
    const int canvasDepth = log2ceil(std::max(ctx->width, ctx->height));   // adjustedLog, TileTree.hs:74-75
    const int tileDepth = log2ceil(ctx->spec.max_tile_size);
    BinParams P{};
    P.entries = devEntries;
    P.nEntries = n;
    P.rootDepth = std::min(canvasDepth, tileDepth);
    P.rootSize = 1 << P.rootDepth;
    P.rootsPerSide = (1 << canvasDepth) / P.rootSize;
    // the tile tree covers the power-of-two square around the canvas (TileTree.hs:88-93): root tiles
    // below the last canvas row still exist (their threads are inactive) and belong to the last strip
    P.rowBegin = ctx->rowBegin;
    P.rowEnd = (ctx->rowEnd >= ctx->height) ? (1 << 30) : ctx->rowEnd;
    P.maxStrands = (uint32_t)ctx->spec.max_strands_per_tile;
    const int cells = std::max(1, P.rootSize / kMinTile);
    P.maxNodes = cells * cells;
    P.maxLevels = 2 * std::max(0, P.rootDepth - 3) + 2;
    P.threadsPerTile = ctx->spec.threads_per_tile;
    // buildRasterJobs swaps the two arguments of accumulateRasterJobs (OpenCL/CallKernels.hs:254 vs
    // Raster/Job.hs:151-156): tiles per job = threadsPerTile, columns per tile = maxTilesPerCall.
    P.tilesPerCall = ctx->spec.threads_per_tile;
    P.columnsPerTile = ctx->spec.max_tiles_per_call;
    if (P.rootsPerSide > 32768 / 1) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "canvas too large for the binning grid");
    const size_t nRoots = (size_t)P.rootsPerSide * P.rootsPerSide;
    // A 4K frame is about as many root tiles as the chip has SMs, each with hundreds of shapes to split six levels deep: large
    // CTAs (bin_subdivide) and several CTAs per root tile (bin_emit).  A 16K canvas is thousands of root tiles of a few dozen
    // shapes that never split: small CTAs, one per root tile (S5: 0.09 ms against 0.21 the other way round).
    const bool fewRoots = nRoots <= 512;

    GUDNI_TRY(devEnsure(ctx, ctx->binWork[0], nRoots * 8 * sizeof(uint32_t)));
    GUDNI_TRY(devEnsure(ctx, ctx->binWork[1], nRoots * 2 * (size_t)P.maxNodes * sizeof(BinNode)));
    GUDNI_TRY(devEnsure(ctx, ctx->binCounters, 64));
    size_t wantArena = (size_t)32 * (size_t)std::max(n, 1) + ((size_t)4 << 20);
    if (ctx->binWork[2].cap < wantArena * 4) GUDNI_TRY(devEnsure(ctx, ctx->binWork[2], wantArena * 4));
    if (!ctx->pinned) {
        GUDNI_CUDA_TRY(ctx, cudaMallocHost(&ctx->pinned, 4096));
        ctx->pinnedCap = 4096;
    }
    uint32_t* w = ctx->binWork[0].as<uint32_t>();
    P.rootCount = w; P.rootStrands = w + nRoots; P.rootStart = w + 2 * nRoots; P.rootCursor = w + 3 * nRoots;
    P.leafCount = w + 4 * nRoots; P.refCount = w + 5 * nRoots; P.tileOffset = w + 6 * nRoots; P.shapeOffset = w + 7 * nRoots;
    P.frontier = ctx->binWork[1].as<BinNode>();
    P.counters = ctx->binCounters.as<unsigned long long>();

    volatile unsigned long long* host = static_cast<volatile unsigned long long*>(ctx->pinned);
    for (int attempt = 0; attempt < 6; attempt++) {
        P.arena = ctx->binWork[2].as<uint32_t>();
        P.arenaCap = ctx->binWork[2].cap / 4;
        GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(w, 0, nRoots * 2 * sizeof(uint32_t), ctx->stream));
        GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(P.counters, 0, 64, ctx->stream));
        const int blocks = (n + 255) / 256;
        if (n) { bin_root_count<<<blocks, 256, 0, ctx->stream>>>(P); ctx->launches++; }
        bin_root_scan<<<1, 1024, 0, ctx->stream>>>(P); ctx->launches++;
        if (n) { bin_root_fill<<<blocks, 256, 0, ctx->stream>>>(P); ctx->launches++; }
        bin_subdivide<<<(unsigned)nRoots, fewRoots ? kSubdivideThreads : 256, 0, ctx->stream>>>(P); ctx->launches++;
        bin_leaf_scan<<<1, 1024, 0, ctx->stream>>>(P); ctx->launches++;
        GUDNI_CUDA_TRY(ctx, cudaGetLastError());
        // the kernels above are queued: host work that can go on beside them (the upload of the geometry heap, shim.cu)
        if (whileBinning) GUDNI_TRY(whileBinning(ctx));
        GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pinned, P.counters, 32, cudaMemcpyDeviceToHost, ctx->stream));
        GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (!host[kOverflow]) break;
        if (attempt == 5) return ctxFail(ctx, GUDNI_ERR_OOM, "tile binning scratch exhausted");
        GUDNI_TRY(devEnsure(ctx, ctx->binWork[2], ctx->binWork[2].cap * 4));   // grow and redo
    }
    const int64_t nTiles = (int64_t)host[kTotalTiles], nRefs = (int64_t)host[kTotalRefs];
    GUDNI_TRY(devEnsure(ctx, ctx->tiles, (size_t)std::max<int64_t>(nTiles, 1) * 32));
    GUDNI_TRY(devEnsure(ctx, ctx->shapes, (size_t)nRefs * 16 + 16));
    GUDNI_TRY(devEnsure(ctx, ctx->tileThreadBase, (size_t)std::max<int64_t>(nTiles, 1) * 4));
    P.tiles = ctx->tiles.as<gudni_tile>();
    P.shapes = ctx->shapes.as<gudni_shape>();
    P.tileThreadBase = ctx->tileThreadBase.as<int32_t>();
    bin_emit<<<(unsigned)nRoots * (fewRoots ? kEmitParts : 1u), 256, 0, ctx->stream>>>(P); ctx->launches++;
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    ctx->nTiles = nTiles;
    ctx->nShapes = nRefs;
    ctx->nColumns = nTiles * (int64_t)P.columnsPerTile;
    ctx->binUsed = true;
    return GUDNI_OK;
}

}  // namespace gudni_bin
#endif  // GUDNI_HOST_EMULATION
