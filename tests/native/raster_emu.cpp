// raster_emu.cpp — TEST HELPER: this repo's raster kernels (gudni_b200/csrc/raster_kernels.cu with
// raster_device.cuh / raster_warp.cuh, the very text nvcc compiles) built with g++ against a host stand-in for
// the CUDA device language (emu/cuda_runtime.h) and run under a cooperative SIMT emulator: every CUDA thread a
// fiber, warp collectives and __syncthreads as rendezvous points, one CTA at a time.  It is slow (small scenes
// only) and it is not a performance model; what it gives is the kernels' logic — queues, warp-cooperative
// sweep, colour cache, spill replay — checked against the oracle in the CPU test suite, deadlock detection for
// mismatched collectives, and a way to try a kernel change before a GPU is at hand.
// The shared-reciprocal division (div3) runs as on the device, from a model of rcp.approx whose error the tests
// can set (raster_emu_set_rcp_error).  Not product, not oracle.
#define GUDNI_HOST_EMULATION 1
#include <cuda_runtime.h>   // emu/cuda_runtime.h

namespace cuemu {
int rcpUlpError = 0;
int scheduleMode = 0;              // 0: lanes in ascending order, 1: descending, 2: a fresh pseudo-random order every pass
static uint64_t scheduleRng = 0x9E3779B97F4A7C15ull;
State S;
unsigned char dynamicShared[228 * 1024];
static void (*g_entry)(void*);
static void* g_args;
static void trampoline() {
    g_entry(g_args);
    S.cur->done = true;
    S.progress = true;
}
void runBlock(void (*entry)(void*), void* args, dim3 grid, dim3 block, uint3 bidx) {
    const size_t n = block.x;
    constexpr size_t kStack = 192 * 1024;
    g_entry = entry;
    g_args = args;
    S.bIdx = bidx;
    S.bDim = uint3{block.x, 1, 1};
    S.gDim = uint3{grid.x, 1, 1};
    if (S.fibers.size() < n) S.fibers.resize(n);
    S.warps.assign((n + 31) / 32, WarpSync{});
    S.barArrived = S.barRelease = 0;
    for (size_t i = 0; i < S.fibers.size(); i++) S.fibers[i].done = true;
    for (size_t i = 0; i < n; i++) {
        Fiber& f = S.fibers[i];
        if (f.stack.size() < kStack) f.stack.resize(kStack);
        f.tid = uint3{(unsigned)i, 0, 0};
        f.warp = (int)(i / 32);
        f.lane = (int)(i % 32);
        f.done = false;
        f.wait = kWaitNone;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = &S.scheduler;
        makecontext(&f.ctx, trampoline, 0);
    }
    size_t remaining = n;
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = scheduleMode == 1 ? n - 1 - i : i;
    while (remaining) {
        S.progress = false;
        if (scheduleMode == 2)      // between two rendezvous points the threads run in an arbitrary order, as on hardware:
            for (size_t i = n; i > 1; i--) {   // code that needs a __syncwarp it does not have shows up as a wrong image
                scheduleRng = scheduleRng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[i - 1], order[(scheduleRng >> 33) % i]);
            }
        for (size_t k = 0; k < n; k++) {
            const size_t i = order[k];
            Fiber& f = S.fibers[i];
            if (f.done || !canRun(f)) continue;
            S.cur = &f;
            S.switches++;
            swapcontext(&S.scheduler, &f.ctx);
            if (f.done) {
                remaining--;
                tryComplete(f.warp);     // the lanes it leaves behind may now be complete
            }
        }
        if (!S.progress && remaining) {
            fprintf(stderr, "cuemu: deadlock in block %u: %zu threads wait on collectives that cannot complete\n", bidx.x, remaining);
            for (size_t w = 0; w < S.warps.size(); w++)
                fprintf(stderr, "  warp %zu: arrived %08x release %08x live %08x op %d\n", w, S.warps[w].arrived,
                        S.warps[w].release, liveMask((int)w), S.warps[w].op);
            abort();
        }
    }
}
}  // namespace cuemu

#include "../../gudni_b200/csrc/raster_kernels.cu"
#include "../../gudni_b200/csrc/binning.cu"
#include "../../gudni_b200/csrc/strands.cu"

#include <cstring>
#include <set>
#include <tuple>

using namespace gudni_dev;

static size_t g_storeEntriesOverride;
static size_t g_streamChunksOverride;
static size_t g_refSlabsOverride;
static int g_laneShift;
static int g_batches = 1;
static int g_rowBegin = 0, g_rowEnd = 0;    // raster_emu_set_strip: rows of the canvas this "context" renders (0,0 = all)
namespace {

struct FrameInputs {
    const void* geometry; size_t geometryBytes;
    const float* substances; const uint8_t* pictureBytes; const gudni_picture_use* pictureUses; const float* background;
    int width, height; const gudni_spec* spec;
};

// rasterTiles + rasterSpill of raster_kernels.cu with host buffers and emulated launches
void rasterStage(const FrameInputs& in, const gudni_shape* shapes, int64_t nShapes, const void* boundsRecords, int boundsStride,
                 int64_t nBoundsRecords, const gudni_tile* tiles, const int32_t* threadBase, int nTiles, int64_t nColumns,
                 uint32_t* out, int32_t* dbgThresholds, int32_t* dbgShapeBits, int64_t* stats) {
    FrameParams P{};
    int depth = 0;
    while ((1 << depth) < in.spec->threads_per_tile) depth++;
    std::vector<unsigned long long> counters(kCountersBytes / 8 + 8, 0ull);
    const int spillCapacity = 1 << 16, spillSlots = 128;
    std::vector<unsigned long long> spillList(spillCapacity);
    const size_t threads = (size_t)nTiles * in.spec->threads_per_tile;
    const size_t entries = g_storeEntriesOverride ? g_storeEntriesOverride
                                                  : std::max<size_t>({threads * 24, (size_t)in.width * in.height / 4, (size_t)1 << 16});
    std::vector<float4> thrStore(entries);
    std::vector<uint32_t> hdrStore(entries + 16);   // (+16 words: bulk copies of headers are rounded out to 16 bytes)
    std::vector<ThreadRec> recs(std::max<size_t>(threads, 32));
    std::vector<uint32_t> order(std::max(nTiles, 1));
    std::vector<float2> bounds(in.geometryBytes / 16 + 2);
    std::vector<float4> spillThr((size_t)spillSlots * in.spec->max_thresholds);
    std::vector<uint32_t> spillHdr((size_t)spillSlots * in.spec->max_thresholds);
    P.geometry = static_cast<const uint8_t*>(in.geometry);
    P.shapes = shapes;
    std::vector<gudni_tile> tilesCopy(tiles, tiles + std::max(nTiles, 0));   // the shim's device copy (tile_order_kernel may edit it)
    P.tiles = tilesCopy.data();
    P.tileThreadBase = threadBase;
    P.substances = reinterpret_cast<const float4*>(in.substances);
    P.pictureData = in.pictureBytes;
    P.pictureUses = in.pictureUses;
    P.out = out;
    P.background = make_float4(in.background[0], in.background[1], in.background[2], in.background[3]);
    P.width = in.width; P.height = in.height;
    P.rowBegin = g_rowEnd ? g_rowBegin : 0;
    P.rowEnd = g_rowEnd ? g_rowEnd : in.height;
    P.rowOrigin = 0;                                   // `out` is the whole canvas here
    P.computeDepth = depth;
    P.maxShape = in.spec->max_shapes;
    P.maxThresholds = in.spec->max_thresholds;
    P.dbgThresholds = dbgThresholds;
    P.dbgShapeBits = dbgShapeBits;
    P.counters = counters.data();
    P.spillList = spillList.data();
    P.spillCapacity = spillCapacity;
    P.thrStore = thrStore.data();
    P.hdrStore = hdrStore.data();
    P.storeCap = entries;
    P.threadRecs = recs.data();
    P.strandBounds = bounds.data();
    P.tileOrder = order.data();
    P.laneShift = g_laneShift;
    // section-stream pool (raster_split.cuh): generous (the shim sizes it from the last frame's demand and retries a frame that ran dry), or what raster_emu_set_stream_chunks forces
    const size_t chunks = g_streamChunksOverride ? g_streamChunksOverride
                                                 : std::max<size_t>(threads * 96 + (size_t)in.width * in.height / 16, (size_t)1 << 14);
    std::vector<uint2> streamPool(chunks * kChunkRecs);
    P.streamPool = streamPool.data();
    P.streamCapChunks = (unsigned int)chunks;
    const size_t refSlabs = g_refSlabsOverride ? g_refSlabsOverride : std::max<size_t>(threads / 4 + (size_t)in.width * in.height / 64, 64);
    std::vector<ulonglong2> stackKeys(refSlabs * kRefSlab);
    std::vector<float4> stackColors(refSlabs * kRefSlab);
    std::vector<uint2> refSlabInfo(refSlabs);
    P.stackKeys = stackKeys.data();
    P.stackColors = stackColors.data();
    P.refSlabs = refSlabInfo.data();
    P.refCapSlabs = (unsigned int)refSlabs;
    if (dbgThresholds) for (int64_t i = 0; i < nColumns; i++) dbgThresholds[i] = -1;
    if (dbgShapeBits) for (int64_t i = 0; i < nColumns; i++) dbgShapeBits[i] = -1;
    if (nTiles > 0) {
        if (nBoundsRecords > 0)
            cuemu::launch(strand_bounds_kernel, dim3((unsigned)((nBoundsRecords + 255) / 256)), dim3(256), P.geometry, in.geometryBytes,
                          static_cast<const uint8_t*>(boundsRecords), boundsStride, (int)nBoundsRecords, bounds.data(), counters.data());
        cuemu::launch(tile_order_kernel, dim3(1), dim3(256), tilesCopy.data(), 0, nTiles, order.data(), counters.data());
        // the batches of rasterTiles (one stream each on the GPU), here one after the other
        const int batches = std::max(1, std::min({g_batches, kMaxBatches, nTiles}));
        const unsigned int regionSlabs = (unsigned int)refSlabs / (unsigned int)batches;
        unsigned int* work = reinterpret_cast<unsigned int*>(counters.data() + 32);
        std::vector<unsigned int> wideList((size_t)(nTiles + kMaxBatches) * (size_t)(in.spec->threads_per_tile / 8));
        for (int b = 0; b < batches; b++) {
            const int tilesHere = (nTiles - b + batches - 1) / batches;
            P.work = work + (size_t)b * kWorkWords;
            P.batchStride = batches;
            P.batchCount = batches;
            P.batchIndex = b;
            P.refSlabBase = (unsigned int)b * regionSlabs;
            P.refCapSlabs = regionSlabs;
            P.wideList = wideList.data() + (size_t)b * ((size_t)(nTiles + batches - 1) / batches) * (size_t)(in.spec->threads_per_tile / 8);
            cuemu::launch(raster_generate_kernel, dim3(2), dim3(in.spec->threads_per_tile), P, 0, tilesHere);
            cuemu::launch(raster_sort_kernel, dim3(2), dim3(kSortWarpsPerCta * 32), P, 0, tilesHere);
            cuemu::launch(raster_slice_kernel, dim3(2), dim3(kSliceWarpsPerCta * 32), P, 0, tilesHere);
            cuemu::launch(raster_slice_wide_kernel, dim3(2), dim3(32), P);
            cuemu::launch(raster_resolve_kernel, dim3(2), dim3(kResolveWarpsPerCta * 32), P, 0, tilesHere);
            cuemu::launch(raster_composite_kernel, dim3(2), dim3(kCompositeWarpsPerCta * 32), P);
            cuemu::launch(raster_accumulate_kernel, dim3(2), dim3(kAccumulateWarpsPerCta * 32), P, 0, tilesHere);
            cuemu::launch(raster_picture_kernel, dim3(2), dim3(kColorWarpsPerCta * 32), P, 0, tilesHere);
        }
        cuemu::launch(raster_spill_kernel, dim3(1), dim3(spillSlots), P, spillThr.data(), spillHdr.data(), spillSlots);
    }
    (void)nShapes;
    if (getenv("GUDNI_EMU_STACKS")) {   // how well the resolve kernel's per-warp cache deduplicates: numbered stacks vs distinct (tile, stack) pairs
        std::set<std::tuple<unsigned, unsigned long long, unsigned long long>> distinct;
        unsigned long long numbered = 0;
        const unsigned int* work = reinterpret_cast<const unsigned int*>(counters.data() + 32);
        for (unsigned s = 0; s < work[kWorkRefSlabs] && s < P.refCapSlabs; s++) {
            const uint2 info = refSlabInfo[s];
            numbered += info.y;
            for (unsigned i = 0; i < info.y; i++) {
                const ulonglong2 k = stackKeys[(size_t)s * kRefSlab + i];
                distinct.insert(std::make_tuple(info.x, k.x, k.y));
            }
        }
        fprintf(stderr, "[emu] stacks numbered %llu, distinct per tile %zu, tiles %d\n", numbered, distinct.size(), nTiles);
    }
    if (stats) {
        stats[0] = (int64_t)counters[kCntThresholds];
        stats[1] = (int64_t)counters[kCntSpilled];
        stats[2] = (int64_t)counters[kCntOverflow];
        stats[3] = (int64_t)cuemu::S.switches;
        stats[4] = (int64_t)counters[kCntNonFinite];
        stats[5] = (int64_t)counters[kCntExhausted];
        stats[6] = (int64_t)counters[kCntStreamCursor];
        stats[7] = (int64_t)counters[kCntRefSlabs];
    }
}

}  // namespace
// (g_storeEntriesOverride, declared above: raster_emu_set_store_entries forces the threshold store to run out)
namespace {

int log2ceil(int x) { int d = 0; while ((1 << d) < x) d++; return d; }

// binScene of binning.cu with host buffers and emulated launches
struct Binned {
    std::vector<gudni_tile> tiles;
    std::vector<gudni_shape> shapes;
    std::vector<int32_t> threadBase;
    int64_t nColumns = 0;
    bool overflow = false;
};
void binStage(const FrameInputs& in, const gudni_shape_entry* entries, int n, Binned& out) {
    const int canvasDepth = log2ceil(std::max(in.width, in.height));
    const int tileDepth = log2ceil(in.spec->max_tile_size);
    BinParams P{};
    P.entries = entries;
    P.nEntries = n;
    P.rootDepth = std::min(canvasDepth, tileDepth);
    P.rootSize = 1 << P.rootDepth;
    P.rootsPerSide = (1 << canvasDepth) / P.rootSize;
    // as binScene: the tile tree covers the power-of-two square around the canvas; root tiles below the last
    // canvas row belong to the last strip
    P.rowBegin = g_rowEnd ? g_rowBegin : 0;
    P.rowEnd = (!g_rowEnd || g_rowEnd >= in.height) ? (1 << 30) : g_rowEnd;
    P.maxStrands = (uint32_t)in.spec->max_strands_per_tile;
    const int cells = std::max(1, P.rootSize / kMinTile);
    P.maxNodes = cells * cells;
    P.maxLevels = 2 * std::max(0, P.rootDepth - 3) + 2;
    P.threadsPerTile = in.spec->threads_per_tile;
    P.tilesPerCall = in.spec->threads_per_tile;
    P.columnsPerTile = in.spec->max_tiles_per_call;
    const size_t nRoots = (size_t)P.rootsPerSide * P.rootsPerSide;
    std::vector<uint32_t> w(nRoots * 8, 0u);
    std::vector<BinNode> frontier(nRoots * 2 * (size_t)P.maxNodes);
    std::vector<uint32_t> arena((size_t)32 * std::max(n, 1) + ((size_t)4 << 20));
    std::vector<unsigned long long> counters(8, 0ull);
    P.rootCount = w.data(); P.rootStrands = w.data() + nRoots; P.rootStart = w.data() + 2 * nRoots; P.rootCursor = w.data() + 3 * nRoots;
    P.leafCount = w.data() + 4 * nRoots; P.refCount = w.data() + 5 * nRoots; P.tileOffset = w.data() + 6 * nRoots; P.shapeOffset = w.data() + 7 * nRoots;
    P.frontier = frontier.data();
    P.counters = counters.data();
    P.arena = arena.data();
    P.arenaCap = arena.size();
    const int blocks = (n + 255) / 256;
    if (n) cuemu::launch(bin_root_count, dim3(blocks), dim3(256), P);
    cuemu::launch(bin_root_scan, dim3(1), dim3(1024), P);
    if (n) cuemu::launch(bin_root_fill, dim3(blocks), dim3(256), P);
    cuemu::launch(bin_subdivide, dim3((unsigned)nRoots), dim3(256), P);
    cuemu::launch(bin_leaf_scan, dim3(1), dim3(1024), P);
    out.overflow = counters[kOverflow] != 0 || counters[kTooManyShapes] != 0;
    const int64_t nTiles = (int64_t)counters[kTotalTiles], nRefs = (int64_t)counters[kTotalRefs];
    out.tiles.assign(std::max<int64_t>(nTiles, 1), gudni_tile{});
    out.shapes.assign(nRefs + 1, gudni_shape{});
    out.threadBase.assign(std::max<int64_t>(nTiles, 1), 0);
    P.tiles = out.tiles.data();
    P.shapes = out.shapes.data();
    P.tileThreadBase = out.threadBase.data();
    cuemu::launch(bin_emit, dim3((unsigned)nRoots * kEmitParts), dim3(256), P);
    out.tiles.resize(nTiles);
    out.shapes.resize(nRefs);
    out.threadBase.resize(nTiles);
    out.nColumns = nTiles * (int64_t)P.columnsPerTile;
}

}  // namespace

extern "C" {

// 0 restores the shim's sizing rule.  A store that is too small makes the generate kernel hand whole warps to
// the replay kernel (raster_warp.cuh generateWarp): slow, not wrong.
void raster_emu_set_store_entries(size_t n) { g_storeEntriesOverride = n; }
// the same for the section-stream pool between the slice and the colour kernel (chunks of 16 records)
void raster_emu_set_stream_chunks(size_t n) { g_streamChunksOverride = n; }
// ... and for the table of distinct shape stacks (slabs of 128 numbers)
void raster_emu_set_ref_slabs(size_t n) { g_refSlabsOverride = n; }
// the render kernels' units are 32 >> shift column-threads wide (what the shim picks for launches of few tiles)
void raster_emu_set_lane_shift(int shift) { g_laneShift = shift; }
// batches per launch (rasterTiles of raster_kernels.cu): emulated one after the other
void raster_emu_set_batches(int n) { g_batches = n < 1 ? 1 : n; }

// order in which the emulator runs the threads of a CTA between rendezvous points (see runBlock)
void raster_emu_set_schedule(int mode) { cuemu::scheduleMode = mode; }

// ulps by which the modelled rcp.approx misses the correctly rounded reciprocal (0, +-1, +-2 ...)
void raster_emu_set_rcp_error(int ulps) { cuemu::rcpUlpError = ulps; }

// div3<true> (raster_device.cuh) against the host's IEEE division on n pseudo-random operand sets, the same
// generator as the device's selftest_div3_kernel; returns the number of differing quotients.
uint64_t raster_emu_selftest_div3(uint64_t n, uint64_t seed) {
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        unsigned long long z = seed + i * 0x9E3779B97F4A7C15ull;
        float v[4];
        for (int k = 0; k < 4; k++) {
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const unsigned mode = (unsigned)(z >> 60);
            float f = (float)((z >> 20) & 0xFFFFFFull) * (1.0f / 16777216.0f);
            if (mode == 0) f = __uint_as_float((unsigned)(z >> 8) & 0x7FFFFFFFu);
            else if (mode == 1) f = f * 0x1p-58f;
            else if (mode == 2) f = 0.0f;
            else if (mode == 3) f = 1.0f;
            v[k] = f;
        }
        const float d = v[3];
        if (!(d > 0.0f) || v[0] != v[0] || v[1] != v[1] || v[2] != v[2] || d != d) continue;
        float qx, qy, qz;
        div3<true>(v[0], v[1], v[2], d, qx, qy, qz);
        const volatile float rx = v[0] / d, ry = v[1] / d, rz = v[2] / d;
        bad += (__float_as_uint(qx) != __float_as_uint(rx)) + (__float_as_uint(qy) != __float_as_uint(ry)) +
               (__float_as_uint(qz) != __float_as_uint(rz));
    }
    return bad;
}

// gudni_b200_frame_strip: whole root-tile rows [row_begin, row_end) of the canvas; (0, 0) restores the whole frame.
void raster_emu_set_strip(int row_begin, int row_end) { g_rowBegin = row_begin; g_rowEnd = row_end; }

// Level 1: the jobs' shapes and tiles laid end to end as the shim lays them (shape_start rebased,
// thread_base = first column-thread of each tile).  stats[0..4] (room for 8) = thresholds, spilled threads, overflowed
// threads, context switches, the infinite-coordinate flag.
int raster_emu_frame(const void* geometry, size_t geometry_bytes, const float* substances, const uint8_t* picture_bytes,
                     const gudni_picture_use* picture_uses, const float* background, int width, int height,
                     const gudni_spec* spec, const gudni_shape* shapes, int64_t n_shapes, const gudni_tile* tiles,
                     const int32_t* thread_base, int n_tiles, int64_t n_columns, uint32_t* out, int32_t* dbg_thresholds,
                     int32_t* dbg_shape_bits, int64_t* stats) {
    FrameInputs in{geometry, geometry_bytes, substances, picture_bytes, picture_uses, background, width, height, spec};
    rasterStage(in, shapes, n_shapes, shapes, (int)sizeof(gudni_shape), n_shapes, tiles, thread_base, n_tiles, n_columns, out,
                dbg_thresholds, dbg_shape_bits, stats);
    return 0;
}

// Levels 2 and 3: tile binning (and, with `raw_shapes`, strand building) through the emulated kernels as well.
// Level 2: `entries` + geometry given.  Level 3: raw outlines given, geometry/entries built and returned through
// geometry_out / entries_out (capacities in bytes / records).  Binned tiles and shape lists come back through
// tiles_out / shapes_out for comparison with the oracle's tile tree.  sizes[0..4] = entries, geometry bytes,
// tiles, shape refs, columns.  Returns 0, or 1 if the binning scratch overflowed.
int raster_emu_scene(const void* geometry, size_t geometry_bytes, const gudni_shape_entry* entries, int n_entries,
                     const gudni_outline_shape* raw_shapes, int n_raw_shapes, const gudni_outline* outlines,
                     const gudni_curve_pair* pairs, const gudni_transform* transforms,
                     const float* substances, const uint8_t* picture_bytes, const gudni_picture_use* picture_uses,
                     const float* background, int width, int height, const gudni_spec* spec, uint32_t* out,
                     uint8_t* geometry_out, size_t geometry_capacity, gudni_shape_entry* entries_out, int64_t entry_capacity,
                     gudni_tile* tiles_out, int64_t tile_capacity, gudni_shape* shapes_out, int64_t shape_capacity,
                     int32_t* dbg_thresholds, int32_t* dbg_shape_bits, int64_t column_capacity, int64_t* sizes, int64_t* stats) {
    std::vector<uint8_t> builtGeometry;
    std::vector<gudni_shape_entry> builtEntries;
    if (raw_shapes) {
        using namespace gudni_strands;
        buildReorderTable(cTable);
        const int n = n_raw_shapes, blocks = (n + kBlock - 1) / kBlock;
        std::vector<ShapeMeasure> measures(std::max(n, 1));
        std::vector<BlockSums> sums(std::max(blocks, 1));
        unsigned long long totals[3] = {0, 0, 0};
        if (n) {
            cuemu::launch(strand_measure_kernel, dim3(blocks), dim3(kBlock), raw_shapes, n, outlines, pairs, transforms, width,
                          height, measures.data(), sums.data());
            cuemu::launch(strand_scan_kernel, dim3(1), dim3(kScanThreads), sums.data(), blocks, totals);
        }
        builtGeometry.assign((size_t)totals[1] * 16 + 16, 0xCD);
        builtEntries.assign(totals[0] + 1, gudni_shape_entry{});
        if (totals[0])
            cuemu::launch(strand_emit_kernel, dim3(blocks), dim3(kBlock), raw_shapes, n, outlines, pairs, transforms,
                          (const ShapeMeasure*)measures.data(), (const BlockSums*)sums.data(), builtGeometry.data(),
                          builtEntries.data());
        builtGeometry.resize((size_t)totals[1] * 16);
        builtEntries.resize(totals[0]);
        geometry = builtGeometry.data();
        geometry_bytes = builtGeometry.size();
        entries = builtEntries.data();
        n_entries = (int)builtEntries.size();
        if (geometry_out && geometry_capacity >= geometry_bytes && geometry_bytes) memcpy(geometry_out, geometry, geometry_bytes);
        if (entries_out && entry_capacity >= n_entries && n_entries) memcpy(entries_out, entries, (size_t)n_entries * sizeof(gudni_shape_entry));
    }
    static const uint8_t emptyGeometry[16] = {0};
    static const gudni_shape_entry noEntry{};
    if (!geometry) geometry = emptyGeometry;
    if (!entries) entries = &noEntry;
    FrameInputs in{geometry, geometry_bytes, substances, picture_bytes, picture_uses, background, width, height, spec};
    Binned b;
    binStage(in, entries, n_entries, b);
    sizes[0] = n_entries; sizes[1] = (int64_t)geometry_bytes; sizes[2] = (int64_t)b.tiles.size();
    sizes[3] = (int64_t)b.shapes.size(); sizes[4] = b.nColumns;
    if (b.overflow) return 1;
    if (tiles_out && tile_capacity >= (int64_t)b.tiles.size() && !b.tiles.empty()) memcpy(tiles_out, b.tiles.data(), b.tiles.size() * sizeof(gudni_tile));
    if (shapes_out && shape_capacity >= (int64_t)b.shapes.size() && !b.shapes.empty()) memcpy(shapes_out, b.shapes.data(), b.shapes.size() * sizeof(gudni_shape));
    const bool taps = dbg_thresholds && dbg_shape_bits && column_capacity >= b.nColumns;
    static const gudni_shape noShape{};
    rasterStage(in, b.shapes.empty() ? &noShape : b.shapes.data(), (int64_t)b.shapes.size(), entries, (int)sizeof(gudni_shape_entry),
                n_entries, b.tiles.data(), b.threadBase.data(), (int)b.tiles.size(), b.nColumns, out,
                taps ? dbg_thresholds : nullptr, taps ? dbg_shape_bits : nullptr, stats);
    return 0;
}

}  // extern "C"
