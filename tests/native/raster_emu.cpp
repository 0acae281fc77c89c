// raster_emu.cpp — TEST HELPER: this repo's raster kernels (gudni_b200/csrc/raster_kernels.cu with
// raster_device.cuh / raster_warp.cuh, the very text nvcc compiles) built with g++ against a host stand-in for
// the CUDA device language (emu/cuda_runtime.h) and run under a cooperative SIMT emulator: every CUDA thread a
// fiber, warp collectives and __syncthreads as rendezvous points, one CTA at a time.  It is slow (small scenes
// only) and it is not a performance model; what it gives is the kernels' logic — queues, warp-cooperative
// sweep, colour cache, spill replay — checked against the oracle in the CPU test suite, deadlock detection for
// mismatched collectives, and a way to try a kernel change before a GPU is at hand.
// Division uses the plain IEEE `/` (-DGUDNI_NO_DIV3: div3 is bit-identical to it by construction and checked
// on the device by gudni_b200_debug_selftest).  Not product, not oracle.
#define GUDNI_HOST_EMULATION 1
#define GUDNI_NO_DIV3 1
#include <cuda_runtime.h>   // emu/cuda_runtime.h

namespace cuemu {
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
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = &S.scheduler;
        makecontext(&f.ctx, trampoline, 0);
    }
    size_t remaining = n;
    while (remaining) {
        S.progress = false;
        for (size_t i = 0; i < n; i++) {
            Fiber& f = S.fibers[i];
            if (f.done) continue;
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

#include <cstring>

using namespace gudni_dev;

extern "C" {

// One frame at level 1 of the ABI: the jobs' shapes and tiles laid end to end as the shim lays them
// (shape_start rebased, thread_base = first column-thread of each tile).  Returns 0; fills the image, the
// per-thread taps (may be null) and stats[0..3] = thresholds, spilled threads, overflowed threads, switches.
int raster_emu_frame(const void* geometry, size_t geometry_bytes, const float* substances, const uint8_t* picture_bytes,
                     const gudni_picture_use* picture_uses, const float* background, int width, int height,
                     const gudni_spec* spec, const gudni_shape* shapes, int64_t n_shapes, const gudni_tile* tiles,
                     const int32_t* thread_base, int n_tiles, int64_t n_columns, uint32_t* out, int32_t* dbg_thresholds,
                     int32_t* dbg_shape_bits, int64_t* stats) {
    FrameParams P{};
    int depth = 0;
    while ((1 << depth) < spec->threads_per_tile) depth++;
    std::vector<unsigned long long> counters(kCountersBytes / 8 + 8, 0ull);
    const int spillCapacity = 1 << 16, spillSlots = 128;
    std::vector<unsigned long long> spillList(spillCapacity);
    const size_t threads = (size_t)n_tiles * spec->threads_per_tile;
    const size_t entries = std::max<size_t>({threads * 24, (size_t)width * height / 4, (size_t)1 << 16});
    std::vector<float4> thrStore(entries);
    std::vector<uint32_t> hdrStore(entries);
    std::vector<ThreadRec> recs(std::max<size_t>(threads, 32));
    std::vector<uint32_t> order(std::max(n_tiles, 1));
    std::vector<float2> bounds(geometry_bytes / 16 + 2);
    std::vector<float4> spillThr((size_t)spillSlots * spec->max_thresholds);
    std::vector<uint32_t> spillHdr((size_t)spillSlots * spec->max_thresholds);
    P.geometry = static_cast<const uint8_t*>(geometry);
    P.shapes = shapes;
    P.tiles = tiles;
    P.tileThreadBase = thread_base;
    P.substances = reinterpret_cast<const float4*>(substances);
    P.pictureData = picture_bytes;
    P.pictureUses = picture_uses;
    P.out = out;
    P.background = make_float4(background[0], background[1], background[2], background[3]);
    P.width = width; P.height = height;
    P.rowBegin = 0; P.rowEnd = height; P.rowOrigin = 0;
    P.computeDepth = depth;
    P.maxShape = spec->max_shapes;
    P.maxThresholds = spec->max_thresholds;
    P.dbgThresholds = dbg_thresholds;
    P.dbgShapeBits = dbg_shape_bits;
    P.counters = counters.data();
    P.spillList = spillList.data();
    P.spillCapacity = spillCapacity;
    P.thrStore = thrStore.data();
    P.hdrStore = hdrStore.data();
    P.storeCap = entries;
    P.threadRecs = recs.data();
    P.strandBounds = bounds.data();
    P.tileOrder = order.data();
    P.numStreams = std::max(1, std::min(3, n_tiles));
    if (dbg_thresholds) for (int64_t i = 0; i < n_columns; i++) dbg_thresholds[i] = -1;
    if (dbg_shape_bits) for (int64_t i = 0; i < n_columns; i++) dbg_shape_bits[i] = -1;
    if (n_tiles > 0) {
        if (n_shapes > 0)
            cuemu::launch(strand_bounds_kernel, dim3((unsigned)((n_shapes + 255) / 256)), dim3(256), P.geometry,
                          reinterpret_cast<const uint8_t*>(shapes), (int)sizeof(gudni_shape), (int)n_shapes, bounds.data());
        cuemu::launch(tile_order_kernel, dim3(1), dim3(256), tiles, 0, n_tiles, order.data());
        cuemu::launch(raster_generate_kernel, dim3(2), dim3(kGenWarpsPerCta * 32), P, 0, n_tiles);
        cuemu::launch(raster_sweep_kernel, dim3(2), dim3(kSweepWarpsPerCta * 32), P, 0, n_tiles);
        cuemu::launch(raster_spill_kernel, dim3(1), dim3(spillSlots), P, spillThr.data(), spillHdr.data(), spillSlots);
    }
    if (stats) {
        stats[0] = (int64_t)counters[kCntThresholds];
        stats[1] = (int64_t)counters[kCntSpilled];
        stats[2] = (int64_t)counters[kCntOverflow];
        stats[3] = (int64_t)cuemu::S.switches;
    }
    return 0;
}

}  // extern "C"
