// strands.cu — kernels and launcher of level 3 (strand_build.cuh has the per-shape logic and the
// reference citations).  Three launches per frame:
//   strand_measure_kernel  thread / shape: transformed bounding box, canvas culling, strand count,
//                          16-byte units of geometry                      -> ShapeMeasure[] (24 B / shape)
//   strand_scan_kernel     one CTA, tiles of 8,192 shapes: exclusive sums of (kept, units, strands)
//                          -> entry index and geometry offset per shape, totals for the host
//   strand_emit_kernel     thread / kept shape: gudni_shape_entry + the strands in tree order
// Bound: HBM in principle (S4: 6.4 MB in, 35 MB out); in this first version a thread writes its shape's
// 288 bytes alone, so the stores are 32-byte sectors rather than full lines.
#include "context.cuh"
#include "strand_build.cuh"
#include "strands.cuh"

namespace gudni_strands {

__constant__ ReorderTable cTable;

struct ScanOut {
    uint32_t entryIndex;     // position among the kept shapes, or 0xFFFFFFFF if culled
    uint32_t geoStart;       // 16-byte units
};

__global__ void __launch_bounds__(256) strand_measure_kernel(const gudni_outline_shape* shapes, int nShapes,
                                                             const gudni_outline* outlines, const gudni_curve_pair* pairs,
                                                             const gudni_transform* transforms, int width, int height,
                                                             ShapeMeasure* measures) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nShapes) return;
    ShapeMeasure m = measureShape(shapes[i], outlines, pairs, transforms);
    if (culled(m, width, height)) m.units = 0xFFFFFFFFu;   // marks the shape as dropped
    measures[i] = m;
}

constexpr int kScanThreads = 1024;
constexpr int kScanPerThread = 8;

// totals: [0] kept shapes, [1] geometry units, [2] strands
__global__ void __launch_bounds__(kScanThreads) strand_scan_kernel(const ShapeMeasure* measures, int nShapes, ScanOut* out,
                                                                   unsigned long long* totals) {
    __shared__ unsigned long long warpKept[32], warpUnits[32];
    __shared__ unsigned long long carryKept, carryUnits, carryStrands;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carryKept = carryUnits = carryStrands = 0ull;
    unsigned long long strandsMine = 0ull;
    __syncthreads();
    for (int base = 0; base < nShapes; base += kScanThreads * kScanPerThread) {
        const int first = base + threadIdx.x * kScanPerThread;
        uint32_t units[kScanPerThread];
        unsigned long long kept = 0ull, sumUnits = 0ull;
        for (int k = 0; k < kScanPerThread; k++) {
            const int i = first + k;
            units[k] = 0xFFFFFFFFu;
            if (i < nShapes) {
                units[k] = measures[i].units;
                if (units[k] != 0xFFFFFFFFu) { kept++; sumUnits += units[k]; strandsMine += measures[i].strands; }
            }
        }
        // exclusive scan of (kept, sumUnits) over the CTA: warp shuffle, then the warp totals
        unsigned long long incK = kept, incU = sumUnits;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long k2 = __shfl_up_sync(0xFFFFFFFFu, incK, d), u2 = __shfl_up_sync(0xFFFFFFFFu, incU, d);
            if (lane >= d) { incK += k2; incU += u2; }
        }
        if (lane == 31) { warpKept[warp] = incK; warpUnits[warp] = incU; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long wk = warpKept[lane], wu = warpUnits[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long k2 = __shfl_up_sync(0xFFFFFFFFu, wk, d), u2 = __shfl_up_sync(0xFFFFFFFFu, wu, d);
                if (lane >= d) { wk += k2; wu += u2; }
            }
            warpKept[lane] = wk;      // inclusive over warps
            warpUnits[lane] = wu;
        }
        __syncthreads();
        unsigned long long exK = carryKept + (warp ? warpKept[warp - 1] : 0ull) + (incK - kept);
        unsigned long long exU = carryUnits + (warp ? warpUnits[warp - 1] : 0ull) + (incU - sumUnits);
        for (int k = 0; k < kScanPerThread; k++) {
            const int i = first + k;
            if (i >= nShapes) break;
            ScanOut o;
            if (units[k] == 0xFFFFFFFFu) { o.entryIndex = 0xFFFFFFFFu; o.geoStart = 0u; }
            else { o.entryIndex = (uint32_t)exK; o.geoStart = (uint32_t)exU; exK++; exU += units[k]; }
            out[i] = o;
        }
        __syncthreads();
        if (threadIdx.x == 0) { carryKept += warpKept[31]; carryUnits += warpUnits[31]; }
        __syncthreads();
    }
    // strands: a plain CTA reduction
    for (int d = 16; d > 0; d >>= 1) strandsMine += __shfl_down_sync(0xFFFFFFFFu, strandsMine, d);
    if (lane == 0) atomicAdd(&carryStrands, strandsMine);
    __syncthreads();
    if (threadIdx.x == 0) { totals[0] = carryKept; totals[1] = carryUnits; totals[2] = carryStrands; }
}

__global__ void __launch_bounds__(256) strand_emit_kernel(const gudni_outline_shape* shapes, int nShapes,
                                                          const gudni_outline* outlines, const gudni_curve_pair* pairs,
                                                          const gudni_transform* transforms, const ShapeMeasure* measures,
                                                          const ScanOut* scan, uint8_t* geometry, gudni_shape_entry* entries) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nShapes) return;
    const ScanOut so = scan[i];
    if (so.entryIndex == 0xFFFFFFFFu) return;
    const gudni_outline_shape s = shapes[i];
    const ShapeMeasure m = measures[i];
    gudni_shape_entry e;
    e.tag = s.tag;
    e.geo_start = so.geoStart;                    // appendGeoRef counts in 16-byte units, Serialize.hs:76-85
    e.num_strands = m.strands;
    e.left = m.left; e.top = m.top; e.right = m.right; e.bottom = m.bottom;
    entries[so.entryIndex] = e;
    emitShape(s, outlines, pairs, transforms, &cTable, geometry, 16ull * (uint64_t)so.geoStart);
}

}  // namespace gudni_strands

namespace gudni_launch {

int strandTableInit(gudni_ctx* ctx) {
    gudni_strands::ReorderTable t;
    gudni_strands::buildReorderTable(t);
    GUDNI_CUDA_TRY(ctx, cudaMemcpyToSymbol(gudni_strands::cTable, &t, sizeof t));
    return GUDNI_OK;
}

// Builds ctx->geometry and ctx->entries from device-resident outline data.  Synchronises once (the
// geometry buffer is sized from the scan's totals).
int buildStrands(gudni_ctx* ctx, const void* devShapes, int nShapes, const void* devOutlines, const void* devPairs,
                 const void* devTransforms) {
    using namespace gudni_strands;
    ctx->nEntries = 0;
    ctx->geometryBytes = 0;
    ctx->geometryPtr = ctx->geometry.ptr;
    ctx->builtStrands = 0;
    // the later stages take these pointers even when there is nothing behind them
    GUDNI_TRY(devEnsure(ctx, ctx->geometry, 16));
    GUDNI_TRY(devEnsure(ctx, ctx->entries, 32));
    ctx->geometryPtr = ctx->geometry.ptr;
    if (nShapes == 0) return GUDNI_OK;
    GUDNI_TRY(devEnsure(ctx, ctx->strandMeasures, (size_t)nShapes * sizeof(ShapeMeasure)));
    GUDNI_TRY(devEnsure(ctx, ctx->strandScan, (size_t)nShapes * sizeof(ScanOut) + 32));
    GUDNI_TRY(devEnsure(ctx, ctx->strandTotals, 32));
    GUDNI_TRY(devEnsure(ctx, ctx->entries, (size_t)nShapes * sizeof(gudni_shape_entry)));
    const int blocks = (nShapes + 255) / 256;
    strand_measure_kernel<<<blocks, 256, 0, ctx->stream>>>(
        static_cast<const gudni_outline_shape*>(devShapes), nShapes, static_cast<const gudni_outline*>(devOutlines),
        static_cast<const gudni_curve_pair*>(devPairs), static_cast<const gudni_transform*>(devTransforms), ctx->width,
        ctx->height, ctx->strandMeasures.as<ShapeMeasure>());
    ctx->launches++;
    strand_scan_kernel<<<1, kScanThreads, 0, ctx->stream>>>(ctx->strandMeasures.as<ShapeMeasure>(), nShapes,
                                                            ctx->strandScan.as<ScanOut>(),
                                                            ctx->strandTotals.as<unsigned long long>());
    ctx->launches++;
    unsigned long long totals[3] = {0, 0, 0};
    GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(totals, ctx->strandTotals.ptr, sizeof totals, cudaMemcpyDeviceToHost, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (totals[1] >= (1ull << 32))
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: geometry heap would exceed 2^32 16-byte units");
    const size_t geoBytes = (size_t)totals[1] * 16;
    GUDNI_TRY(devEnsure(ctx, ctx->geometry, std::max<size_t>(geoBytes, 16)));
    if (totals[0]) {
        ctx->launches++;
        strand_emit_kernel<<<blocks, 256, 0, ctx->stream>>>(
            static_cast<const gudni_outline_shape*>(devShapes), nShapes, static_cast<const gudni_outline*>(devOutlines),
            static_cast<const gudni_curve_pair*>(devPairs), static_cast<const gudni_transform*>(devTransforms),
            ctx->strandMeasures.as<ShapeMeasure>(), ctx->strandScan.as<ScanOut>(), ctx->geometry.as<uint8_t>(),
            ctx->entries.as<gudni_shape_entry>());
    }
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    ctx->geometryPtr = ctx->geometry.ptr;
    ctx->geometryBytes = geoBytes;
    ctx->nEntries = (int)totals[0];
    ctx->builtStrands = (int64_t)totals[2];
    return GUDNI_OK;
}

}  // namespace gudni_launch
