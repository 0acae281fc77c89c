// strands.cu — kernels and launcher of level 3 (strand_build.cuh has the per-shape logic and the
// reference citations).  Three launches per frame:
//   strand_measure_kernel  thread / shape: transformed bounding box, canvas culling, strand count,
//                          16-byte units of geometry -> ShapeMeasure[] (24 B / shape), and per CTA of 256
//                          shapes the sums of (kept, units, strands)
//   strand_scan_kernel     one CTA: those per-CTA sums become exclusive prefixes; totals for the host
//                          (a first version scanned all shapes in this one CTA: 185 us of the 304 on S4)
//   strand_emit_kernel     thread / shape: CTA-local scan + the CTA's prefix give the entry index and the
//                          heap offset; gudni_shape_entry + the strands in tree order
// Bound: HBM in principle (S4: 6.4 MB in, 35 MB out); in this first version a thread writes its shape's
// 288 bytes alone, so the stores are 32-byte sectors rather than full lines.
#include "context.cuh"
#include "strand_build.cuh"
#include "strands.cuh"

namespace gudni_strands {

__constant__ ReorderTable cTable;

constexpr int kBlock = 256;          // shapes per CTA in the measure and emit kernels

struct BlockSums {                   // per CTA of the measure kernel; after the scan: exclusive prefixes
    uint32_t kept, strands;
    unsigned long long units;
};

// exclusive scan of (kept, units) over the CTA's threads; totals in *sumKept / *sumUnits.  All threads call it.
__device__ __forceinline__ void blockScan(uint32_t kept, unsigned long long units, uint32_t& exKept,
                                          unsigned long long& exUnits, uint32_t* sumKept, unsigned long long* sumUnits) {
    __shared__ uint32_t warpKept[32];
    __shared__ unsigned long long warpUnits[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = (blockDim.x + 31) >> 5;
    uint32_t incK = kept;
    unsigned long long incU = units;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t k2 = __shfl_up_sync(0xFFFFFFFFu, incK, d);
        const unsigned long long u2 = __shfl_up_sync(0xFFFFFFFFu, incU, d);
        if (lane >= d) { incK += k2; incU += u2; }
    }
    __syncthreads();                                   // the arrays may still be read from a previous call
    if (lane == 31) { warpKept[warp] = incK; warpUnits[warp] = incU; }
    __syncthreads();
    if (warp == 0) {
        uint32_t wk = lane < nWarps ? warpKept[lane] : 0u;
        unsigned long long wu = lane < nWarps ? warpUnits[lane] : 0ull;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t k2 = __shfl_up_sync(0xFFFFFFFFu, wk, d);
            const unsigned long long u2 = __shfl_up_sync(0xFFFFFFFFu, wu, d);
            if (lane >= d) { wk += k2; wu += u2; }
        }
        warpKept[lane] = wk;                           // inclusive over warps
        warpUnits[lane] = wu;
    }
    __syncthreads();
    exKept = (warp ? warpKept[warp - 1] : 0u) + (incK - kept);
    exUnits = (warp ? warpUnits[warp - 1] : 0ull) + (incU - units);
    if (sumKept) *sumKept = warpKept[nWarps - 1];
    if (sumUnits) *sumUnits = warpUnits[nWarps - 1];
}

__global__ void __launch_bounds__(kBlock) strand_measure_kernel(const gudni_outline_shape* shapes, int nShapes,
                                                                const gudni_outline* outlines, const gudni_curve_pair* pairs,
                                                                const gudni_transform* transforms, int width, int height,
                                                                ShapeMeasure* measures, BlockSums* blockSums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t kept = 0u, strands = 0u;
    unsigned long long units = 0ull;
    if (i < nShapes) {
        ShapeMeasure m = measureShape(shapes[i], outlines, pairs, transforms);
        if (culled(m, width, height)) m.units = 0xFFFFFFFFu;   // marks the shape as dropped
        else { kept = 1u; units = m.units; strands = m.strands; }
        measures[i] = m;
    }
    uint32_t exK, sumK;
    unsigned long long exU, sumU;
    blockScan(kept, units, exK, exU, &sumK, &sumU);
    // strands: only the CTA's total is needed
    __shared__ uint32_t strandTotal;
    if (threadIdx.x == 0) strandTotal = 0u;
    __syncthreads();
    for (int d = 16; d > 0; d >>= 1) strands += __shfl_down_sync(0xFFFFFFFFu, strands, d);
    if ((threadIdx.x & 31) == 0 && strands) atomicAdd(&strandTotal, strands);
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockSums b;
        b.kept = sumK; b.strands = strandTotal; b.units = sumU;
        blockSums[blockIdx.x] = b;
    }
}

constexpr int kScanThreads = 1024;

// One CTA: the per-CTA sums of the measure kernel become exclusive prefixes in place.
// totals: [0] kept shapes, [1] geometry units, [2] strands
__global__ void __launch_bounds__(kScanThreads) strand_scan_kernel(BlockSums* blockSums, int nBlocks, unsigned long long* totals) {
    __shared__ unsigned long long carryKept, carryUnits, carryStrands;
    if (threadIdx.x == 0) carryKept = carryUnits = carryStrands = 0ull;
    unsigned long long strandsMine = 0ull;
    __syncthreads();
    for (int base = 0; base < nBlocks; base += kScanThreads) {
        const int b = base + threadIdx.x;
        BlockSums in{0u, 0u, 0ull};
        if (b < nBlocks) in = blockSums[b];
        strandsMine += in.strands;
        uint32_t exK, sumK;
        unsigned long long exU, sumU;
        blockScan(in.kept, in.units, exK, exU, &sumK, &sumU);
        if (b < nBlocks) {
            BlockSums out;
            out.kept = (uint32_t)(carryKept + exK);
            out.strands = in.strands;
            out.units = carryUnits + exU;
            blockSums[b] = out;
        }
        __syncthreads();
        if (threadIdx.x == 0) { carryKept += sumK; carryUnits += sumU; }
        __syncthreads();
    }
    for (int d = 16; d > 0; d >>= 1) strandsMine += __shfl_down_sync(0xFFFFFFFFu, strandsMine, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&carryStrands, strandsMine);
    __syncthreads();
    if (threadIdx.x == 0) { totals[0] = carryKept; totals[1] = carryUnits; totals[2] = carryStrands; }
}

__global__ void __launch_bounds__(kBlock) strand_emit_kernel(const gudni_outline_shape* shapes, int nShapes,
                                                             const gudni_outline* outlines, const gudni_curve_pair* pairs,
                                                             const gudni_transform* transforms, const ShapeMeasure* measures,
                                                             const BlockSums* blockOffsets, uint8_t* geometry,
                                                             gudni_shape_entry* entries) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    ShapeMeasure m;
    m.units = 0xFFFFFFFFu;
    if (i < nShapes) m = measures[i];
    const bool keep = m.units != 0xFFFFFFFFu;
    uint32_t exK;
    unsigned long long exU;
    blockScan(keep ? 1u : 0u, keep ? (unsigned long long)m.units : 0ull, exK, exU, nullptr, nullptr);
    if (!keep) return;
    const BlockSums off = blockOffsets[blockIdx.x];
    const gudni_outline_shape s = shapes[i];
    const unsigned long long geoStart = off.units + exU;
    gudni_shape_entry e;
    e.tag = s.tag;
    e.geo_start = (uint32_t)geoStart;             // appendGeoRef counts in 16-byte units, Serialize.hs:76-85
    e.num_strands = m.strands;
    e.left = m.left; e.top = m.top; e.right = m.right; e.bottom = m.bottom;
    entries[off.kept + exK] = e;
    emitShape(s, outlines, pairs, transforms, &cTable, geometry, 16ull * geoStart);
}

}  // namespace gudni_strands

#ifndef GUDNI_HOST_EMULATION   // the emulator drives the kernels itself (tests/native/raster_emu.cpp)
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
    ctx->geometry.generation = ctx->entries.generation = 0;   // what the input cache knew about these buffers is gone
    GUDNI_TRY(devEnsure(ctx, ctx->entries, 32));
    ctx->geometryPtr = ctx->geometry.ptr;
    if (nShapes == 0) return GUDNI_OK;
    GUDNI_TRY(devEnsure(ctx, ctx->strandMeasures, (size_t)nShapes * sizeof(ShapeMeasure)));
    const int blocks = (nShapes + kBlock - 1) / kBlock;
    GUDNI_TRY(devEnsure(ctx, ctx->strandScan, (size_t)blocks * sizeof(BlockSums)));
    GUDNI_TRY(devEnsure(ctx, ctx->strandTotals, 32));
    GUDNI_TRY(devEnsure(ctx, ctx->entries, (size_t)nShapes * sizeof(gudni_shape_entry)));
    strand_measure_kernel<<<blocks, kBlock, 0, ctx->stream>>>(
        static_cast<const gudni_outline_shape*>(devShapes), nShapes, static_cast<const gudni_outline*>(devOutlines),
        static_cast<const gudni_curve_pair*>(devPairs), static_cast<const gudni_transform*>(devTransforms), ctx->width,
        ctx->height, ctx->strandMeasures.as<ShapeMeasure>(), ctx->strandScan.as<BlockSums>());
    ctx->launches++;
    strand_scan_kernel<<<1, kScanThreads, 0, ctx->stream>>>(ctx->strandScan.as<BlockSums>(), blocks,
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
        strand_emit_kernel<<<blocks, kBlock, 0, ctx->stream>>>(
            static_cast<const gudni_outline_shape*>(devShapes), nShapes, static_cast<const gudni_outline*>(devOutlines),
            static_cast<const gudni_curve_pair*>(devPairs), static_cast<const gudni_transform*>(devTransforms),
            ctx->strandMeasures.as<ShapeMeasure>(), ctx->strandScan.as<BlockSums>(), ctx->geometry.as<uint8_t>(),
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
#endif  // GUDNI_HOST_EMULATION
