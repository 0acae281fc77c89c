// raster_kernels.cu — the two raster kernels (fused generate+sort+sweep, and the HBM-queue replay
// of spilled threads) and their launch wrappers.  Device arithmetic is in raster_device.cuh.
#include "raster_kernels.cuh"

#include <algorithm>

#include "raster_warp.cuh"

using namespace gudni_dev;

constexpr int kWarpsPerCta = 4;

// Fused generate -> sort -> sweep, persistent: the grid is sized to what the chip holds resident and
// every warp pulls (tile, 32-column group) units from a global counter until the frame is done, so
// a warp that draws a cheap unit moves on instead of waiting at a CTA barrier for the slowest warp
// of its tile, and the tail of the frame is spread over all SMs.  The unit is the reference's
// work-group sliced by warps: `Work2D numTiles threadsPerTile` (OpenCL/CallKernels.hs:141-142).
__global__ void __launch_bounds__(kWarpsPerCta * 32) raster_warps_kernel(const FrameParams P, int tileBase, int nTiles,
                                                                         unsigned int* workCounter) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smemRaw);
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch& W = scratch[warp];
    QueueCold<kQueueCap - kQueueHot> cold;
    LaneQueue q;
    q.cold = &cold;
    q.thrHot = W.qThr + lane;
    q.hdrHot = W.qHdr + lane;
    const int warpShift = P.computeDepth - 5;                    // warps per tile = threadsPerTile / 32
    const unsigned totalUnits = (unsigned)nTiles << warpShift;
    const uint32_t denseCap = (uint32_t)min(P.maxShape, kWarpTableCap);
    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd(workCounter, 1u);
        unit = __shfl_sync(full, unit, 0);
        if (unit >= totalUnits) break;
        const int tileIndex = tileBase + (int)(unit >> warpShift);
        const int column = (int)((unit & ((1u << warpShift) - 1u)) << 5) + lane;
        const gudni_tile tile = P.tiles[tileIndex];
        int generated = -1;
        int failed = 0;
        if (tile.shape_count <= denseCap) {
            failed = rasterWarpDense(P, W, q, tile, tileIndex, column, generated);
        } else {
            // a tile that stopped splitting at the 8-pixel floor with more shapes than stack bits:
            // lane-private sweep with the bit -> shape table, colours through global memory
            const ThreadGeom g = threadGeom(P, tile, column);
            if (g.active) {
                const bool ok = rasterThread(P, *reinterpret_cast<const TileTable*>(&W), 0u, g, q,
                                             P.tileThreadBase[tileIndex] + column, generated);
                failed = ok ? 0 : 1;
            }
        }
        __syncwarp();
        // statistics: thresholds of lanes that completed here (spilled lanes are counted by the replay)
        unsigned int mine = (!failed && generated > 0) ? (unsigned)generated : 0u;
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(full, mine, d);
        if (lane == 0 && mine) atomicAdd(&P.counters[kCntThresholds], (unsigned long long)mine);
        if (failed) {
            // replayed by raster_spill_kernel against an HBM queue of MAXTHRESHOLDS entries
            const unsigned long long slot = atomicAdd(&P.counters[kCntSpilled], 1ull);
            if (slot < (unsigned long long)P.spillCapacity)
                P.spillList[slot] = ((unsigned long long)tileIndex << 32) | (unsigned long long)column;
        }
    }
}

// Replay of spilled column-threads.  Persistent: each thread owns one HBM queue slot and walks the
// spill list with a grid stride, so the scratch footprint is fixed (slots x MAXTHRESHOLDS x 20 B)
// regardless of how many threads spilled.
__global__ void __launch_bounds__(128) raster_spill_kernel(const FrameParams P, float4* thr, uint32_t* hdr, int slots) {
    // spilled threads come from arbitrary tiles: the table is empty (count 0) and every layer
    // takes the global-memory path
    __shared__ TileTable sTable;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n = P.counters[kCntSpilled];
    if (n > (unsigned long long)P.spillCapacity) n = (unsigned long long)P.spillCapacity;
    for (unsigned long long i = slot; i < n; i += (unsigned long long)slots) {
        const unsigned long long packed = P.spillList[i];
        const int tileIndex = (int)(packed >> 32), column = (int)(packed & 0xFFFFFFFFull);
        const gudni_tile tile = P.tiles[tileIndex];
        const ThreadGeom g = threadGeom(P, tile, column);
        HbmQueue q;
        q.thr = thr + slot;
        q.hdr = hdr + slot;
        q.stride = (size_t)slots;
        q.cap = P.maxThresholds;
        // the first kernel does not count the thresholds of a thread it hands over
        int generated;
        bool ok = rasterThread(P, sTable, 0u, g, q, P.tileThreadBase[tileIndex] + column, generated);
        if (generated > 0) atomicAdd(&P.counters[kCntThresholds], (unsigned long long)generated);
        if (!ok) atomicAdd(&P.counters[kCntOverflow], 1ull);
    }
}

namespace gudni_launch {

int rasterTiles(gudni_ctx* ctx, const FrameParams& P, int tileBase, int nTiles) {
    if (nTiles <= 0) return GUDNI_OK;
    static int ctasPerSm = 0, numSms = 0;
    if (!ctasPerSm) {
        GUDNI_CUDA_TRY(ctx, cudaFuncSetAttribute(raster_warps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(kWarpsPerCta * sizeof(WarpScratch))));
        GUDNI_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, raster_warps_kernel, kWarpsPerCta * 32,
                                                                          kWarpsPerCta * sizeof(WarpScratch)));
        GUDNI_CUDA_TRY(ctx, cudaDeviceGetAttribute(&numSms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (ctasPerSm < 1) ctasPerSm = 1;
    }
    unsigned int* workCounter = reinterpret_cast<unsigned int*>(ctx->counters.as<unsigned long long>() + 3);
    GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(workCounter, 0, sizeof(unsigned int), ctx->stream));
    const long long units = (long long)nTiles * (ctx->spec.threads_per_tile / 32);
    const int grid = (int)std::min<long long>((long long)ctasPerSm * numSms, (units + kWarpsPerCta - 1) / kWarpsPerCta);
    raster_warps_kernel<<<grid, kWarpsPerCta * 32, kWarpsPerCta * sizeof(WarpScratch), ctx->stream>>>(P, tileBase, nTiles, workCounter);
    ctx->launches++;
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    return GUDNI_OK;
}

int rasterSpill(gudni_ctx* ctx, const FrameParams& P) {
    const int threads = 128;
    const int blocks = ctx->spillSlots / threads;
    raster_spill_kernel<<<blocks, threads, 0, ctx->stream>>>(P, ctx->spillThr.as<float4>(), ctx->spillHdr.as<uint32_t>(),
                                                             ctx->spillSlots);
    ctx->launches++;
    GUDNI_CUDA_TRY(ctx, cudaGetLastError());
    return GUDNI_OK;
}

}  // namespace gudni_launch
