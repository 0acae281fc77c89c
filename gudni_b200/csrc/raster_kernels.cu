// raster_kernels.cu — the two raster kernels (fused generate+sort+sweep, and the HBM-queue replay
// of spilled threads) and their launch wrappers.  Device arithmetic is in raster_device.cuh.
#include "raster_kernels.cuh"

using namespace gudni_dev;

// Fused generate -> sort -> sweep: one CTA per tile, one thread per (column, slab) exactly as the
// reference's NDRange `Work2D numTiles threadsPerTile` with work-group [1, threadsPerTile]
// (OpenCL/CallKernels.hs:141-142), but a single launch covers every tile of the frame and the
// three phases never leave the SM.
template <int CAP>
__global__ void __launch_bounds__(1024) raster_tiles_kernel(const FrameParams P, int tileBase) {
    __shared__ unsigned long long sThresholds;
    __shared__ gudni_tile sTile;
    __shared__ TileTable sTable;
    const int tileIndex = tileBase + blockIdx.x;
    if (threadIdx.x == 0) {
        sThresholds = 0ull;
        sTile = P.tiles[tileIndex];
    }
    __syncthreads();
    fillTileTable(P, sTable, sTile.shape_start, sTile.shape_count);
    __syncthreads();
    const int column = threadIdx.x;
    const ThreadGeom g = threadGeom(P, sTile, column);
    if (g.active) {
        ChipQueue<CAP> q;
        const int threadId = P.tileThreadBase[tileIndex] + column;
        int generated;
        bool ok = rasterThread(P, sTable, min(sTile.shape_count, (uint32_t)kTileTableCap), g, q, threadId, generated);
        if (ok) {
            atomicAdd(&sThresholds, (unsigned long long)generated);
        } else {
            // replayed by raster_spill_kernel against an HBM queue of MAXTHRESHOLDS entries
            unsigned long long slot = atomicAdd(&P.counters[kCntSpilled], 1ull);
            if (slot < (unsigned long long)P.spillCapacity) P.spillList[slot] = ((unsigned long long)tileIndex << 32) | (unsigned long long)column;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sThresholds) atomicAdd(&P.counters[kCntThresholds], sThresholds);
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
    const int threads = ctx->spec.threads_per_tile;
    raster_tiles_kernel<kChipQueueCapacity><<<nTiles, threads, 0, ctx->stream>>>(P, tileBase);
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
