// context.cuh — the rasterizer context behind the C ABI (include/gudni_b200.h): device buffers,
// stream, events, error text.  Replaces what `Rasterizer` / `OpenCLState` hold in the reference
// (OpenCL/Rasterizer.hs:52-60) and the per-job clCreateBuffer/clReleaseMemObject traffic of
// generateCall (OpenCL/CallKernels.hs:124-127, 175-178): every buffer here is allocated once,
// grown geometrically, and reused across jobs and frames.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/gudni_b200.h"
#ifndef GUDNI_HOST_EMULATION   // (the emulated kernels need nothing of the host side)
#include "hostcopy.cuh"
#endif

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    // input cache (gudni_b200_frame_begin_cached): what the buffer holds, as the caller named it
    uint64_t generation = 0;
    size_t bytesHeld = 0;
    template <class T>
    T* as() const { return static_cast<T*>(ptr); }
};

struct gudni_ctx {
    int device = 0;
    gudni_spec spec{};
    int computeDepth = 8;
    cudaStream_t stream = nullptr;      // stream in use (own or caller's)
    cudaStream_t ownStream = nullptr;
    cudaStream_t copyStream = nullptr;
    std::string err;
    int64_t launches = 0;

    // frame constants
    DevBuf geometry, substances, pictures, pictureUses;
    const void *geometryPtr = nullptr, *substancesPtr = nullptr, *picturesPtr = nullptr, *pictureUsesPtr = nullptr;
    size_t geometryBytes = 0, pictureBytes = 0;
    int nSubstances = 0, nPictureUses = 0;
    float background[4] = {0, 0, 0, 1};
    int width = 0, height = 0, frameNumber = 0;
    int rowBegin = 0, rowEnd = 0;
    bool inFrame = false;

    // per-frame tile / shape arrays (all jobs laid end to end)
    DevBuf shapes, tiles, tileThreadBase;
    int64_t nShapes = 0, nTiles = 0, nColumns = 0;
    int64_t rasteredTiles = 0;   // tiles already covered by a raster launch this frame
    int64_t rasteredShapes = 0;  // ... and the shape records of those tiles' jobs

    // output
    DevBuf frame;
    void* externalTarget = nullptr;   // gudni_b200_frame_target
    int externalRowOrigin = 0;
    const uint32_t* hostTarget = nullptr;   // gudni_b200_frame_target_host: the target is this page-locked host bitmap

    // counters / spill
    DevBuf counters;        // 8 x u64
    DevBuf spillList;       // u32 per spilled thread
    int spillCapacity = 0;
    DevBuf spillThr, spillHdr;
    DevBuf thrStore, hdrStore, threadRecs;   // generate -> sweep hand-over
    DevBuf tileOrder;                        // tiles of a launch by decreasing shape count
    DevBuf wideList;                         // units with a thread for raster_slice_wide_kernel (one region per batch of a launch)
    DevBuf strandBounds;                     // per-strand y range, geometry_bytes / 16 entries
    unsigned long long storeCap = 0;
    unsigned long long storeDemand = 0;   // thresholds the generate kernel wanted to store last frame
    DevBuf streamPool;                    // slice -> colour hand-over: section streams, 128-byte chunks
    unsigned long long streamCapChunks = 0;
    unsigned long long streamDemand = 0;  // chunks the slice kernel drew last frame
    // resident CTAs per SM of the persistent kernels on this context's device
    bool occupancyKnown = false;
    int genCtasPerSm = 0, sortCtasPerSm = 0, sliceCtasPerSm = 0, colorCtasPerSm = 0, numSms = 0;
    int resolveCtasPerSm = 0, compositeCtasPerSm = 0, accumulateCtasPerSm = 0;
    DevBuf stackKeys, stackColors, refSlabs;   // the frame's table of distinct shape stacks
    unsigned long long refCapSlabs = 0;
    unsigned long long refDemand = 0;          // slabs of stack numbers drawn last frame
    int spillSlots = 0;
    // a launch's tiles go through the kernels in `batches` interleaved batches, each on its own stream (rasterTiles)
    int batches = 0;         // 0: chosen per launch (rasterTiles)
    int batchSplitPercent = 0;   // two ordered batches: the expensive batch's share of the tiles (0: half; GUDNI_BATCH_SPLIT)
    int batchOrdered = -1;   // batches are runs of the cost order (the first one the most expensive tiles) instead of interleaved; -1: chosen with `batches`
    std::vector<cudaStream_t> batchStreams;   // batches beyond the first
    std::vector<cudaEvent_t> evJoin;
    cudaEvent_t evFork = nullptr;
    int64_t uploadsSkipped = 0;           // input-cache hits (gudni_b200_frame_begin_cached)
    int64_t retriedFrames = 0;            // frames rasterized twice because a per-frame buffer was undersized

    // taps
    bool debug = false;
    DevBuf dbgThresholds, dbgShapeBits;

    // binning (level 2)
    DevBuf entries;
    const void* entriesPtr = nullptr;
    int nEntries = 0;
    bool binUsed = false;
    DevBuf binWork[8];
    DevBuf binCounters;

    // strand building (level 3)
    DevBuf olShapes, olOutlines, olPairs, olTransforms;   // uploaded outline data (host-pointer variant)
    DevBuf strandMeasures, strandScan, strandTotals;
    int64_t builtStrands = 0;
    int64_t outlineInputBytes = 0;
    bool strandsUsed = false;

    // pageable caller memory <-> device through page-locked staging and a few copy threads (hostcopy.cuh)
#ifndef GUDNI_HOST_EMULATION
    HostCopier copier;
#endif
    int copyThreads = 4;     // GUDNI_COPY_THREADS; 0: leave pageable memory to the driver
    // pinned staging for host transfers
    void* pinned = nullptr;
    size_t pinnedCap = 0;

    // timing
    cudaEvent_t evFrameBegin = nullptr, evUploadDone = nullptr, evBinDone = nullptr, evRasterDone = nullptr,
                evDownloadDone = nullptr, evFirstKernel = nullptr, evStrandsDone = nullptr, evGeometryUp = nullptr;
    const void* geometryDeferredSrc = nullptr;   // frame_begin's geometry heap, not copied yet (the entries go first)
    bool geometryPending = false;   // the geometry heap is still crossing PCIe on the copy stream (waitGeometry, shim.cu)
    bool geometryTimed = false;     // ... this frame uploaded it there
    bool firstKernelRecorded = false;
    float lastFrameMs = 0.f;
    gudni_stats lastStats{};
};

inline int ctxFail(gudni_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define GUDNI_CUDA_TRY(ctx, expr)                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return ctxFail((ctx), e_ == cudaErrorMemoryAllocation ? GUDNI_ERR_OOM : GUDNI_ERR_CUDA,      \
                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
    } while (0)

// Grow `b` to at least `bytes` (geometric growth).  Contents are preserved when `keep` > 0 bytes.
inline int devEnsure(gudni_ctx* ctx, DevBuf& b, size_t bytes, size_t keep = 0) {
    if (bytes <= b.cap) return GUDNI_OK;
    size_t ncap = b.cap ? b.cap : 4096;
    while (ncap < bytes) ncap *= 2;
    void* np = nullptr;
    GUDNI_CUDA_TRY(ctx, cudaMalloc(&np, ncap));
    if (keep && b.ptr) {
        GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(np, b.ptr, keep, cudaMemcpyDeviceToDevice, ctx->stream));
        GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    } else if (b.ptr) {
        GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = np;
    b.cap = ncap;
    return GUDNI_OK;
}

#define GUDNI_TRY(expr)            \
    do {                           \
        int rc_ = (expr);          \
        if (rc_ != GUDNI_OK) return rc_; \
    } while (0)
