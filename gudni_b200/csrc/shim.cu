// shim.cu — the C ABI of include/gudni_b200.h.  Each entry point names the reference function it
// stands in for (paths relative to /root/reference/src/Graphics/Gudni/).
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "binning.cuh"
#include "context.cuh"
#include "raster_kernels.cuh"
#include "strands.cuh"

using gudni_dev::FrameParams;

namespace {

constexpr int kSpillListCapacity = 1 << 22;   // column-threads that may leave the on-chip queue per frame
constexpr int kSpillSlots = 148 * 128;        // HBM queue slots of the replay kernel (one CTA of 128 per SM) to start with
constexpr int kSpillSlotsMax = 148 * 128 * 8; // ... and at most (a frame that replays many threads: see frame_end)

bool isPow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
int log2i(int x) { int d = 0; while ((1 << d) < x) d++; return d; }

FrameParams makeParams(gudni_ctx* ctx) {
    FrameParams P{};
    P.geometry = static_cast<const uint8_t*>(ctx->geometryPtr);
    P.shapes = ctx->shapes.as<gudni_shape>();
    P.tiles = ctx->tiles.as<gudni_tile>();
    P.tileThreadBase = ctx->tileThreadBase.as<int32_t>();
    P.substances = static_cast<const float4*>(ctx->substancesPtr);
    P.pictureData = static_cast<const uint8_t*>(ctx->picturesPtr);
    P.pictureUses = static_cast<const gudni_picture_use*>(ctx->pictureUsesPtr);
    if (ctx->externalTarget) {
        P.out = static_cast<uint32_t*>(ctx->externalTarget);
        P.rowOrigin = ctx->externalRowOrigin;
    } else {
        P.out = ctx->frame.as<uint32_t>();
        P.rowOrigin = ctx->rowBegin;
    }
    P.background = make_float4(ctx->background[0], ctx->background[1], ctx->background[2], ctx->background[3]);
    P.width = ctx->width;
    P.height = ctx->height;
    P.rowBegin = ctx->rowBegin;
    P.rowEnd = ctx->rowEnd;
    P.computeDepth = ctx->computeDepth;
    P.maxShape = ctx->spec.max_shapes;
    P.maxThresholds = ctx->spec.max_thresholds;
    P.dbgThresholds = ctx->debug ? ctx->dbgThresholds.as<int32_t>() : nullptr;
    P.dbgShapeBits = ctx->debug ? ctx->dbgShapeBits.as<int32_t>() : nullptr;
    P.counters = ctx->counters.as<unsigned long long>();
    P.spillList = ctx->spillList.as<unsigned long long>();
    P.spillCapacity = ctx->spillCapacity;
    P.thrStore = ctx->thrStore.as<float4>();
    P.hdrStore = ctx->hdrStore.as<uint32_t>();
    P.storeCap = ctx->storeCap;
    P.threadRecs = ctx->threadRecs.as<gudni_dev::ThreadRec>();
    P.strandBounds = ctx->strandBounds.as<float2>();
    P.tileOrder = ctx->tileOrder.as<uint32_t>();
    P.streamPool = ctx->streamPool.as<uint2>();
    P.streamCapChunks = (unsigned int)std::min<unsigned long long>(ctx->streamCapChunks, 0xFFFFFFF0ull);
    P.stackKeys = ctx->stackKeys.as<ulonglong2>();
    P.stackColors = ctx->stackColors.as<float4>();
    P.refSlabs = ctx->refSlabs.as<uint2>();
    P.refCapSlabs = (unsigned int)std::min<unsigned long long>(ctx->refCapSlabs, (0x3FFFFFFFull / 128));
    return P;
}

int ensureFrameBuffer(gudni_ctx* ctx) {
    if (ctx->externalTarget) return GUDNI_OK;
    size_t bytes = (size_t)ctx->width * (size_t)(ctx->rowEnd - ctx->rowBegin) * 4;
    return devEnsure(ctx, ctx->frame, std::max<size_t>(bytes, 4));
}

// Hand-over buffers between the kernels.  Threshold store (generate -> sort -> slice): a first guess of 20 thresholds
// per column-thread or one per four pixels, whichever is more.  Section
// stream pool (slice -> colour): 6 chunks of 16 records per column-thread + one per 64 pixels.  Both first
// guesses are capped by a byte budget; afterwards each is sized 25 % above what the previous frame drew.  A frame
// that runs one of them dry hands the threads that found it empty to the replay kernel and frame_end then
// rasterizes the frame again with the measured demand (see retryExhausted), so an undersized guess costs time on
// the first frame, never pixels.  One 32-byte record per thread.
constexpr int kDefaultBatches = 0;
constexpr size_t kFirstGuessBudget = (size_t)6 << 30;
int ensureHandover(gudni_ctx* ctx, int64_t totalTiles) {
    const size_t threads = (size_t)totalTiles * (size_t)ctx->spec.threads_per_tile;
    const size_t pixels = (size_t)ctx->width * (size_t)(ctx->rowEnd - ctx->rowBegin);
    const size_t guess = std::min(std::max(threads * 20, pixels / 4), kFirstGuessBudget / 20);
    const size_t entries = std::max({guess, (size_t)(ctx->storeDemand + ctx->storeDemand / 4), (size_t)1 << 20});
    GUDNI_TRY(devEnsure(ctx, ctx->thrStore, entries * 16));
    GUDNI_TRY(devEnsure(ctx, ctx->hdrStore, entries * 4 + 64));   // (+64: the sort kernel's bulk copies of header words are rounded out to 16 bytes)
    ctx->storeCap = std::min(ctx->thrStore.cap / 16, (ctx->hdrStore.cap - 64) / 4);
    const size_t chunkGuess = std::min(threads * 6 + pixels / 64, kFirstGuessBudget / 128);
    const size_t chunks = std::max({chunkGuess, (size_t)(ctx->streamDemand + ctx->streamDemand / 4), (size_t)1 << 14});
    GUDNI_TRY(devEnsure(ctx, ctx->streamPool, chunks * 128));
    ctx->streamCapChunks = ctx->streamPool.cap / 128;
    // stack table: slabs of 128 numbers; first guess two slabs per (tile, 32-column group) + one number per 2 pixels
    const size_t slabGuess = std::min(threads / 16 + pixels / 256, kFirstGuessBudget / (128 * 32));
    const size_t slabs = std::max({slabGuess, (size_t)(ctx->refDemand + ctx->refDemand / 4), (size_t)1 << 10});
    GUDNI_TRY(devEnsure(ctx, ctx->stackKeys, slabs * 128 * 16));
    GUDNI_TRY(devEnsure(ctx, ctx->stackColors, slabs * 128 * 16));
    GUDNI_TRY(devEnsure(ctx, ctx->refSlabs, slabs * 8));
    ctx->refCapSlabs = std::min({ctx->stackKeys.cap / (128 * 16), ctx->stackColors.cap / (128 * 16), ctx->refSlabs.cap / 8});
    GUDNI_TRY(devEnsure(ctx, ctx->threadRecs, std::max<size_t>(threads, 32) * sizeof(gudni_dev::ThreadRec)));
    GUDNI_TRY(devEnsure(ctx, ctx->tileOrder, (size_t)std::max<int64_t>(totalTiles, 1) * 4, (size_t)ctx->rasteredTiles * 4));
    // one entry per unit (narrowest units: 8 column-threads) of a launch, and a tile of slack per batch
    GUDNI_TRY(devEnsure(ctx, ctx->wideList, (size_t)(totalTiles + gudni_dev::kMaxBatches) * (size_t)(ctx->spec.threads_per_tile / 8) * 4));
    return GUDNI_OK;
}

int ensureDebug(gudni_ctx* ctx, int64_t columnsBefore, int64_t columnsAfter) {
    if (!ctx->debug) return GUDNI_OK;
    GUDNI_TRY(devEnsure(ctx, ctx->dbgThresholds, (size_t)columnsAfter * 4, (size_t)columnsBefore * 4));
    GUDNI_TRY(devEnsure(ctx, ctx->dbgShapeBits, (size_t)columnsAfter * 4, (size_t)columnsBefore * 4));
    size_t n = (size_t)(columnsAfter - columnsBefore) * 4;
    GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->dbgThresholds.as<char>() + columnsBefore * 4, 0xFF, n, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->dbgShapeBits.as<char>() + columnsBefore * 4, 0xFF, n, ctx->stream));
    return GUDNI_OK;
}

void markFirstKernel(gudni_ctx* ctx) {
    if (!ctx->firstKernelRecorded) {
        cudaEventRecord(ctx->evFirstKernel, ctx->stream);
        ctx->firstKernelRecorded = true;
    }
}

int beginFrameCommon(gudni_ctx* ctx, const float bg[4], int width, int height, int frame) {
    if (width <= 0 || height <= 0) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "frame_begin: bad bitmap size %dx%d", width, height);
    std::memcpy(ctx->background, bg, 16);
    ctx->width = width;
    ctx->height = height;
    ctx->frameNumber = frame;
    ctx->rowBegin = 0;
    ctx->rowEnd = height;
    ctx->nShapes = ctx->nTiles = ctx->nColumns = 0;
    ctx->rasteredTiles = 0;
    ctx->rasteredShapes = 0;
    ctx->firstKernelRecorded = false;
    ctx->binUsed = false;
    ctx->strandsUsed = false;
    ctx->geometryDeferredSrc = nullptr;
    ctx->geometryPending = ctx->geometryTimed = false;
    ctx->inFrame = true;
    GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->counters.ptr, 0, 256, ctx->stream));
    return GUDNI_OK;
}

// host -> device on `stream`: pageable memory of some size through the context's staging ring and copy threads
// (hostcopy.cuh), everything else as it is
int copyIn(gudni_ctx* ctx, void* dev, const void* host, size_t bytes, cudaStream_t stream) {
    if (bytes >= HostCopier::kWorthIt && ctx->copyThreads > 0 && HostCopier::pageable(host) && ctx->copier.start(ctx->copyThreads)) {
        GUDNI_CUDA_TRY(ctx, ctx->copier.upload(dev, host, bytes, stream));
        return GUDNI_OK;
    }
    GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream));
    return GUDNI_OK;
}
// device -> host on `stream` (queued; for pageable memory: done when this returns)
int copyOut(gudni_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t stream) {
    if (bytes >= HostCopier::kWorthIt && ctx->copyThreads > 0 && HostCopier::pageable(host) && ctx->copier.start(ctx->copyThreads)) {
        GUDNI_CUDA_TRY(ctx, ctx->copier.download(host, dev, bytes, stream));
        return GUDNI_OK;
    }
    GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream));
    return GUDNI_OK;
}

// `generation` != 0: skip the copy if the buffer already holds `bytes` bytes uploaded under that generation
int uploadTo(gudni_ctx* ctx, DevBuf& buf, const void* src, size_t bytes, uint64_t generation = 0, cudaStream_t stream = nullptr,
             bool* copied = nullptr) {
    if (copied) *copied = false;
    if (generation != 0 && buf.ptr && buf.generation == generation && buf.bytesHeld == bytes) {
        ctx->uploadsSkipped++;
        return GUDNI_OK;
    }
    GUDNI_TRY(devEnsure(ctx, buf, std::max<size_t>(bytes, 16)));
    if (bytes) GUDNI_TRY(copyIn(ctx, buf.ptr, src, bytes, stream ? stream : ctx->stream));
    if (copied) *copied = bytes != 0;
    buf.generation = generation;
    buf.bytesHeld = bytes;
    return GUDNI_OK;
}

// The geometry heap is by far the largest input and the binning does not read it: it crosses PCIe on the copy stream
// while the entries are uploaded and binned on the main one, and the first kernel that walks strands waits for it here.
int waitGeometry(gudni_ctx* ctx) {
    if (ctx->geometryDeferredSrc) {      // frame_begin only noted it: the entries go first (one copy engine serves both)
        GUDNI_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->evFrameBegin, 0));   // (after whatever the caller's stream still held)
        GUDNI_TRY(copyIn(ctx, ctx->geometry.ptr, ctx->geometryDeferredSrc, ctx->geometryBytes, ctx->copyStream));
        GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evGeometryUp, ctx->copyStream));
        ctx->geometryDeferredSrc = nullptr;
        ctx->geometryPending = true;
        ctx->geometryTimed = true;
        return GUDNI_OK;                 // the copy is on its way; the next call makes the main stream wait for it
    }
    if (ctx->geometryPending) {
        GUDNI_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evGeometryUp, 0));
        ctx->geometryPending = false;
    }
    return GUDNI_OK;
}

// queue the deferred copy of the geometry heap, if any, without making the main stream wait for it yet
int startGeometry(gudni_ctx* ctx) { return ctx->geometryDeferredSrc ? waitGeometry(ctx) : GUDNI_OK; }
// ... and make the main stream wait for it
int needGeometry(gudni_ctx* ctx) {
    GUDNI_TRY(startGeometry(ctx));
    return waitGeometry(ctx);
}

}  // namespace

extern "C" {

// setupOpenCL + determineRasterSpec, OpenCL/Setup.hs:71-87, 102-147
int gudni_b200_init(int device, const gudni_spec* want, gudni_spec* got, gudni_ctx** out) {
    if (!out) return GUDNI_ERR_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return GUDNI_ERR_NO_DEVICE;
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return GUDNI_ERR_NO_DEVICE;
    }
    if (device >= count) return GUDNI_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GUDNI_ERR_NO_DEVICE;
    if (prop.major < 10) return GUDNI_ERR_NO_DEVICE;   // sm_100a code only; there is no fallback path
    gudni_ctx* ctx = new (std::nothrow) gudni_ctx();
    if (!ctx) return GUDNI_ERR_OOM;
    ctx->device = device;
    gudni_spec spec = {256, 256, 256, 1024, 1022, 127};   // canonical spec, BASELINE.md §3
    if (want) spec = *want;
    if (!isPow2(spec.max_tile_size) || !isPow2(spec.threads_per_tile) || spec.threads_per_tile < 32 ||
        spec.threads_per_tile > 1024 || spec.max_tile_size < 8 || spec.max_thresholds < 4 || spec.max_shapes < 1 ||
        spec.max_shapes > gudni_dev::kMaxShapeLimit || spec.max_tiles_per_call < 1 || spec.max_strands_per_tile < 1) {
        delete ctx;
        return GUDNI_ERR_ARGUMENT;
    }
    ctx->spec = spec;
    ctx->computeDepth = log2i(spec.threads_per_tile);
    if (got) *got = spec;
    auto fail = [&](int code) { gudni_b200_destroy(ctx); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(GUDNI_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking) != cudaSuccess) return fail(GUDNI_ERR_CUDA);
    ctx->stream = ctx->ownStream;
    if (cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess) return fail(GUDNI_ERR_CUDA);
    cudaEvent_t* evs[] = {&ctx->evFrameBegin, &ctx->evUploadDone, &ctx->evBinDone, &ctx->evRasterDone,
                          &ctx->evDownloadDone, &ctx->evFirstKernel, &ctx->evStrandsDone, &ctx->evGeometryUp};
    for (cudaEvent_t* e : evs)
        if (cudaEventCreate(e) != cudaSuccess) return fail(GUDNI_ERR_CUDA);
    if (gudni_launch::strandTableInit(ctx) != GUDNI_OK) return fail(GUDNI_ERR_CUDA);
    if (devEnsure(ctx, ctx->counters, gudni_dev::kCountersBytes) != GUDNI_OK) return fail(GUDNI_ERR_OOM);
    // batches per raster launch (rasterTiles): GUDNI_BATCHES overrides the default
    ctx->batches = kDefaultBatches;       // 0: rasterTiles decides (one batch, or two when the frame is stored into a peer's canvas)
    if (const char* e = std::getenv("GUDNI_BATCHES")) ctx->batches = std::max(1, std::min(atoi(e), gudni_dev::kMaxBatches));
    if (const char* e = std::getenv("GUDNI_BATCH_ORDERED")) ctx->batchOrdered = atoi(e) != 0 ? 1 : 0;
    if (const char* e = std::getenv("GUDNI_COPY_THREADS")) ctx->copyThreads = std::max(0, std::min(atoi(e), 16));
    if (const char* e = std::getenv("GUDNI_BATCH_SPLIT")) ctx->batchSplitPercent = std::max(0, std::min(atoi(e), 99));
    ctx->spillCapacity = kSpillListCapacity;
    if (devEnsure(ctx, ctx->spillList, (size_t)kSpillListCapacity * 8) != GUDNI_OK) return fail(GUDNI_ERR_OOM);
    ctx->spillSlots = kSpillSlots;
    if (devEnsure(ctx, ctx->spillThr, (size_t)kSpillSlots * spec.max_thresholds * 16) != GUDNI_OK) return fail(GUDNI_ERR_OOM);
    if (devEnsure(ctx, ctx->spillHdr, (size_t)kSpillSlots * spec.max_thresholds * 4) != GUDNI_OK) return fail(GUDNI_ERR_OOM);
    *out = ctx;
    return GUDNI_OK;
}

void gudni_b200_destroy(gudni_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->geometry, &ctx->substances, &ctx->pictures, &ctx->pictureUses, &ctx->shapes, &ctx->tiles,
                      &ctx->tileThreadBase, &ctx->frame, &ctx->counters, &ctx->spillList, &ctx->spillThr, &ctx->spillHdr,
                      &ctx->dbgThresholds, &ctx->dbgShapeBits, &ctx->entries, &ctx->binCounters, &ctx->thrStore, &ctx->hdrStore,
                      &ctx->threadRecs, &ctx->streamPool, &ctx->stackKeys, &ctx->stackColors, &ctx->refSlabs, &ctx->strandBounds, &ctx->tileOrder, &ctx->wideList, &ctx->olShapes, &ctx->olOutlines, &ctx->olPairs,
                      &ctx->olTransforms, &ctx->strandMeasures, &ctx->strandScan, &ctx->strandTotals};
    for (DevBuf* b : bufs)
        if (b->ptr) cudaFree(b->ptr);
    for (DevBuf& b : ctx->binWork)
        if (b.ptr) cudaFree(b.ptr);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaEvent_t evs[] = {ctx->evFrameBegin, ctx->evUploadDone, ctx->evBinDone, ctx->evRasterDone, ctx->evDownloadDone,
                         ctx->evFirstKernel, ctx->evStrandsDone, ctx->evGeometryUp};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    for (cudaStream_t st : ctx->batchStreams) cudaStreamDestroy(st);
    for (cudaEvent_t e : ctx->evJoin) cudaEventDestroy(e);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    delete ctx;
}

const char* gudni_b200_last_error(gudni_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// queueRasterJobs' frame-constant uploads, OpenCL/CallKernels.hs:229-235
int gudni_b200_frame_begin_cached(gudni_ctx* ctx, const void* geometry, size_t geometry_bytes, const float* substances,
                                  int n_substances, const uint8_t* picture_bytes, size_t n_picture_bytes,
                                  const gudni_picture_use* picture_uses, int n_picture_uses, const float background_rgba[4],
                                  int width, int height, int frame_number, const gudni_generations* gen) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if ((geometry_bytes && !geometry) || (n_substances && !substances) || (n_picture_bytes && !picture_bytes) ||
        (n_picture_uses && !picture_uses) || !background_rgba || n_substances < 0 || n_picture_uses < 0)
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "frame_begin: null or negative-sized input");
    if (n_substances >= (1 << 30)) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "frame_begin: more than 2^30 substances");
    const gudni_generations none{};
    if (!gen) gen = &none;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evFrameBegin, ctx->stream));
    GUDNI_TRY(beginFrameCommon(ctx, background_rgba, width, height, frame_number));
    // the geometry heap: sized (and looked up in the input cache) now, copied by startGeometry / waitGeometry
    ctx->geometryTimed = false;
    ctx->geometryPending = false;
    ctx->geometryDeferredSrc = nullptr;
    if (gen->geometry != 0 && ctx->geometry.ptr && ctx->geometry.generation == gen->geometry && ctx->geometry.bytesHeld == geometry_bytes) {
        ctx->uploadsSkipped++;
    } else {
        GUDNI_TRY(devEnsure(ctx, ctx->geometry, std::max<size_t>(geometry_bytes, 16)));
        if (geometry_bytes) ctx->geometryDeferredSrc = geometry;
        ctx->geometry.generation = gen->geometry;
        ctx->geometry.bytesHeld = geometry_bytes;
    }
    GUDNI_TRY(uploadTo(ctx, ctx->substances, substances, (size_t)n_substances * 16, gen->substances));
    GUDNI_TRY(uploadTo(ctx, ctx->pictures, picture_bytes, n_picture_bytes, gen->pictures));
    GUDNI_TRY(uploadTo(ctx, ctx->pictureUses, picture_uses, (size_t)n_picture_uses * sizeof(gudni_picture_use), gen->picture_uses));
    ctx->geometryPtr = ctx->geometry.ptr;
    ctx->substancesPtr = ctx->substances.ptr;
    ctx->picturesPtr = ctx->pictures.ptr;
    ctx->pictureUsesPtr = ctx->pictureUses.ptr;
    ctx->geometryBytes = geometry_bytes;
    ctx->pictureBytes = n_picture_bytes;
    ctx->nSubstances = n_substances;
    ctx->nPictureUses = n_picture_uses;
    GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evUploadDone, ctx->stream));
    return GUDNI_OK;
}

int gudni_b200_frame_begin(gudni_ctx* ctx, const void* geometry, size_t geometry_bytes, const float* substances,
                           int n_substances, const uint8_t* picture_bytes, size_t n_picture_bytes,
                           const gudni_picture_use* picture_uses, int n_picture_uses, const float background_rgba[4],
                           int width, int height, int frame_number) {
    return gudni_b200_frame_begin_cached(ctx, geometry, geometry_bytes, substances, n_substances, picture_bytes, n_picture_bytes,
                                         picture_uses, n_picture_uses, background_rgba, width, height, frame_number, nullptr);
}

int gudni_b200_frame_begin_device(gudni_ctx* ctx, const void* dev_geometry, size_t geometry_bytes,
                                  const void* dev_substances, int n_substances, const void* dev_picture_bytes,
                                  size_t n_picture_bytes, const void* dev_picture_uses, int n_picture_uses,
                                  const float background_rgba[4], int width, int height, int frame_number) {
    if (!ctx || !background_rgba) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evFrameBegin, ctx->stream));
    GUDNI_TRY(beginFrameCommon(ctx, background_rgba, width, height, frame_number));
    ctx->geometryPtr = dev_geometry;
    ctx->substancesPtr = dev_substances;
    ctx->picturesPtr = dev_picture_bytes;
    ctx->pictureUsesPtr = dev_picture_uses;
    ctx->geometryBytes = geometry_bytes;
    ctx->pictureBytes = n_picture_bytes;
    ctx->nSubstances = n_substances;
    ctx->nPictureUses = n_picture_uses;
    GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evUploadDone, ctx->stream));
    return GUDNI_OK;
}

int gudni_b200_frame_strip(gudni_ctx* ctx, int row_begin, int row_end) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "frame_strip outside a frame");
    if (ctx->nTiles) return ctxFail(ctx, GUDNI_ERR_STATE, "frame_strip after raster calls");
    const int tileRows = ctx->spec.max_tile_size;
    if (row_begin < 0 || row_end > ctx->height || row_begin >= row_end || (row_begin % tileRows) != 0 ||
        (row_end != ctx->height && (row_end % tileRows) != 0))
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "frame_strip: rows [%d,%d) are not whole root-tile rows of %d", row_begin,
                       row_end, tileRows);
    ctx->rowBegin = row_begin;
    ctx->rowEnd = row_end;
    return GUDNI_OK;
}

// rasterize the tiles of the jobs queued since the last launch (level 1)
constexpr int64_t kLaunchTiles = 2048;
static int launchPendingJobs(gudni_ctx* ctx) {
    const int64_t pending = ctx->nTiles - ctx->rasteredTiles;
    if (pending <= 0) return GUDNI_OK;
    GUDNI_TRY(ensureHandover(ctx, ctx->nTiles));
    markFirstKernel(ctx);
    GUDNI_TRY(devEnsure(ctx, ctx->strandBounds, ctx->geometryBytes / 2 + 16));
    GUDNI_TRY(needGeometry(ctx));
    GUDNI_TRY(gudni_launch::strandBounds(ctx, ctx->geometryPtr, ctx->shapes.as<gudni_shape>() + ctx->rasteredShapes,
                                         (int)sizeof(gudni_shape), (int)(ctx->nShapes - ctx->rasteredShapes),
                                         ctx->strandBounds.as<float2>()));
    GUDNI_TRY(gudni_launch::rasterTiles(ctx, makeParams(ctx), (int)ctx->rasteredTiles, (int)pending));
    ctx->rasteredTiles = ctx->nTiles;
    ctx->rasteredShapes = ctx->nShapes;
    return GUDNI_OK;
}

// raster + generateCall, OpenCL/CallKernels.hs:88-206
int gudni_b200_raster_job(gudni_ctx* ctx, const gudni_shape* shapes, int n_shapes, const gudni_tile* tiles, int n_tiles,
                          int columns_allocated, int job_index) {
    (void)job_index;   // only feeds the inert random field in the reference (K.cl:1713)
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_job outside a frame");
    if (n_shapes < 0 || n_tiles < 0 || (n_shapes && !shapes) || (n_tiles && !tiles) || columns_allocated < 0)
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_job: null or negative-sized input");
    if (n_tiles == 0) return GUDNI_OK;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int G = ctx->spec.threads_per_tile;
    // rebase the job onto the frame-wide arrays; validate what the kernels index with
    std::vector<gudni_tile> t(tiles, tiles + n_tiles);
    std::vector<int32_t> base(n_tiles);
    for (int i = 0; i < n_tiles; i++) {
        const gudni_tile& ti = tiles[i];
        if ((uint64_t)ti.shape_start + ti.shape_count > (uint64_t)n_shapes)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_job: tile %d shape slice out of range", i);
        if (ti.column_allocation < 0 || (int64_t)ti.column_allocation + G > (int64_t)columns_allocated)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_job: tile %d column allocation out of range", i);
        if (ti.shape_count > 65535u)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_job: tile %d lists more than 65535 shapes", i);
        if (ti.h_depth < 0 || ti.h_depth > 15 || ti.v_depth < 0 || ti.v_depth > 15)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_job: tile %d depth out of range", i);
        t[i].shape_start = ti.shape_start + (uint32_t)ctx->nShapes;
        base[i] = (int32_t)(ctx->nColumns + ti.column_allocation);
    }
    GUDNI_TRY(devEnsure(ctx, ctx->shapes, (size_t)(ctx->nShapes + n_shapes) * 16 + 16, (size_t)ctx->nShapes * 16));
    GUDNI_TRY(devEnsure(ctx, ctx->tiles, (size_t)(ctx->nTiles + n_tiles) * 32, (size_t)ctx->nTiles * 32));
    GUDNI_TRY(devEnsure(ctx, ctx->tileThreadBase, (size_t)(ctx->nTiles + n_tiles) * 4, (size_t)ctx->nTiles * 4));
    GUDNI_TRY(ensureFrameBuffer(ctx));
    GUDNI_TRY(ensureDebug(ctx, ctx->nColumns, ctx->nColumns + columns_allocated));
    if (n_shapes)
        GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->shapes.as<gudni_shape>() + ctx->nShapes, shapes, (size_t)n_shapes * 16,
                                            cudaMemcpyHostToDevice, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tiles.as<gudni_tile>() + ctx->nTiles, t.data(), (size_t)n_tiles * 32,
                                        cudaMemcpyHostToDevice, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tileThreadBase.as<int32_t>() + ctx->nTiles, base.data(), (size_t)n_tiles * 4,
                                        cudaMemcpyHostToDevice, ctx->stream));
    ctx->nShapes += n_shapes;
    ctx->nTiles += n_tiles;
    ctx->nColumns += columns_allocated;
    // The reference launches its kernels once per job (<= G tiles); a job that small leaves most of a
    // B200 idle, so jobs are collected and rasterized together, kLaunchTiles at a time or at frame_end
    // (nothing reads the bitmap before frame_end).
    if (ctx->nTiles - ctx->rasteredTiles >= kLaunchTiles) GUDNI_TRY(launchPendingJobs(ctx));
    return GUDNI_OK;
}

static int rasterSceneCommon(gudni_ctx* ctx, const void* devEntries, int n_entries) {
    GUDNI_TRY(ensureFrameBuffer(ctx));
    markFirstKernel(ctx);
    GUDNI_TRY(devEnsure(ctx, ctx->strandBounds, ctx->geometryBytes / 2 + 16));
    // binning first: it reads the entries only, so it runs while the geometry heap is still on its way (waitGeometry)
    // (the copy of the geometry heap is queued once the first binning kernels are: startGeometry)
    GUDNI_TRY(gudni_bin::binScene(ctx, static_cast<const gudni_shape_entry*>(devEntries), n_entries, startGeometry));
    GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evBinDone, ctx->stream));
    GUDNI_TRY(needGeometry(ctx));
    GUDNI_TRY(gudni_launch::strandBounds(ctx, ctx->geometryPtr, devEntries, (int)sizeof(gudni_shape_entry), n_entries,
                                         ctx->strandBounds.as<float2>()));
    GUDNI_TRY(ensureDebug(ctx, 0, ctx->nColumns));
    GUDNI_TRY(ensureHandover(ctx, ctx->nTiles));
    GUDNI_TRY(gudni_launch::rasterTiles(ctx, makeParams(ctx), 0, (int)ctx->nTiles));
    ctx->rasteredTiles = ctx->nTiles;
    return GUDNI_OK;
}

// buildTileTree/addShapeToTree + buildRasterJobs + the job loop of queueRasterJobs
int gudni_b200_raster_scene_cached(gudni_ctx* ctx, const gudni_shape_entry* entries, int n_entries, uint64_t generation) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_scene outside a frame");
    if (ctx->nTiles) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_scene after other raster calls of the frame");
    if (n_entries < 0 || (n_entries && !entries)) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_scene: bad entries");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_TRY(uploadTo(ctx, ctx->entries, entries, (size_t)n_entries * sizeof(gudni_shape_entry), generation));
    return rasterSceneCommon(ctx, ctx->entries.ptr, n_entries);
}
int gudni_b200_raster_scene(gudni_ctx* ctx, const gudni_shape_entry* entries, int n_entries) {
    return gudni_b200_raster_scene_cached(ctx, entries, n_entries, 0);
}

int gudni_b200_raster_scene_device(gudni_ctx* ctx, const void* dev_entries, int n_entries) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_scene outside a frame");
    if (ctx->nTiles) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_scene after other raster calls of the frame");
    if (n_entries < 0 || (n_entries && !dev_entries)) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_scene: bad entries");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return rasterSceneCommon(ctx, dev_entries, n_entries);
}

// onShape's geometry work + enclose + outlineToStrands (Raster/Serialize.hs:148-177, Enclosure.hs:62-73,
// Strand.hs:153-178), then level 2
static int rasterOutlinesCommon(gudni_ctx* ctx, const void* devShapes, int n_shapes, const void* devOutlines,
                                const void* devPairs, const void* devTransforms, int64_t inputBytes) {
    GUDNI_TRY(ensureFrameBuffer(ctx));
    markFirstKernel(ctx);
    GUDNI_TRY(needGeometry(ctx));   // (a heap handed to frame_begin is copied before this one replaces it, as it always was)
    GUDNI_TRY(gudni_launch::buildStrands(ctx, devShapes, n_shapes, devOutlines, devPairs, devTransforms));
    GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evStrandsDone, ctx->stream));
    ctx->strandsUsed = true;
    ctx->outlineInputBytes = inputBytes;
    return rasterSceneCommon(ctx, ctx->entries.ptr, ctx->nEntries);
}

int gudni_b200_raster_outlines(gudni_ctx* ctx, const gudni_outline_shape* shapes, int n_shapes, const gudni_outline* outlines,
                               int n_outlines, const gudni_curve_pair* pairs, int64_t n_pairs,
                               const gudni_transform* transforms, int n_transforms) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_outlines outside a frame");
    if (ctx->nTiles) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_outlines after other raster calls of the frame");
    if (n_shapes < 0 || n_outlines < 0 || n_pairs < 0 || n_transforms < 0 || (n_shapes && !shapes) ||
        (n_outlines && !outlines) || (n_pairs && !pairs) || (n_transforms && !transforms))
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: null or negative-sized input");
    if (n_pairs >= (1ll << 32)) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: more than 2^32 curve pairs");
    // the kernels index with these: check every slice once on the host
    for (int i = 0; i < n_outlines; i++)
        if ((uint64_t)outlines[i].first_pair + outlines[i].n_pairs > (uint64_t)n_pairs)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: outline %d runs past the pair array", i);
    for (int i = 0; i < n_shapes; i++) {
        if ((uint64_t)shapes[i].first_outline + shapes[i].n_outlines > (uint64_t)n_outlines)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: shape %d outline slice out of range", i);
        if ((uint64_t)shapes[i].first_transform + shapes[i].n_transforms > (uint64_t)n_transforms)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: shape %d transform slice out of range", i);
    }
    for (int i = 0; i < n_transforms; i++)
        if (transforms[i].kind > GUDNI_TRANSFORM_ROTATE)
            return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: transform %d has unknown kind %u", i, transforms[i].kind);
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_TRY(uploadTo(ctx, ctx->olShapes, shapes, (size_t)n_shapes * sizeof(gudni_outline_shape)));
    GUDNI_TRY(uploadTo(ctx, ctx->olOutlines, outlines, (size_t)n_outlines * sizeof(gudni_outline)));
    GUDNI_TRY(uploadTo(ctx, ctx->olPairs, pairs, (size_t)n_pairs * sizeof(gudni_curve_pair)));
    GUDNI_TRY(uploadTo(ctx, ctx->olTransforms, transforms, (size_t)n_transforms * sizeof(gudni_transform)));
    const int64_t inputBytes = (int64_t)n_shapes * 32 + (int64_t)n_outlines * 8 + n_pairs * 16 + (int64_t)n_transforms * 16;
    return rasterOutlinesCommon(ctx, ctx->olShapes.ptr, n_shapes, ctx->olOutlines.ptr, ctx->olPairs.ptr, ctx->olTransforms.ptr,
                                inputBytes);
}

int gudni_b200_raster_outlines_device(gudni_ctx* ctx, const void* dev_shapes, int n_shapes, const void* dev_outlines,
                                      int n_outlines, const void* dev_pairs, int64_t n_pairs, const void* dev_transforms,
                                      int n_transforms) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_outlines outside a frame");
    if (ctx->nTiles) return ctxFail(ctx, GUDNI_ERR_STATE, "raster_outlines after other raster calls of the frame");
    if (n_shapes < 0 || n_outlines < 0 || n_pairs < 0 || n_transforms < 0 || (n_shapes && (!dev_shapes || !dev_outlines || !dev_pairs)))
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "raster_outlines: null or negative-sized input");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t inputBytes = (int64_t)n_shapes * 32 + (int64_t)n_outlines * 8 + n_pairs * 16 + (int64_t)n_transforms * 16;
    return rasterOutlinesCommon(ctx, dev_shapes, n_shapes, dev_outlines, dev_pairs, dev_transforms, inputBytes);
}

int gudni_b200_debug_strands(gudni_ctx* ctx, void* geometry, size_t geometry_capacity, size_t* geometry_bytes,
                             gudni_shape_entry* entries, int64_t entry_capacity, int64_t* n_entries) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (geometry_bytes) *geometry_bytes = ctx->geometryBytes;
    if (n_entries) *n_entries = ctx->nEntries;
    GUDNI_TRY(needGeometry(ctx));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (geometry) {
        if (geometry_capacity < ctx->geometryBytes) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "debug_strands: geometry buffer too small");
        if (ctx->geometryBytes)
            GUDNI_CUDA_TRY(ctx, cudaMemcpy(geometry, ctx->geometryPtr, ctx->geometryBytes, cudaMemcpyDeviceToHost));
    }
    if (entries) {
        if (entry_capacity < ctx->nEntries) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "debug_strands: entry buffer too small");
        if (ctx->nEntries)
            GUDNI_CUDA_TRY(ctx, cudaMemcpy(entries, ctx->entries.ptr, (size_t)ctx->nEntries * sizeof(gudni_shape_entry),
                                           cudaMemcpyDeviceToHost));
    }
    return GUDNI_OK;
}

// OutputPtr read-back, OpenCL/Instances.hs:60-75
int gudni_b200_frame_end(gudni_ctx* ctx, uint32_t* out_bgra, gudni_stats* stats) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (!ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "frame_end outside a frame");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_TRY(ensureFrameBuffer(ctx));
    GUDNI_TRY(needGeometry(ctx));   // (a frame without raster calls: the caller's buffer must not be in use when this returns)
    if (ctx->nTiles) {
        GUDNI_TRY(launchPendingJobs(ctx));
        GUDNI_TRY(gudni_launch::rasterSpill(ctx, makeParams(ctx)));
    } else {
        markFirstKernel(ctx);
    }
    const size_t rows = (size_t)(ctx->rowEnd - ctx->rowBegin);
    unsigned long long counters[32] = {0};
    unsigned long long binCounters[8] = {0};
    for (int attempt = 0;; attempt++) {
        GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evRasterDone, ctx->stream));
        if (out_bgra && out_bgra != ctx->hostTarget) {   // (a host target already holds the pixels: frame_target_host)
            const uint32_t* src = ctx->externalTarget
                                      ? static_cast<const uint32_t*>(ctx->externalTarget) +
                                            (size_t)(ctx->rowBegin - ctx->externalRowOrigin) * ctx->width
                                      : ctx->frame.as<uint32_t>();
            GUDNI_TRY(copyOut(ctx, out_bgra, src, rows * ctx->width * 4, ctx->stream));
        }
        GUDNI_CUDA_TRY(ctx, cudaEventRecord(ctx->evDownloadDone, ctx->stream));
        GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(counters, ctx->counters.ptr, 256, cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->binUsed)
            GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(binCounters, ctx->binCounters.ptr, 64, cudaMemcpyDeviceToHost, ctx->stream));
        GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        // A frame that sends many threads to the lane-private replay (tiles over MAXSHAPE at the 8-pixel floor, queues over the
        // on-chip capacity) finds it four warps per SM wide: its HBM queues are per slot, and a context starts with few.  The
        // next frame gets a slot for every such thread, up to eight CTAs per SM (20,000 circles on 512 x 512: 646 -> 328 ms at
        // four CTAs per SM).
        if (counters[gudni_dev::kCntSpilled] > (unsigned long long)ctx->spillSlots && ctx->spillSlots < kSpillSlotsMax) {
            int slots = ctx->spillSlots;
            while (slots < kSpillSlotsMax && (unsigned long long)slots < counters[gudni_dev::kCntSpilled]) slots *= 2;
            if (devEnsure(ctx, ctx->spillThr, (size_t)slots * ctx->spec.max_thresholds * 16) == GUDNI_OK &&
                devEnsure(ctx, ctx->spillHdr, (size_t)slots * ctx->spec.max_thresholds * 4) == GUDNI_OK)
                ctx->spillSlots = slots;
            else
                ctx->err.clear();   // (no room: the replay stays as narrow as it was)
        }
        ctx->storeDemand = counters[gudni_dev::kCntStoreCursor];
        ctx->streamDemand = counters[gudni_dev::kCntStreamCursor];
        ctx->refDemand = counters[gudni_dev::kCntRefSlabs];
        // A per-frame buffer ran dry (threshold store, stream pool) or more threads were handed to the replay than
        // its list holds (those were dropped): the demand is known now, so size the buffers from it and rasterize
        // the frame's tiles again — everything the kernels read is still resident.  Once: a second failure is reported.
        const bool dropped = counters[gudni_dev::kCntSpilled] > (unsigned long long)ctx->spillCapacity;
        if (attempt > 0 || !ctx->nTiles || counters[gudni_dev::kCntNonFinite] || !(dropped || counters[gudni_dev::kCntExhausted])) break;
        if (dropped) {
            size_t cap = (size_t)ctx->spillCapacity;
            while (cap < counters[gudni_dev::kCntSpilled] + counters[gudni_dev::kCntSpilled] / 4) cap *= 2;
            GUDNI_TRY(devEnsure(ctx, ctx->spillList, cap * 8));
            ctx->spillCapacity = (int)std::min<size_t>(cap, (size_t)1 << 30);
        }
        GUDNI_TRY(ensureHandover(ctx, ctx->nTiles));
        GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->counters.ptr, 0, 256, ctx->stream));
        GUDNI_TRY(gudni_launch::rasterTiles(ctx, makeParams(ctx), 0, (int)ctx->nTiles));
        GUDNI_TRY(gudni_launch::rasterSpill(ctx, makeParams(ctx)));
        ctx->retriedFrames++;
    }
    ctx->inFrame = false;
#ifdef GUDNI_STATS
    fprintf(stderr, "[stats] records %llu ready-hits %llu pending-hits %llu new %llu slow %llu rounds %llu flushes %llu logged %llu\n",
            counters[8], counters[9], counters[10], counters[11], counters[12], counters[13], counters[14], counters[15]);
#endif
    if (counters[gudni_dev::kCntNonFinite] & 2ull)
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "a shape's strands run outside the geometry heap (geo_start, strand count or a strand's size word): nothing was rasterized");
    if (counters[gudni_dev::kCntNonFinite])
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "geometry holds a point at infinity: the reference's curve bisection does not terminate on it; nothing was rasterized");
    if (binCounters[4])
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "a minimum-size tile lists more than 65535 shapes: unsupported by the raster kernels");
    gudni_stats s{};
    s.n_tiles = ctx->nTiles;
    s.n_shape_refs = ctx->nShapes;
    s.n_thresholds = (int64_t)counters[gudni_dev::kCntThresholds];
    s.n_spilled_threads = (int64_t)counters[gudni_dev::kCntSpilled];
    // threads the replay list could not hold even after the retry: their pixels were not written
    const int64_t droppedThreads = std::max<int64_t>(0, s.n_spilled_threads - ctx->spillCapacity);
    s.n_overflow_threads = (int64_t)counters[gudni_dev::kCntOverflow];
    if (droppedThreads > 0)
        return ctxFail(ctx, GUDNI_ERR_OOM, "%lld column-threads could not be replayed (replay list full after one retry): their pixels are missing",
                       (long long)droppedThreads);
    s.algorithmic_bytes = (int64_t)ctx->geometryBytes + 16 * ctx->nShapes + 32 * ctx->nTiles + 16 * (int64_t)ctx->nSubstances +
                          24 * (int64_t)ctx->nPictureUses + (int64_t)ctx->pictureBytes + 4 * (int64_t)ctx->width * (int64_t)rows;
    cudaEventElapsedTime(&s.ms_upload, ctx->evFrameBegin, ctx->evUploadDone);
    if (ctx->geometryTimed) {       // the geometry heap went over the copy stream
        float g = 0.f;
        if (cudaEventElapsedTime(&g, ctx->evFrameBegin, ctx->evGeometryUp) == cudaSuccess) s.ms_upload = std::max(s.ms_upload, g);
    }
    cudaEventElapsedTime(&s.ms_raster, ctx->evFirstKernel, ctx->evRasterDone);
    cudaEventElapsedTime(&s.ms_download, ctx->evRasterDone, ctx->evDownloadDone);
    s.ms_bin = 0.f;
    s.ms_strands = 0.f;
    if (ctx->binUsed) {
        cudaEventElapsedTime(&s.ms_bin, ctx->evFirstKernel, ctx->evBinDone);
        s.ms_raster -= s.ms_bin;
    }
    if (ctx->strandsUsed) {       // level 3: first kernel .. strands done .. bin done .. raster done
        cudaEventElapsedTime(&s.ms_strands, ctx->evFirstKernel, ctx->evStrandsDone);
        s.ms_bin -= s.ms_strands;
        // the outline data is what crosses the boundary instead of the geometry heap it becomes
        s.algorithmic_bytes += ctx->outlineInputBytes + 32 * (int64_t)ctx->nEntries;
    }
    ctx->lastFrameMs = s.ms_raster + s.ms_bin + s.ms_strands;
    ctx->lastStats = s;
    if (stats) *stats = s;
    return GUDNI_OK;
}

int gudni_b200_frame_device_ptr(gudni_ctx* ctx, void** dev_bgra, size_t* n_bytes) {
    if (!ctx || !dev_bgra) return GUDNI_ERR_ARGUMENT;
    if (ctx->width <= 0) return ctxFail(ctx, GUDNI_ERR_STATE, "no frame yet");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_TRY(ensureFrameBuffer(ctx));
    *dev_bgra = ctx->externalTarget ? ctx->externalTarget : ctx->frame.ptr;
    if (n_bytes) *n_bytes = (size_t)ctx->width * (size_t)(ctx->rowEnd - ctx->rowBegin) * 4;
    return GUDNI_OK;
}

int gudni_b200_frame_target(gudni_ctx* ctx, void* dev_bgra, int row_origin) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (ctx->nTiles && ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "frame_target after raster calls");
    if (row_origin < 0) return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "frame_target: negative row origin");
    ctx->externalTarget = dev_bgra;
    ctx->externalRowOrigin = dev_bgra ? row_origin : 0;
    ctx->hostTarget = nullptr;
    return GUDNI_OK;
}

int gudni_b200_frame_target_host(gudni_ctx* ctx, uint32_t* host_bgra) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (ctx->nTiles && ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "frame_target_host after raster calls");
    ctx->hostTarget = nullptr;
    if (!host_bgra) return gudni_b200_frame_target(ctx, nullptr, 0);
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void* dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, host_bgra, 0) != cudaSuccess) {
        cudaGetLastError();
        return ctxFail(ctx, GUDNI_ERR_ARGUMENT, "frame_target_host: the bitmap is not page-locked (gudni_b200_host_register)");
    }
    GUDNI_TRY(gudni_b200_frame_target(ctx, dev, 0));
    ctx->hostTarget = host_bgra;
    return GUDNI_OK;
}

int gudni_b200_set_stream(gudni_ctx* ctx, void* cuda_stream) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    if (ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "set_stream inside a frame");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->ownStream;
    return GUDNI_OK;
}

int gudni_b200_ipc_export_frame(gudni_ctx* ctx, void* handle_64b) {
    if (!ctx || !handle_64b) return GUDNI_ERR_ARGUMENT;
    if (!ctx->frame.ptr) return ctxFail(ctx, GUDNI_ERR_STATE, "no frame buffer to export");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    GUDNI_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->frame.ptr));
    std::memcpy(handle_64b, &h, 64);
    return GUDNI_OK;
}
int gudni_b200_ipc_open(gudni_ctx* ctx, const void* handle_64b, void** dev_ptr) {
    if (!ctx || !handle_64b || !dev_ptr) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle_64b, 64);
    GUDNI_CUDA_TRY(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GUDNI_OK;
}
int gudni_b200_ipc_close(gudni_ctx* ctx, void* dev_ptr) {
    if (!ctx || !dev_ptr) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return GUDNI_OK;
}

int gudni_b200_device_alloc(gudni_ctx* ctx, size_t bytes, void** dev_ptr) {
    if (!ctx || !dev_ptr) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaMalloc(dev_ptr, std::max<size_t>(bytes, 16)));
    return GUDNI_OK;
}
int gudni_b200_device_free(gudni_ctx* ctx, void* dev_ptr) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaFree(dev_ptr));
    return GUDNI_OK;
}
int gudni_b200_upload(gudni_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes) {
    if (!ctx || (bytes && (!dev_dst || !host_src))) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (bytes) GUDNI_TRY(copyIn(ctx, dev_dst, host_src, bytes, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GUDNI_OK;
}
int gudni_b200_download(gudni_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    if (!ctx || (bytes && (!host_dst || !dev_src))) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (bytes) GUDNI_TRY(copyOut(ctx, host_dst, dev_src, bytes, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GUDNI_OK;
}
int gudni_b200_host_register(gudni_ctx* ctx, void* host_ptr, size_t bytes) {
    if (!ctx || !host_ptr || !bytes) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return GUDNI_OK;
}
int gudni_b200_host_unregister(gudni_ctx* ctx, void* host_ptr) {
    if (!ctx || !host_ptr) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaHostUnregister(host_ptr));
    return GUDNI_OK;
}
int gudni_b200_sync(gudni_ctx* ctx) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GUDNI_OK;
}
int gudni_b200_last_frame_ms(gudni_ctx* ctx, float* ms) {
    if (!ctx || !ms) return GUDNI_ERR_ARGUMENT;
    *ms = ctx->lastFrameMs;
    return GUDNI_OK;
}
int gudni_b200_launch_count(gudni_ctx* ctx, int64_t* n) {
    if (!ctx || !n) return GUDNI_ERR_ARGUMENT;
    *n = ctx->launches;
    return GUDNI_OK;
}

int gudni_b200_debug_selftest(gudni_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!ctx || !mismatches) return GUDNI_ERR_ARGUMENT;
    if (ctx->inFrame) return ctxFail(ctx, GUDNI_ERR_STATE, "selftest inside a frame");
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned long long* dev = ctx->counters.as<unsigned long long>() + 16;
    GUDNI_CUDA_TRY(ctx, cudaMemsetAsync(dev, 0, 8, ctx->stream));
    GUDNI_TRY(gudni_launch::selftestDiv3(ctx, n, seed, dev));
    unsigned long long host = 0;
    GUDNI_CUDA_TRY(ctx, cudaMemcpyAsync(&host, dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *mismatches = host;
    return GUDNI_OK;
}

int gudni_b200_debug_enable(gudni_ctx* ctx, int on) {
    if (!ctx) return GUDNI_ERR_ARGUMENT;
    ctx->debug = on != 0;
    return GUDNI_OK;
}
int gudni_b200_debug_thread_counts(gudni_ctx* ctx, int32_t* n_thresholds, int32_t* shape_bits, int64_t capacity,
                                   int64_t* n_threads) {
    if (!ctx || !n_threads) return GUDNI_ERR_ARGUMENT;
    *n_threads = ctx->nColumns;
    if (!ctx->debug) return ctxFail(ctx, GUDNI_ERR_STATE, "debug taps are not enabled");
    if (capacity < ctx->nColumns) return GUDNI_OK;   // caller sizes its buffers from *n_threads and calls again
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_thresholds && ctx->nColumns)
        GUDNI_CUDA_TRY(ctx, cudaMemcpy(n_thresholds, ctx->dbgThresholds.ptr, (size_t)ctx->nColumns * 4, cudaMemcpyDeviceToHost));
    if (shape_bits && ctx->nColumns)
        GUDNI_CUDA_TRY(ctx, cudaMemcpy(shape_bits, ctx->dbgShapeBits.ptr, (size_t)ctx->nColumns * 4, cudaMemcpyDeviceToHost));
    return GUDNI_OK;
}
int gudni_b200_debug_binned(gudni_ctx* ctx, gudni_tile* tiles, int64_t tile_capacity, int64_t* n_tiles, gudni_shape* shapes,
                            int64_t shape_capacity, int64_t* n_shapes) {
    if (!ctx || !n_tiles || !n_shapes) return GUDNI_ERR_ARGUMENT;
    *n_tiles = ctx->nTiles;
    *n_shapes = ctx->nShapes;
    GUDNI_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    GUDNI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (tiles && tile_capacity >= ctx->nTiles && ctx->nTiles)
        GUDNI_CUDA_TRY(ctx, cudaMemcpy(tiles, ctx->tiles.ptr, (size_t)ctx->nTiles * 32, cudaMemcpyDeviceToHost));
    if (shapes && shape_capacity >= ctx->nShapes && ctx->nShapes)
        GUDNI_CUDA_TRY(ctx, cudaMemcpy(shapes, ctx->shapes.ptr, (size_t)ctx->nShapes * 16, cudaMemcpyDeviceToHost));
    return GUDNI_OK;
}

}  // extern "C"
