// multi.cu — one frame on several GPUs of one box, behind the C ABI (include/gudni_b200.h, gudni_b200_multi_*).
//
// The reference has no counterpart: one OpenCLState is one device (OpenCL/Setup.hs:118-120) and queueRasterJobs
// (OpenCL/CallKernels.hs:218-242) feeds it every job.  Tiles and columns are independent, so the path shards with no
// data-path exchange: the canvas is cut into strips of whole root-tile rows, every device bins and rasterizes the
// shapes that touch its strip (its own gudni_ctx, its own host thread issuing its launches), and only finished pixels
// move:
//   * to the caller's host bitmap — every device copies its strip over its own PCIe link straight into the rows
//     of `out_bgra` it owns (no inter-GPU traffic at all for a host presenter), and/or
//   * to a canvas on the presenting device — every other device pushes its strip over NVLink
//     (cudaMemcpyPeerAsync with peer access enabled: one process, so no NCCL communicator is needed; the
//     one-process-per-GPU harness in gudni_b200/strips.py gathers with NCCL send/recv instead).
// The strips are re-cut from the measured per-device times of the previous frame (rebalanceRows), contiguous and
// in device order, so that the devices finish together.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "context.cuh"

namespace {

struct Strip {
    int rowBegin = 0, rowEnd = 0;
};

// Contiguous runs of root-tile rows per device, balanced by a cheap work estimate: shape-box area clipped to each
// tile row plus half the row's pixel count (covers empty rows).  First frame only; afterwards rebalanceRows.
std::vector<Strip> partitionRows(const gudni_shape_entry* e, int n, int width, int height, int tileRows, int nDevices) {
    const int nRows = (height + tileRows - 1) / tileRows;
    std::vector<double> weight(nRows, 0.0);
    for (int i = 0; i < n; i++) {
        const double top = std::min(std::max((double)e[i].top, 0.0), (double)height);
        const double bottom = std::min(std::max((double)e[i].bottom, 0.0), (double)height);
        const double w = std::min(std::max((double)e[i].right, 0.0), (double)width) - std::min(std::max((double)e[i].left, 0.0), (double)width);
        if (!(bottom > top) || !(w > 0.0)) continue;
        for (int r = (int)(top / tileRows); r < nRows && r * (double)tileRows < bottom; r++) {
            const double y0 = r * (double)tileRows, y1 = std::min((r + 1) * (double)tileRows, (double)height);
            weight[r] += std::max(std::min(bottom, y1) - std::max(top, y0), 0.0) * w;
        }
    }
    for (double& w : weight) w += (double)width * tileRows * 0.5;
    const int k = std::min(nDevices, nRows);
    std::vector<double> cum(nRows + 1, 0.0);
    for (int r = 0; r < nRows; r++) cum[r + 1] = cum[r] + weight[r];
    std::vector<int> bounds{0};
    for (int d = 1; d < k; d++) {
        const double target = cum[nRows] * d / k;
        int idx = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        idx = std::max(bounds.back() + 1, std::min(idx, nRows - (k - d)));
        bounds.push_back(idx);
    }
    bounds.push_back(nRows);
    std::vector<Strip> strips(nDevices);
    for (int d = 0; d < nDevices; d++) {
        if (d < k) strips[d] = Strip{bounds[d] * tileRows, std::min(bounds[d + 1] * tileRows, height)};
        else strips[d] = Strip{height, height};   // more devices than tile rows: idle
    }
    return strips;
}

// Feedback partition: the cost of a tile row is taken uniform inside last frame's strip and the canvas is cut
// again, contiguous and in device order, minimising the latest finish (exact, by dynamic programming over
// tile rows; a few dozen rows, at most 16 devices).
std::vector<Strip> rebalanceRows(const std::vector<Strip>& last, const std::vector<double>& ms, int height, int tileRows) {
    const int nRows = (height + tileRows - 1) / tileRows;
    const int n = (int)last.size();
    if (nRows < n) return last;
    std::vector<double> cost(nRows, 0.0);
    for (int d = 0; d < n; d++) {
        const int a = last[d].rowBegin / tileRows, b = (last[d].rowEnd + tileRows - 1) / tileRows;
        for (int r = a; r < b && r < nRows; r++) cost[r] = std::max(ms[d], 1e-3) / (b - a);
    }
    std::vector<double> pre(nRows + 1, 0.0);
    for (int r = 0; r < nRows; r++) pre[r + 1] = pre[r] + cost[r];
    const double inf = 1e300;
    std::vector<std::vector<double>> best(n + 1, std::vector<double>(nRows + 1, inf));
    std::vector<std::vector<int>> cut(n + 1, std::vector<int>(nRows + 1, 0));
    best[0][0] = 0.0;
    for (int k = 1; k <= n; k++)
        for (int j = k; j <= nRows - (n - k); j++)
            for (int i = k - 1; i < j; i++) {
                const double v = std::max(best[k - 1][i], pre[j] - pre[i]);
                if (v < best[k][j]) { best[k][j] = v; cut[k][j] = i; }
            }
    std::vector<int> bounds{nRows};
    for (int k = n; k > 0; k--) bounds.push_back(cut[k][bounds.back()]);
    std::reverse(bounds.begin(), bounds.end());
    std::vector<Strip> out(n);
    for (int d = 0; d < n; d++) out[d] = Strip{bounds[d] * tileRows, std::min(bounds[d + 1] * tileRows, height)};
    return out;
}

struct FrameJob {
    const void* geometry; size_t geometryBytes;
    const float* substances; int nSubstances;
    const uint8_t* pictures; size_t pictureBytes;
    const gudni_picture_use* uses; int nUses;
    float background[4];
    int width, height, frameNumber;
    const gudni_shape_entry* entries; int nEntries;
    gudni_generations generations;
    uint32_t* outHost;
};

}  // namespace

struct gudni_multi {
    struct Worker {
        gudni_multi* owner = nullptr;
        int index = 0, device = 0;
        gudni_ctx* ctx = nullptr;
        std::thread thread;
        // hand-off
        std::mutex m;
        std::condition_variable cv;
        int posted = 0, done = 0;
        bool quit = false;
        // per frame
        Strip strip;
        int rc = GUDNI_OK;
        std::string err;
        gudni_stats stats{};
        double finishedAt = 0.0, landedAt = 0.0;   // seconds since the frame began: raster done, strip where it belongs
        // the strip's entries and what they were cut from (input cache)
        std::vector<gudni_shape_entry> subset;
        Strip subsetStrip;
        uint64_t subsetGeneration = 0;
        int subsetOf = -1;
        void* stripBuf = nullptr;     // device strip when a presenting device gathers
        size_t stripCap = 0;
    };
    std::deque<Worker> workers;   // (a Worker holds a mutex: it never moves)
    gudni_spec spec{};
    std::string err;
    FrameJob job{};
    std::vector<Strip> strips;
    std::vector<double> lastMs;
    int lastWidth = 0, lastHeight = 0, lastEntries = -1;
    int presenting = -1;              // index of the device that holds the gathered canvas, -1: none
    void* canvas = nullptr;           // on workers[presenting].device
    size_t canvasCap = 0;
    std::chrono::steady_clock::time_point frameStart;
};

namespace {

double sinceStart(gudni_multi* mm) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - mm->frameStart).count();
}

void runFrame(gudni_multi::Worker& w) {
    gudni_multi* mm = w.owner;
    const FrameJob& j = mm->job;
    w.rc = GUDNI_OK;
    w.err.clear();
    w.stats = gudni_stats{};
    const Strip s = w.strip;
    if (s.rowEnd <= s.rowBegin) { w.finishedAt = w.landedAt = sinceStart(mm); return; }
    auto fail = [&](int rc) { w.rc = rc; w.err = gudni_b200_last_error(w.ctx); };
    // the strip's entries: shapes whose box touches its rows, scene order kept (what a strip's tile tree is built from)
    const bool sameCut = j.generations.entries != 0 && w.subsetGeneration == j.generations.entries &&
                         w.subsetStrip.rowBegin == s.rowBegin && w.subsetStrip.rowEnd == s.rowEnd && w.subsetOf == j.nEntries;
    if (!sameCut) {
        w.subset.clear();
        for (int i = 0; i < j.nEntries; i++)
            if (j.entries[i].top < (float)s.rowEnd && j.entries[i].bottom > (float)s.rowBegin) w.subset.push_back(j.entries[i]);
        w.subsetStrip = s;
        w.subsetGeneration = j.generations.entries;
        w.subsetOf = j.nEntries;
    }
    const bool gather = mm->presenting >= 0;
    const bool presenter = gather && mm->presenting == w.index;
    const size_t stripBytes = (size_t)(s.rowEnd - s.rowBegin) * j.width * 4;
    if (gather) {
        if (presenter) {
            if (int rc = gudni_b200_frame_target(w.ctx, mm->canvas, 0)) return fail(rc);
        } else {
            if (w.stripCap < stripBytes) {
                cudaSetDevice(w.device);
                if (w.stripBuf) cudaFree(w.stripBuf);
                w.stripBuf = nullptr; w.stripCap = 0;
                if (cudaMalloc(&w.stripBuf, stripBytes) != cudaSuccess) { w.rc = GUDNI_ERR_OOM; w.err = "strip buffer"; return; }
                w.stripCap = stripBytes;
            }
            if (int rc = gudni_b200_frame_target(w.ctx, w.stripBuf, s.rowBegin)) return fail(rc);
        }
    } else {
        if (int rc = gudni_b200_frame_target(w.ctx, nullptr, 0)) return fail(rc);
    }
    gudni_generations gen = j.generations;
    if (int rc = gudni_b200_frame_begin_cached(w.ctx, j.geometry, j.geometryBytes, j.substances, j.nSubstances, j.pictures, j.pictureBytes,
                                               j.uses, j.nUses, j.background, j.width, j.height, j.frameNumber, &gen))
        return fail(rc);
    if (int rc = gudni_b200_frame_strip(w.ctx, s.rowBegin, s.rowEnd)) return fail(rc);
    if (int rc = gudni_b200_raster_scene_cached(w.ctx, w.subset.data(), (int)w.subset.size(), sameCut ? j.generations.entries : 0))
        return fail(rc);
    // host presenter: this device's rows of the caller's bitmap, over this device's own PCIe link
    uint32_t* hostRows = j.outHost ? j.outHost + (size_t)s.rowBegin * j.width : nullptr;
    if (int rc = gudni_b200_frame_end(w.ctx, hostRows, &w.stats)) return fail(rc);
    w.finishedAt = sinceStart(mm);
    if (gather && !presenter) {   // device presenter: the strip goes over NVLink into the canvas rows it owns
        cudaSetDevice(w.device);
        uint8_t* dst = static_cast<uint8_t*>(mm->canvas) + (size_t)s.rowBegin * j.width * 4;
        cudaError_t e = cudaMemcpyPeerAsync(dst, mm->workers[mm->presenting].device, w.stripBuf, w.device, stripBytes, w.ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w.ctx->stream);
        if (e != cudaSuccess) { w.rc = GUDNI_ERR_CUDA; w.err = std::string("strip gather: ") + cudaGetErrorString(e); return; }
    }
    w.landedAt = sinceStart(mm);
}

void workerLoop(gudni_multi::Worker* w) {
    int seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->quit || w->posted != seen; });
            if (w->quit) return;
            seen = w->posted;
        }
        runFrame(*w);
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->done = seen;
        }
        w->cv.notify_all();
    }
}

}  // namespace

extern "C" {

int gudni_b200_multi_init(int n_devices, const int* devices, const gudni_spec* want, gudni_spec* got, gudni_multi** out) {
    if (!out || n_devices < 1 || n_devices > GUDNI_MULTI_MAX_DEVICES) return GUDNI_ERR_ARGUMENT;
    *out = nullptr;
    gudni_multi* mm = new (std::nothrow) gudni_multi();
    if (!mm) return GUDNI_ERR_OOM;
    for (int i = 0; i < n_devices; i++) mm->workers.emplace_back();
    for (int i = 0; i < n_devices; i++) {
        gudni_multi::Worker& w = mm->workers[i];
        w.owner = mm;
        w.index = i;
        w.device = devices ? devices[i] : i;
        const int rc = gudni_b200_init(w.device, want, &mm->spec, &w.ctx);
        if (rc != GUDNI_OK) {
            for (int k = 0; k < i; k++) gudni_b200_destroy(mm->workers[k].ctx);
            delete mm;
            return rc;
        }
    }
    // NVLink peer access towards every other device (the gather pushes strips to the presenting one)
    for (int i = 0; i < n_devices; i++)
        for (int k = 0; k < n_devices; k++) {
            const int a = mm->workers[i].device, b = mm->workers[k].device;
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (can) {
                cudaSetDevice(a);
                const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess) cudaGetLastError();   // already enabled: fine
            }
        }
    for (auto& w : mm->workers) w.thread = std::thread(workerLoop, &w);
    if (got) *got = mm->spec;
    *out = mm;
    return GUDNI_OK;
}

void gudni_b200_multi_destroy(gudni_multi* mm) {
    if (!mm) return;
    for (auto& w : mm->workers) {
        {
            std::lock_guard<std::mutex> lk(w.m);
            w.quit = true;
        }
        w.cv.notify_all();
        if (w.thread.joinable()) w.thread.join();
        cudaSetDevice(w.device);
        if (w.stripBuf) cudaFree(w.stripBuf);
        gudni_b200_destroy(w.ctx);
    }
    if (mm->canvas && mm->presenting >= 0) {
        cudaSetDevice(mm->workers[mm->presenting].device);
        cudaFree(mm->canvas);
    }
    delete mm;
}

const char* gudni_b200_multi_last_error(gudni_multi* mm) { return mm ? mm->err.c_str() : "null handle"; }

int gudni_b200_multi_set_presenting(gudni_multi* mm, int device_index) {
    if (!mm || device_index < -1 || device_index >= (int)mm->workers.size()) return GUDNI_ERR_ARGUMENT;
    if (mm->canvas && mm->presenting >= 0 && mm->presenting != device_index) {
        cudaSetDevice(mm->workers[mm->presenting].device);
        cudaFree(mm->canvas);
        mm->canvas = nullptr;
        mm->canvasCap = 0;
    }
    mm->presenting = device_index;
    return GUDNI_OK;
}

int gudni_b200_multi_canvas(gudni_multi* mm, void** dev_bgra, int* device) {
    if (!mm || !dev_bgra) return GUDNI_ERR_ARGUMENT;
    if (mm->presenting < 0 || !mm->canvas) { mm->err = "no presenting device, or no frame yet"; return GUDNI_ERR_STATE; }
    *dev_bgra = mm->canvas;
    if (device) *device = mm->workers[mm->presenting].device;
    return GUDNI_OK;
}

int gudni_b200_multi_frame(gudni_multi* mm, const void* geometry, size_t geometry_bytes, const float* substances, int n_substances,
                           const uint8_t* picture_bytes, size_t n_picture_bytes, const gudni_picture_use* picture_uses,
                           int n_picture_uses, const float background_rgba[4], int width, int height, int frame_number,
                           const gudni_shape_entry* entries, int n_entries, const gudni_generations* generations,
                           uint32_t* out_bgra, gudni_multi_stats* stats) {
    if (!mm || !background_rgba || width <= 0 || height <= 0 || n_entries < 0 || (n_entries && !entries)) return GUDNI_ERR_ARGUMENT;
    const int n = (int)mm->workers.size();
    const int tileRows = mm->spec.max_tile_size;
    FrameJob& j = mm->job;
    j = FrameJob{geometry, geometry_bytes, substances, n_substances, picture_bytes, n_picture_bytes, picture_uses, n_picture_uses,
                 {background_rgba[0], background_rgba[1], background_rgba[2], background_rgba[3]}, width, height, frame_number,
                 entries, n_entries, generations ? *generations : gudni_generations{}, out_bgra};
    // strips: from the shapes' boxes for a new scene, from the measured times of the last frame otherwise
    const bool sameScene = mm->lastWidth == width && mm->lastHeight == height && mm->lastEntries == n_entries && !mm->strips.empty();
    if (!sameScene) mm->strips = partitionRows(entries, n_entries, width, height, tileRows, n);
    else if (n > 1) mm->strips = rebalanceRows(mm->strips, mm->lastMs, height, tileRows);
    mm->lastWidth = width; mm->lastHeight = height; mm->lastEntries = n_entries;
    if (mm->presenting >= 0) {
        const size_t bytes = (size_t)width * height * 4;
        if (mm->canvasCap < bytes) {
            cudaSetDevice(mm->workers[mm->presenting].device);
            if (mm->canvas) cudaFree(mm->canvas);
            mm->canvas = nullptr; mm->canvasCap = 0;
            if (cudaMalloc(&mm->canvas, bytes) != cudaSuccess) { mm->err = "canvas allocation failed"; return GUDNI_ERR_OOM; }
            mm->canvasCap = bytes;
        }
    }
    mm->frameStart = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) {
        gudni_multi::Worker& w = mm->workers[i];
        {
            std::lock_guard<std::mutex> lk(w.m);
            w.strip = mm->strips[i];
            w.posted++;
        }
        w.cv.notify_all();
    }
    int rc = GUDNI_OK;
    for (auto& w : mm->workers) {
        std::unique_lock<std::mutex> lk(w.m);
        w.cv.wait(lk, [&] { return w.done == w.posted; });
        if (w.rc != GUDNI_OK && rc == GUDNI_OK) {
            rc = w.rc;
            mm->err = "device " + std::to_string(w.device) + ": " + w.err;
        }
    }
    const double frameMs = sinceStart(mm) * 1e3;
    mm->lastMs.assign(n, 0.0);
    gudni_multi_stats ms{};
    ms.n_devices = n;
    ms.ms_frame = (float)frameMs;
    double lastFinished = 0.0, lastLanded = 0.0;
    for (int i = 0; i < n; i++) {
        const gudni_multi::Worker& w = mm->workers[i];
        mm->lastMs[i] = (double)w.stats.ms_bin + w.stats.ms_raster + w.stats.ms_strands;
        ms.device[i] = w.device;
        ms.row_begin[i] = w.strip.rowBegin;
        ms.row_end[i] = w.strip.rowEnd;
        ms.ms_device[i] = (float)mm->lastMs[i];
        ms.total.n_tiles += w.stats.n_tiles;
        ms.total.n_shape_refs += w.stats.n_shape_refs;
        ms.total.n_thresholds += w.stats.n_thresholds;
        ms.total.n_spilled_threads += w.stats.n_spilled_threads;
        ms.total.n_overflow_threads += w.stats.n_overflow_threads;
        ms.total.algorithmic_bytes += w.stats.algorithmic_bytes;
        ms.total.ms_upload = std::max(ms.total.ms_upload, w.stats.ms_upload);
        ms.total.ms_bin = std::max(ms.total.ms_bin, w.stats.ms_bin);
        ms.total.ms_raster = std::max(ms.total.ms_raster, w.stats.ms_raster);
        ms.total.ms_download = std::max(ms.total.ms_download, w.stats.ms_download);
        lastFinished = std::max(lastFinished, w.finishedAt);
        lastLanded = std::max(lastLanded, w.landedAt);
    }
    ms.ms_gather_exposed = (float)((lastLanded - lastFinished) * 1e3);
    if (stats) *stats = ms;
    return rc;
}

// The two partition functions by themselves (host arithmetic only: callable, and tested, without a GPU).
// rows_out receives n_devices pairs (row_begin, row_end).
int gudni_b200_partition_rows(const gudni_shape_entry* entries, int n_entries, int width, int height, int tile_rows, int n_devices,
                              int* rows_out) {
    if (!rows_out || n_devices < 1 || width <= 0 || height <= 0 || tile_rows <= 0 || n_entries < 0 || (n_entries && !entries))
        return GUDNI_ERR_ARGUMENT;
    const std::vector<Strip> s = partitionRows(entries, n_entries, width, height, tile_rows, n_devices);
    for (int d = 0; d < n_devices; d++) { rows_out[2 * d] = s[d].rowBegin; rows_out[2 * d + 1] = s[d].rowEnd; }
    return GUDNI_OK;
}
int gudni_b200_rebalance_rows(const int* rows_in, const double* ms, int n_devices, int height, int tile_rows, int* rows_out) {
    if (!rows_in || !ms || !rows_out || n_devices < 1 || height <= 0 || tile_rows <= 0) return GUDNI_ERR_ARGUMENT;
    std::vector<Strip> last(n_devices);
    for (int d = 0; d < n_devices; d++) last[d] = Strip{rows_in[2 * d], rows_in[2 * d + 1]};
    const std::vector<Strip> s = rebalanceRows(last, std::vector<double>(ms, ms + n_devices), height, tile_rows);
    for (int d = 0; d < n_devices; d++) { rows_out[2 * d] = s[d].rowBegin; rows_out[2 * d + 1] = s[d].rowEnd; }
    return GUDNI_OK;
}

}  // extern "C"
