// hostcopy.cuh — copies between PAGEABLE caller memory and the device, for the caller that changes nothing
// (the Haskell front end hands the library plain `Pile`s and a malloc'ed bitmap, as it hands them to
// clEnqueueWriteBuffer / clEnqueueReadBuffer: OpenCL/Instances.hs:39-75).
//
// cudaMemcpyAsync on pageable memory goes through the driver's own staging on the calling thread; what it reaches on the
// GPU box is ≈ 18 GB/s in either direction, a third of the link.  Here the context owns a small ring of page-locked
// staging buffers and a few worker threads: a transfer is cut into chunks, a chunk is copied between the caller's
// memory and a staging buffer by all threads at once (a single thread's memcpy is what limits the driver's path), and
// crosses PCIe by DMA while the next chunk is being copied.  Page-locked caller memory (gudni_b200_host_register) does
// not come through here: it is handed to the copy engine as it is.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

struct HostCopier {
    static constexpr size_t kChunk = (size_t)4 << 20;   // bytes per staging buffer
    static constexpr int kRing = 4;                      // staging buffers in flight
    static constexpr size_t kWorthIt = (size_t)1 << 20;  // smaller transfers take the driver's path

    void* stage[kRing] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kRing] = {nullptr, nullptr, nullptr, nullptr};
    bool inFlight[kRing] = {false, false, false, false};   // a DMA queued earlier may still be reading the buffer
    int next = 0;
    int threads = 0;          // workers + the calling thread; 0: not started, < 0: disabled
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable wake, done;
    // the job the workers share: [src, src + bytes) -> dst in `parts` slices
    const char* src = nullptr;
    char* dst = nullptr;
    size_t bytes = 0;
    int parts = 0, nextPart = 0, pendingParts = 0;
    unsigned long long generation = 0;
    bool stop = false;

    ~HostCopier() { shutdown(); }

    // Is `p` ordinary pageable host memory?  (Registered or cudaMallocHost'ed memory, and anything the runtime does
    // not want to tell us about, is not.)
    static bool pageable(const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeUnregistered;
    }

    bool start(int wanted) {
        if (threads != 0) return threads > 0;
        threads = -1;
        if (wanted < 1) return false;
        for (int i = 0; i < kRing; i++) {
            if (cudaMallocHost(&stage[i], kChunk) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                release();
                return false;
            }
        }
        try {
            for (int i = 1; i < wanted; i++) workers.emplace_back([this] { work(); });
        } catch (...) {   // no more threads to be had: what started is enough
        }
        threads = 1 + (int)workers.size();
        return true;
    }

    void release() {
        for (int i = 0; i < kRing; i++) {
            if (stage[i]) cudaFreeHost(stage[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
            stage[i] = nullptr;
            ev[i] = nullptr;
        }
    }

    void shutdown() {
        {
            std::lock_guard<std::mutex> lock(m);
            stop = true;
        }
        wake.notify_all();
        for (std::thread& t : workers) t.join();
        workers.clear();
        release();
        threads = -1;
    }

    // takes slices of the current job until none is left; returns with the lock held
    void takeParts(std::unique_lock<std::mutex>& lock) {
        while (nextPart < parts) {
            const int part = nextPart++;
            const size_t per = ((bytes / (size_t)parts) + 4095) & ~(size_t)4095;
            const size_t from = std::min(bytes, per * (size_t)part), to = part == parts - 1 ? bytes : std::min(bytes, per * (size_t)(part + 1));
            lock.unlock();
            if (to > from) std::memcpy(dst + from, src + from, to - from);
            lock.lock();
            if (--pendingParts == 0) done.notify_all();
        }
    }

    void work() {
        std::unique_lock<std::mutex> lock(m);
        unsigned long long seen = 0;
        for (;;) {
            wake.wait(lock, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            takeParts(lock);
        }
    }

    // memcpy by all threads; returns when the last byte is in place
    void copy(void* to, const void* from, size_t n) {
        if (threads <= 1 || n < ((size_t)256 << 10)) {
            std::memcpy(to, from, n);
            return;
        }
        std::unique_lock<std::mutex> lock(m);
        src = static_cast<const char*>(from);
        dst = static_cast<char*>(to);
        bytes = n;
        parts = threads;
        nextPart = 0;
        pendingParts = parts;
        generation++;
        wake.notify_all();
        takeParts(lock);
        done.wait(lock, [&] { return pendingParts == 0; });
    }

    // pageable host -> device, in order on `stream`; the caller's memory is free again when this returns
    cudaError_t upload(void* dev, const void* host, size_t n, cudaStream_t stream) {
        for (size_t off = 0; off < n; off += kChunk) {
            const int i = next;
            next = (next + 1) % kRing;
            const size_t len = std::min(kChunk, n - off);
            cudaError_t e = inFlight[i] ? cudaEventSynchronize(ev[i]) : cudaSuccess;   // the DMA that last read this buffer
            if (e != cudaSuccess) return e;
            copy(stage[i], static_cast<const char*>(host) + off, len);
            if ((e = cudaMemcpyAsync(static_cast<char*>(dev) + off, stage[i], len, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(ev[i], stream)) != cudaSuccess) return e;
            inFlight[i] = true;
        }
        // (the staging buffers, not the caller's memory, are what the copy engine still reads)
        return cudaSuccess;
    }

    // device -> pageable host; everything queued on `stream` before is waited for; the bytes are in place on return
    cudaError_t download(void* host, const void* dev, size_t n, cudaStream_t stream) {
        const int chunks = (int)((n + kChunk - 1) / kChunk);
        auto issue = [&](int k) -> cudaError_t {
            const size_t off = (size_t)k * kChunk, len = std::min(kChunk, n - off);
            cudaError_t e = cudaMemcpyAsync(stage[k % kRing], static_cast<const char*>(dev) + off, len, cudaMemcpyDeviceToHost, stream);
            return e != cudaSuccess ? e : cudaEventRecord(ev[k % kRing], stream);
        };
        cudaError_t e;
        for (int i = 0; i < kRing; i++) {      // uploads queued earlier may still be reading the buffers
            if (inFlight[i] && (e = cudaEventSynchronize(ev[i])) != cudaSuccess) return e;
            inFlight[i] = false;
        }
        for (int k = 0; k < std::min(chunks, kRing); k++)
            if ((e = issue(k)) != cudaSuccess) return e;
        for (int k = 0; k < chunks; k++) {
            const size_t off = (size_t)k * kChunk, len = std::min(kChunk, n - off);
            if ((e = cudaEventSynchronize(ev[k % kRing])) != cudaSuccess) return e;
            copy(static_cast<char*>(host) + off, stage[k % kRing], len);
            if (k + kRing < chunks && (e = issue(k + kRing)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
};
