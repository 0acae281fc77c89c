// cuda_runtime.h — HOST STAND-IN used only by tests/native/raster_emu.cpp (found first on its include
// path): enough of the CUDA C++ device language to compile this repo's kernels with g++ and run them under
// a cooperative SIMT emulator, one CTA at a time on one OS thread.  TEST HELPER, not product.
//
//   * every CUDA thread is a fiber (ucontext); a fiber runs until it reaches a warp collective
//     (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp) or __syncthreads, deposits its operand and
//     yields; when every live lane named in the mask (every live thread of the CTA) has arrived the
//     collective completes and the fibers pick up their results.  A collective that can never complete
//     (divergent lanes waiting on different collectives) is reported as a deadlock, not silently resolved;
//   * __shared__ is `static` (one CTA at a time), the dynamic shared array is a fixed host buffer;
//   * atomics are plain read-modify-writes (one OS thread), __ldg is a load;
//   * float arithmetic is the host's IEEE f32 (the emulator is built -ffp-contract=off like the oracle;
//     the library itself is built -fmad=false), __fmaf_rn is fmaf.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <cfloat>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <tuple>
#include <vector>

// after every system header the kernels' sources include (glibc spells attributes __noinline__ too)
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))



typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMalloc(void**, size_t) { return 1; }
inline cudaError_t cudaFree(void*) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }

// ---- vector types ---------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ---- the emulator -------------------------------------------------------------------------------------------
namespace cuemu {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    uint3 tid;
    int warp, lane;
    bool done;
    int wait;      // what the fiber is blocked on (kWait*): the scheduler only switches to fibers that can go on
};
enum { kWaitNone, kWaitWarpFree, kWaitWarpResult, kWaitBarrierFree, kWaitBarrier };
struct WarpSync {
    uint32_t arrived, release;      // lanes that deposited / lanes that still have to pick their result up
    uint64_t vals[32], results[32];
    int op, arg[32];
    uint32_t mask[32];
};
struct State {
    std::vector<Fiber> fibers;
    std::vector<WarpSync> warps;
    Fiber* cur;
    ucontext_t scheduler;
    uint3 bIdx, bDim, gDim;
    // __syncthreads
    unsigned barArrived, barRelease;
    unsigned long long switches;
    bool progress;
};
extern State S;
extern unsigned char dynamicShared[228 * 1024] __attribute__((aligned(128)));

inline void yield() { swapcontext(&S.cur->ctx, &S.scheduler); }
inline void yieldOn(int wait) { S.cur->wait = wait; yield(); S.cur->wait = kWaitNone; }
inline uint32_t liveMask(int warp) {
    uint32_t m = 0;
    for (int l = 0; l < 32; l++) {
        const size_t i = (size_t)warp * 32 + l;
        if (i < S.fibers.size() && !S.fibers[i].done) m |= 1u << l;
    }
    return m;
}
enum { kOpSync, kOpShfl, kOpShflUp, kOpShflDown, kOpShflXor, kOpBallot };

// Completes the warp's pending collective if every live lane named in the masks has arrived.
inline void tryComplete(int warp) {
    WarpSync& w = S.warps[warp];
    if (!w.arrived || w.release) return;
    const uint32_t live = liveMask(warp);
    uint32_t want = 0;
    for (int l = 0; l < 32; l++)
        if (w.arrived & (1u << l)) want |= w.mask[l];
    want &= live;
    if ((w.arrived & want) != want) return;
    if (w.arrived & ~want) { fprintf(stderr, "cuemu: lane outside the mask joined a collective\n"); abort(); }
    uint32_t ballot = 0;
    for (int l = 0; l < 32; l++)
        if ((w.arrived & (1u << l)) && w.vals[l]) ballot |= 1u << l;
    for (int l = 0; l < 32; l++) {
        if (!(w.arrived & (1u << l))) continue;
        int src = l;
        switch (w.op) {
            case kOpShfl: src = w.arg[l] & 31; break;
            case kOpShflUp: src = (l >= w.arg[l]) ? l - w.arg[l] : l; break;
            case kOpShflDown: src = (l + w.arg[l] < 32) ? l + w.arg[l] : l; break;
            case kOpShflXor: src = l ^ w.arg[l]; break;
            default: break;
        }
        if (!(w.arrived & (1u << src))) src = l;     // reading a lane that is not taking part: undefined in CUDA
        w.results[l] = (w.op == kOpBallot) ? (uint64_t)ballot : w.vals[src];
    }
    w.release = w.arrived;
    w.arrived = 0;
    S.progress = true;
}
inline uint64_t collective(int op, uint32_t mask, uint64_t value, int arg) {
    Fiber* f = S.cur;
    WarpSync& w = S.warps[f->warp];
    const uint32_t bit = 1u << f->lane;
    while (w.release) yieldOn(kWaitWarpFree);      // the previous collective is still handing out results
    if (w.arrived && w.op != op) {
        // lanes of one warp sitting in different collectives: legal only if the masks are disjoint; this
        // code base always converges first, so report it
        fprintf(stderr, "cuemu: warp %d mixes collectives %d and %d\n", f->warp, w.op, op);
        abort();
    }
    w.op = op;
    w.vals[f->lane] = value;
    w.arg[f->lane] = arg;
    w.mask[f->lane] = mask;
    w.arrived |= bit;
    S.progress = true;
    tryComplete(f->warp);
    while (!(w.release & bit)) yieldOn(kWaitWarpResult);
    const uint64_t r = w.results[f->lane];
    w.release &= ~bit;
    S.progress = true;
    return r;
}
inline void syncthreads() {
    // two-phase CTA barrier: barArrived counts deposits, barRelease the pick-ups still owed
    while (S.barRelease) yieldOn(kWaitBarrierFree);
    S.barArrived++;
    S.progress = true;
    for (;;) {
        if (S.barRelease) break;                       // completed by another thread
        unsigned live = 0;
        for (const Fiber& f : S.fibers) live += !f.done;
        if (S.barArrived == live) { S.barRelease = live; S.barArrived = 0; break; }
        yieldOn(kWaitBarrier);
    }
    S.barRelease--;
    S.progress = true;
}

// rcp.approx.ftz.f32: the hardware's reciprocal is within about 1 ulp of 1/d; `rcpUlpError` moves the model's
// answer that many ulps off the correctly rounded one, so tests can show that div3 (raster_device.cuh) does not
// depend on which approximation it starts from.
extern int rcpUlpError;
inline float rcpApprox(float d) {
    float r = (float)(1.0 / (double)d);
    for (int i = 0; i < (rcpUlpError < 0 ? -rcpUlpError : rcpUlpError); i++) r = nextafterf(r, rcpUlpError < 0 ? 0.0f : INFINITY);
    return r;
}

// can a blocked fiber make progress if it is switched to now?
inline bool canRun(const Fiber& f) {
    switch (f.wait) {
        case kWaitWarpFree: return S.warps[f.warp].release == 0;
        case kWaitWarpResult: return (S.warps[f.warp].release & (1u << f.lane)) != 0;
        case kWaitBarrierFree: return S.barRelease == 0;
        case kWaitBarrier: {
            if (S.barRelease) return true;
            unsigned live = 0;
            for (const Fiber& g : S.fibers) live += !g.done;
            return S.barArrived == live;        // a thread that finished may have completed the barrier
        }
        default: return true;
    }
}

template <class T> inline uint64_t pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "shuffle operand"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

void runBlock(void (*entry)(void*), void* args, dim3 grid, dim3 block, uint3 bidx);

template <class... A>
struct Call {
    void (*kernel)(A...);
    std::tuple<A...> args;
    static void entry(void* p) { Call* c = static_cast<Call*>(p); std::apply(c->kernel, c->args); }
};
// kernel<<<grid, block>>>(args...)
template <class... P, class... A>
inline void launch(void (*kernel)(P...), dim3 grid, dim3 block, A... args) {
    Call<P...> c{kernel, std::tuple<P...>(args...)};
    for (unsigned b = 0; b < grid.x; b++) runBlock(&Call<P...>::entry, &c, grid, block, uint3{b, 0, 0});
}

}  // namespace cuemu

#define threadIdx (cuemu::S.cur->tid)
#define blockIdx (cuemu::S.bIdx)
#define blockDim (cuemu::S.bDim)
#define gridDim (cuemu::S.gDim)

// ---- warp and CTA collectives ---------------------------------------------------------------------------------
template <class T> inline T __shfl_sync(unsigned m, T v, int src) { return cuemu::unpack<T>(cuemu::collective(cuemu::kOpShfl, m, cuemu::pack(v), src)); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, int d) { return cuemu::unpack<T>(cuemu::collective(cuemu::kOpShflUp, m, cuemu::pack(v), d)); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, int d) { return cuemu::unpack<T>(cuemu::collective(cuemu::kOpShflDown, m, cuemu::pack(v), d)); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int d) { return cuemu::unpack<T>(cuemu::collective(cuemu::kOpShflXor, m, cuemu::pack(v), d)); }
inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)cuemu::collective(cuemu::kOpBallot, m, pred ? 1 : 0, 0); }
inline int __any_sync(unsigned m, int pred) { return cuemu::collective(cuemu::kOpBallot, m, pred ? 1 : 0, 0) != 0; }
inline int __all_sync(unsigned m, int pred) { return cuemu::collective(cuemu::kOpBallot, m, pred ? 0 : 1, 0) == 0; }
inline void __syncwarp(unsigned m = 0xFFFFFFFFu) { cuemu::collective(cuemu::kOpSync, m, 0, 0); }
inline void __syncthreads() { cuemu::syncthreads(); }

// ---- scalar intrinsics ----------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float2int_rz(float f) { return f != f ? 0 : (f >= 2147483648.0f ? 2147483647 : (f <= -2147483648.0f ? (-2147483647 - 1) : (int)f)); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(long long v) { return __builtin_popcountll((unsigned long long)v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
template <class T> inline T atomicAdd(T* p, T v) { T old = *p; *p = (T)(old + v); return old; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o | v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { unsigned int o = *p; *p = o | v; return o; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
