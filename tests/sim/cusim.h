// TEST INFRASTRUCTURE ONLY — "cusim": a tiny CPU interpreter for the subset of CUDA the kernels in
// tudocomp_b200/csrc use, so that kernel LOGIC can be debugged in this GPU-less container before GPU minutes are spent.
//
// It is NOT a CPU fallback: the product (libtdcgpu.so) is compiled by nvcc only and never includes this header; the
// Python package never loads the simulator binary.  tests/test_sim_*.py build the kernels a second time with
// `g++ -DTDC_CUSIM -include tests/sim/cusim.h` into tests/sim/_build/libtdcsim.so and compare against the oracle.
//
// Model: one OS thread; a kernel launch runs CTAs one after another in blockIdx order; inside a CTA every CUDA thread is
// a ucontext fiber scheduled round-robin; __syncthreads / warp collectives are rendezvous points.  Because CTAs run in
// order and to completion, decoupled look-back never spins here (memory-ordering bugs are NOT caught; racecheck on the
// GPU box covers that).
#pragma once
#ifndef TDC_CUSIM
#error "cusim.h is only for -DTDC_CUSIM builds"
#endif

#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }

namespace cusim {

struct Warp {
    uint64_t slot[32];
    uint32_t arrived = 0, departed = 0;
};

struct State {
    uint3 tid{0, 0, 0}, bid{0, 0, 0};
    dim3 bdim, gdim;
    // scheduler
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<uint8_t> done, at_barrier;
    unsigned cur = 0, nthreads = 0, barrier_count = 0, live = 0;
    uint64_t events = 0;  // fibers finished + barriers released + warp collectives completed (stall detection)
    std::vector<Warp> warps;
    const std::function<void()>* body = nullptr;
    unsigned char* dyn_smem = nullptr;
    size_t dyn_cap = 0;
};
inline State& S() {
    static State s;
    return s;
}
static const size_t kStack = 256 * 1024;

inline void yield() {
    State& s = S();
    swapcontext(&s.ctx[s.cur], &s.sched);
}

inline void fiber_entry() {
    State& s = S();
    (*s.body)();
    s.done[s.cur] = 1;
    s.live--;
    s.events++;
    swapcontext(&s.ctx[s.cur], &s.sched);
}

inline void run_block() {
    State& s = S();
    unsigned nt = s.bdim.x * s.bdim.y * s.bdim.z;
    s.nthreads = nt;
    if (s.ctx.size() < nt) {
        size_t old = s.ctx.size();
        s.ctx.resize(nt);
        s.stacks.resize(nt, nullptr);
        for (size_t i = old; i < nt; i++) s.stacks[i] = (char*)malloc(kStack);
    }
    s.done.assign(nt, 0);
    s.at_barrier.assign(nt, 0);
    s.warps.assign((nt + 31) / 32, Warp());
    s.barrier_count = 0;
    s.live = nt;
    for (unsigned t = 0; t < nt; t++) {
        getcontext(&s.ctx[t]);
        s.ctx[t].uc_stack.ss_sp = s.stacks[t];
        s.ctx[t].uc_stack.ss_size = kStack;
        s.ctx[t].uc_link = &s.sched;
        makecontext(&s.ctx[t], (void (*)())fiber_entry, 0);
    }
    uint64_t last_events = s.events;
    unsigned idle_sweeps = 0;
    while (s.live > 0) {
        bool progressed = false;
        for (unsigned t = 0; t < nt; t++) {
            if (s.done[t] || s.at_barrier[t]) continue;
            s.cur = t;
            s.tid.x = t % s.bdim.x;
            s.tid.y = (t / s.bdim.x) % s.bdim.y;
            s.tid.z = t / (s.bdim.x * s.bdim.y);
            swapcontext(&s.sched, &s.ctx[t]);
            progressed = true;
        }
        if (s.barrier_count > 0 && s.barrier_count == s.live) {
            // everyone alive is at the CTA barrier: release
            std::fill(s.at_barrier.begin(), s.at_barrier.end(), 0);
            s.barrier_count = 0;
            progressed = true;
            s.events++;
        }
        if (s.events == last_events) {
            if (++idle_sweeps > 100000) progressed = false;  // fibers only spin: divergent collective / lost barrier
        } else {
            last_events = s.events;
            idle_sweeps = 0;
        }
        if (!progressed) {
            fprintf(stderr, "cusim: deadlock in block %u (barrier_count=%u live=%u)\n", s.bid.x, s.barrier_count, s.live);
            abort();
        }
    }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    State& s = S();
    s.gdim = grid;
    s.bdim = block;
    s.body = &body;
    if (smem > s.dyn_cap) {
        free(s.dyn_smem);
        s.dyn_smem = (unsigned char*)aligned_alloc(128, (smem + 127) / 128 * 128);
        s.dyn_cap = smem;
    }
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                s.bid = uint3{bx, by, bz};
                run_block();
            }
}

// ---- warp rendezvous: every lane named in `mask` deposits a value, all see all ----
inline void warp_exchange(uint32_t mask, uint64_t v, uint64_t out[32]) {
    State& s = S();
    unsigned lane = s.cur & 31;
    Warp& w = s.warps[s.cur >> 5];
    // lanes of the last (partial) warp that do not exist are dropped from the mask
    unsigned base = (s.cur >> 5) * 32;
    if (base + 32 > s.nthreads) mask &= (1u << (s.nthreads - base)) - 1u;
    uint32_t bit = 1u << lane;
    while (w.departed & bit) yield();  // previous collective still draining
    w.slot[lane] = v;
    w.arrived |= bit;
    while ((w.arrived & mask) != mask) yield();
    for (int i = 0; i < 32; i++) out[i] = w.slot[i];
    w.departed |= bit;
    s.events++;
    if ((w.departed & mask) == mask) {
        w.arrived &= ~mask;
        w.departed &= ~mask;
    } else {
        while (w.departed & bit) yield();
    }
}

}  // namespace cusim

#define threadIdx (cusim::S().tid)
#define blockIdx (cusim::S().bid)
#define blockDim (cusim::S().bdim)
#define gridDim (cusim::S().gdim)
static const int warpSize = 32;

inline void __syncthreads() {
    cusim::State& s = cusim::S();
    s.at_barrier[s.cur] = 1;
    s.barrier_count++;
    cusim::yield();
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    uint64_t o[32];
    cusim::warp_exchange(mask, 0, o);
}
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) { }

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl");
    uint64_t in = 0, o[32];
    memcpy(&in, &v, sizeof(T));
    cusim::warp_exchange(mask, in, o);
    int lane = cusim::S().cur & 31;
    int base = lane & ~(width - 1);
    int s = base + (src & (width - 1));
    T r;
    memcpy(&r, &o[s], sizeof(T));
    return r;
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t in = 0, o[32];
    memcpy(&in, &v, sizeof(T));
    cusim::warp_exchange(mask, in, o);
    int lane = cusim::S().cur & 31;
    int base = lane & ~(width - 1);
    int s = lane - (int)delta;
    if (s < base) s = lane;
    T r;
    memcpy(&r, &o[s], sizeof(T));
    return r;
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t in = 0, o[32];
    memcpy(&in, &v, sizeof(T));
    cusim::warp_exchange(mask, in, o);
    int lane = cusim::S().cur & 31;
    int base = lane & ~(width - 1);
    int s = lane + (int)delta;
    if (s >= base + width) s = lane;
    T r;
    memcpy(&r, &o[s], sizeof(T));
    return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    uint64_t in = 0, o[32];
    memcpy(&in, &v, sizeof(T));
    cusim::warp_exchange(mask, in, o);
    int lane = cusim::S().cur & 31;
    int s = lane ^ x;
    (void)width;
    T r;
    memcpy(&r, &o[s], sizeof(T));
    return r;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint64_t o[32];
    cusim::warp_exchange(mask, pred ? 1 : 0, o);
    unsigned r = 0;
    cusim::State& s = cusim::S();
    unsigned base = (s.cur >> 5) * 32;
    for (int i = 0; i < 32; i++)
        if ((mask >> i & 1) && base + i < s.nthreads && o[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
    cusim::State& s = cusim::S();
    unsigned base = (s.cur >> 5) * 32;
    unsigned m = mask;
    if (base + 32 > s.nthreads) m &= (1u << (s.nthreads - base)) - 1u;
    return __ballot_sync(mask, pred) == m;
}
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    uint64_t in = 0, o[32];
    memcpy(&in, &v, sizeof(T));
    cusim::warp_exchange(mask, in, o);
    cusim::State& s = cusim::S();
    unsigned base = (s.cur >> 5) * 32;
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if ((mask >> i & 1) && base + i < s.nthreads && o[i] == in) r |= 1u << i;
    return r;
}
inline unsigned __activemask() { return 0xffffffffu; }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    uint64_t o[32];
    cusim::warp_exchange(mask, v, o);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) r += (unsigned)o[i];
    return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    uint64_t o[32];
    cusim::warp_exchange(mask, v, o);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) r = std::min(r, (unsigned)o[i]);
    return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    uint64_t o[32];
    cusim::warp_exchange(mask, v, o);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) r = std::max(r, (unsigned)o[i]);
    return r;
}

// ---- scalar intrinsics ----
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
using std::max;
using std::min;

// ---- atomics (single OS thread) ----
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- runtime API subset ----
typedef int cudaError_t;
typedef int cudaStream_t;
static const cudaError_t cudaSuccess = 0;
struct cudaEvent_st { std::chrono::steady_clock::time_point t; };
typedef cudaEvent_st* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; char name[64]; int major, minor; };
inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p)); p->multiProcessorCount = 4; p->totalGlobalMem = size_t(8) << 30; strcpy(p->name, "cusim");
    p->major = 10; p->minor = 0; return 0;
}
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = size_t(8) << 30; *t = size_t(8) << 30; return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return 0; }
inline cudaError_t cudaHostUnregister(void*) { return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 1; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 1; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cudaEvent_st(); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0;
}
static const unsigned cudaStreamNonBlocking = 1;
static const unsigned cudaEventDisableTiming = 2;
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new cudaEvent_st(); return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static const unsigned cudaHostRegisterDefault = 0;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

// Launch macro shared with the CUDA build (see tdc_common.cuh): TDC_LAUNCH(kernel, grid, block, smem, stream, args...)
#define TDC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cusim::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
#define TDC_DYN_SMEM(name) unsigned char* name = cusim::S().dyn_smem
namespace tdc { inline void prof_add_bytes(const char*, double) {} }
