// Shared device/host helpers for the tdcgpu kernels (sm_100a).  No torch, no CUB/Thrust: every kernel is hand-written.
#pragma once
#include <cstdint>
#include <cstdio>

#ifndef TDC_CUSIM
#include <cuda_runtime.h>
// every kernel launch goes through this macro: it counts launches and, when profiling is on, brackets the launch with
// CUDA events on the launching stream (tdcgpu_profile_* in include/tdcgpu.h)
#define TDC_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        tdc::LaunchScope ls__(#kernel, (stream));                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
    } while (0)
#define TDC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace tdc {

typedef uint32_t u32;
typedef uint64_t u64;
typedef unsigned long long ull;

static const u32 kWarp = 32;
static const u32 kFull = 0xffffffffu;

// ----------------------------------------------------------------------------------------------------------------
// launch accounting / optional per-kernel event timing (tdcgpu_api.cu)
// ----------------------------------------------------------------------------------------------------------------
#ifndef TDC_CUSIM
struct LaunchScope {
    int slot;
    cudaStream_t st;
    LaunchScope(const char* name, cudaStream_t stream, bool is_kernel = true);  // false: timed (e.g. an NCCL exchange) but not counted as a launch
    ~LaunchScope();
};
#endif
void prof_add_bytes(const char* name, double bytes);  // algorithmic bytes of the launch just issued under `name`

// ----------------------------------------------------------------------------------------------------------------
// error plumbing: kernels never abort; host wrappers return negative codes and stash a message (tdcgpu_last_error)
// ----------------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define TDC_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            tdc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)
#define TDC_TRY(expr)           \
    do {                        \
        int r__ = (expr);       \
        if (r__ < 0) return r__; \
    } while (0)
#define TDC_KCHECK() TDC_CUDA(cudaGetLastError())

static inline u32 bits_for_host(u64 v) {  // tudocomp util.hpp:194 — bits_for(0) == 1
    u32 b = 1;
    while (v >>= 1) b++;
    return b;
}
static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------------------------
// warp / block primitives
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ u32 lanemask_lt() { return (1u << lane_id()) - 1u; }

template <class T>
__device__ __forceinline__ T warp_inclusive_sum(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(kFull, v, d);
        if ((int)lane_id() >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ u32 warp_inclusive_max(u32 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(kFull, v, d);
        if ((int)lane_id() >= d) v = max(v, o);
    }
    return v;
}
__device__ __forceinline__ u32 warp_min(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(kFull, v, d));
    return v;
}
__device__ __forceinline__ u32 warp_max(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(kFull, v, d));
    return v;
}
template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

// Block-wide exclusive sum over one value per thread.  `scratch` must hold 33 T's.  All threads must call.
// Returns the exclusive prefix for this thread; *total receives the block total (valid in every thread).
template <class T>
__device__ __forceinline__ T block_exclusive_sum(T v, T* scratch, T* total) {
    const u32 lane = lane_id(), w = warp_id(), nw = (blockDim.x + 31) >> 5;
    T inc = warp_inclusive_sum(v);
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = lane < nw ? scratch[lane] : T(0);
        T xi = warp_inclusive_sum(x);
        scratch[lane] = xi - x;  // exclusive warp offsets
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    T res = scratch[w] + inc - v;
    *total = scratch[32];
    __syncthreads();  // scratch reusable after return
    return res;
}

// Block-wide inclusive max-scan over one u32 per thread (0 is the identity).  scratch: 33 u32.
__device__ __forceinline__ u32 block_inclusive_max(u32 v, u32* scratch, u32* total) {
    const u32 lane = lane_id(), w = warp_id(), nw = (blockDim.x + 31) >> 5;
    u32 inc = warp_inclusive_max(v);
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 x = lane < nw ? scratch[lane] : 0u;
        u32 xi = warp_inclusive_max(x);
        u32 ex = __shfl_up_sync(kFull, xi, 1);
        if (lane == 0) ex = 0;
        scratch[lane] = ex;
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    u32 res = max(scratch[w], inc);
    *total = scratch[32];
    __syncthreads();
    return res;
}

// ----------------------------------------------------------------------------------------------------------------
// memory helpers
// ----------------------------------------------------------------------------------------------------------------
// Read 8 text bytes starting at an arbitrary (unaligned) byte offset, little-endian (lowest address = lowest byte).
// The text buffer is padded with >= 16 readable bytes past n, so the second aligned word is always in bounds.
__device__ __forceinline__ u64 load_text8(const uint8_t* __restrict__ t, u64 off) {
    const u64* w = reinterpret_cast<const u64*>(t + (off & ~u64(7)));
    const u32 sh = u32(off & 7u) * 8u;
    u64 lo = __ldg(w);
    if (sh == 0) return lo;
    u64 hi = __ldg(w + 1);
    return (lo >> sh) | (hi << (64u - sh));
}

}  // namespace tdc
