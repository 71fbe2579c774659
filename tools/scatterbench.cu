// Development micro-benchmark: cost of 4-byte scatters confined to windows of 2^w elements (L2 write-combining study).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
__device__ __forceinline__ u32 bij(u32 x, int bits) {  // bijection on [0, 2^bits)
    const u32 mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    x = (x * 0x9E3779B1u) & mask; x ^= x >> (bits / 2); x = (x * 0x85EBCA6Bu) & mask; x ^= x >> (bits / 2 + 1); x = (x * 0xC2B2AE35u) & mask;
    return x & mask;
}
__global__ void make_idx(u32* idx, u32* val, u64 n, int wbits) {
    u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    u32 hi = u32(t >> wbits) << wbits;
    idx[t] = hi | bij(u32(t) & ((1u << wbits) - 1u), wbits);
    val[t] = u32(t);
}
__global__ void scatter(const u32* __restrict__ idx, const u32* __restrict__ val, u64 m, u32* __restrict__ dst) {
    const u64 stride = u64(gridDim.x) * blockDim.x;
    for (u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x; t < m; t += stride) dst[idx[t]] = val[t];
}
__global__ void scatter_flat(const u32* __restrict__ idx, const u32* __restrict__ val, u64 m, u32* __restrict__ dst) {
    const u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < m) dst[idx[t]] = val[t];
}
int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 28;
    const u64 n = u64(1) << lg;
    u32 *idx, *val, *dst;
    cudaMalloc(&idx, 4 * n); cudaMalloc(&val, 4 * n); cudaMalloc(&dst, 4 * n);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int wbits = lg; wbits >= 12; wbits -= 2) {
        make_idx<<<unsigned((n + 255) / 256), 256>>>(idx, val, n, wbits);
        for (int variant = 0; variant < 3; variant++) {
            float best = 1e30f;
            for (int it = 0; it < 3; it++) {
                cudaEventRecord(a);
                if (variant == 0) scatter<<<148 * 16, 256>>>(idx, val, n, dst);
                else if (variant == 1) scatter<<<148 * 8, 256>>>(idx, val, n, dst);
                else scatter_flat<<<unsigned((n + 255) / 256), 256>>>(idx, val, n, dst);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            printf("n=2^%d window=2^%d elems (%.2f MiB) variant=%d: %.3f ms  %.2f Gelem/s\n", lg, wbits, 4.0 * (1u << wbits) / 1048576.0, variant, best, n / 1e6 / best);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
