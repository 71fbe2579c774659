// Micro-benchmark of the radix-sort digit pass (development tool, not part of the product library).
// Build one binary per configuration:  nvcc -DRS_THREADS_CFG=.. -DRS_IPT64_CFG=.. -DRS_MIN_CTAS_CFG=.. tools/sortbench.cu
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tudocomp_b200/csrc/radix_sort.cuh"

namespace tdc {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
LaunchScope::LaunchScope(const char*, cudaStream_t s, bool) : slot(-1), st(s) {}
LaunchScope::~LaunchScope() {}
void prof_add_bytes(const char*, double) {}
}  // namespace tdc
using namespace tdc;

__global__ void fill_keys(u64* k, u64 m, u64 seed, int bits) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u64 x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    k[i] = bits >= 64 ? x : (x & ((u64(1) << bits) - 1));
}
__global__ void check_sorted(const u64* k, const u32* v, u64 m, u32* bad) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i + 1 >= m) return;
    if (k[i] > k[i + 1] || (k[i] == k[i + 1] && v[i] > v[i + 1])) atomicAdd(bad, 1u);
}

__global__ void check_sorted_keys(const u64* k, u64 m, int lo, int bits, u32* bad) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i + 1 >= m) return;
    const u64 mask = bits >= 64 - lo ? ~u64(0) : ((u64(1) << bits) - 1);
    if (((k[i] >> lo) & mask) > ((k[i + 1] >> lo) & mask)) atomicAdd(bad, 1u);
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 28;
    const int bits = argc > 2 ? atoi(argv[2]) : 64;
    const bool keysonly = argc > 3 && argv[3][0] == 'k';  // packed records: sort bits [31, 31 + bits) of u64 records, 16 B/elem
    const u64 m = u64(1) << lg;
    SortWorkspace ws;
    ws.sm_count = 148;
    ws.max_tiles = std::max(rs_tiles<u64>(m), rs_tiles_keys(m)) + 1;
    cudaMalloc(&ws.hist, sizeof(u32) * RS_MAX_PASSES * RS_RADIX);
    cudaMalloc(&ws.uniform, sizeof(u32) * RS_MAX_PASSES);
    cudaMalloc(&ws.desc, sizeof(ull) * ws.max_tiles * RS_RADIX);
    cudaMemset(ws.desc, 0, sizeof(ull) * ws.max_tiles * RS_RADIX);
    cudaMallocHost(&ws.h_uniform, sizeof(u32) * RS_MAX_PASSES);
    auto k1 = rs_onesweep_kernel<u64, true>;
    auto k2 = rs_onesweep_kernel<u64, false>;
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<u64>()));
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<u64>()));
    u64* k[2]; u32* v[2]; u32* bad;
    cudaMalloc(&k[0], 8 * m); cudaMalloc(&k[1], 8 * m); cudaMalloc(&v[0], 4 * m); cudaMalloc(&v[1], 4 * m); cudaMalloc(&bad, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    int res = 0;
    for (int it = 0; it < 5; it++) {
        fill_keys<<<unsigned(div_up(m, 256)), 256>>>(k[0], m, 1234 + it, keysonly ? 64 : bits);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        if (keysonly) { if (radix_sort_keys(ws, 0, k, m, 31, 31 + bits, &res) < 0) return 1; }
        else if (radix_sort_pairs<u64>(ws, 0, k, v, m, 0, bits, true, &res) < 0) return 1;
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = std::min(best, ms);
    }
    cudaMemset(bad, 0, 4);
    if (keysonly) check_sorted_keys<<<unsigned(div_up(m, 256)), 256>>>(k[res], m, 31, bits, bad);
    else check_sorted<<<unsigned(div_up(m, 256)), 256>>>(k[res], v[res], m, bad);
    u32 hbad = 0; cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    const int npass = (bits + 7) / 8;
    const double bpe = keysonly ? 16.0 : 24.0;
    printf("%s threads=%d ipt=%d minctas=%d smem=%zu m=2^%d bits=%d: %.3f ms total, %.3f ms/pass incl. histogram, %.1f GB/s per pass (%.0f B/elem), unsorted=%u %s\n",
           keysonly ? "KEYSONLY" : "pairs", RS_THREADS, keysonly ? RsOcc<u64, true>::IPT : RsCfg<u64>::IPT, keysonly ? RsOcc<u64, true>::MIN_CTAS : RsCfg<u64>::MIN_CTAS,
           rs_smem_bytes<u64>(keysonly), lg, bits, best, best / npass, bpe * m / 1e9 / (best / npass / 1e3), bpe, hbad,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    return hbad != 0;
}
