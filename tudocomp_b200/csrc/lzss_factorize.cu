// lzss_lcp factorisation on the GPU (sm_100a): emits exactly the factor list of the loop in
// /root/reference/include/tudocomp/compressors/LZSSLCPCompressor.hpp:60-115.
//
// What the reference does per visited text position i (rank p = ISA[i]):
//   PSV side (:71-77): walk up from p while SA[q] > i, l_up  = min LCP[q+1..p]  (q = nearest rank above with SA[q] < i)
//   NSV side (:82-96): walk down to the first SA[q] < i,   l_dn = min LCP[p+1..q]  (0 if there is none)
//   len = max(l_up, l_dn); if len >= threshold emit (i, SA[l_up == len ? psv : nsv], len) and i += len, else i += 1.
// The naive walks are unbounded.  Here:
//   1. min-trees (fan-out 32) over SA and LCP;
//   2. every rank p finds PSV/NSV and the range minima in one up/down tree walk (abandoned as soon as the running
//      minimum drops below the threshold, because such a side can never produce a factor) and scatters
//      (len << 1 | side) to text order;
//   3. the greedy chain i -> i + max(1, len) from 0: per tile of text positions a shared-memory pointer-jumping
//      computes where each position leaves the tile; speculative per-region walkers plus a short scalar stitch over
//      tile exits find each tile's entry; then one thread per tile walks its part of the chain and marks it;
//   4. visited positions with len > 0 are compacted, in position order, to (pos, src, len) records
//      (lzss::Factor, compressors/lzss/LZSSFactors.hpp:13-20); src is recovered by repeating the winning side's walk.
//      PSV wins ties (:101).
#include "lzss_kernels.cuh"

namespace tdc {

int factorize_lzss_lcp(Ctx& c, u32 threshold) {
    if (threshold < 1) { set_error("lzss_lcp: threshold must be >= 1"); return -5; }
    if ((c.have & (DS_SA | DS_ISA | DS_LCP)) != (DS_SA | DS_ISA | DS_LCP)) {
        set_error("lzss_lcp: SA, ISA and LCP must be built first");
        return -6;
    }
    const u32 n = u32(c.n);
    // (len << 1 | side) words: only a text beyond 2^31 bytes can have a common prefix that long (include/tdcgpu.h, tdcgpu_set_text)
    if (c.max_lcp >= (1u << 31)) { set_error("lzss_lcp: common prefixes of 2^31 bytes or more are not supported (max_lcp = %u)", c.max_lcp); return -5; }
    cudaStream_t st = c.stream;
    c.num_factors = 0;
    c.flen_min = 0xffffffffu;
    c.flen_max = 0;
    c.have_factors = true;
    c.enc.prepared = c.enc.encoded = false;
    if (n <= 1) return 0;

    c.arena.reset();
    // ---- 1. min-trees ----
    MinTree T;
    T.a[0] = c.d_sa;
    T.l[0] = c.d_lcp;
    T.sz[0] = n;
    T.nlev = 1;
    while (T.sz[T.nlev - 1] > 32) {
        if (T.nlev >= MT_MAX_LEVELS) { set_error("min-tree too deep"); return -7; }
        const u32 szi = T.sz[T.nlev - 1], szo = u32(div_up(szi, 32));
        u32* a = c.arena.take<u32>(szo);
        u32* l = c.arena.take<u32>(szo);
        if (!a || !l) { set_error("lzss_lcp: scratch arena too small"); return -2; }
        TDC_LAUNCH(mintree_level_kernel, u32(div_up(div_up(u64(szo), MT_OUT_PER_WARP) * 32, 256)), 256, 0, st, T.a[T.nlev - 1], T.l[T.nlev - 1], szi, a, l, szo);
        T.a[T.nlev] = a;
        T.l[T.nlev] = l;
        T.sz[T.nlev] = szo;
        T.nlev++;
    }
    for (int i = T.nlev; i < MT_MAX_LEVELS; i++) { T.a[i] = nullptr; T.l[i] = nullptr; T.sz[i] = 0; }
    TDC_KCHECK();

    const u32 ntiles = u32(div_up(u64(n), CH_TILE));
    u32* lenside = c.arena.take<u32>(n);
    u32* exitp = c.arena.take<u32>(n);
    u32* entry = c.arena.take<u32>(ntiles);
    u32* fmask = c.arena.take<u32>(u64(ntiles) * (CH_TILE / 32));
    u32* tile_cnt = c.arena.take<u32>(ntiles);
    if (!lenside || !exitp || !entry || !fmask || !tile_cnt) { set_error("lzss_lcp: scratch arena too small"); return -2; }

    // ---- 2. LPF per rank ----
    {
        // (index, value) pairs of the scatter to text order: the index side is SA itself (read only)
        u32* sc_idx[2] = {c.d_sa, c.arena.take<u32>(n)};
        u32* sc_val[2] = {c.arena.take<u32>(n), c.arena.take<u32>(n)};
        if (!sc_idx[1] || !sc_val[0] || !sc_val[1]) { set_error("lzss_lcp: scratch arena too small"); return -2; }
        auto lpf_tile_kernel = tdc::lpf_tile_kernel<false>;
        TDC_CUDA(cudaFuncSetAttribute(lpf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lpf_smem_bytes())));
        TDC_LAUNCH(lpf_tile_kernel, u32(div_up(u64(n), LPF_TILE)), LPF_THREADS, lpf_smem_bytes(), st, T, n, threshold, sc_val[0], LpfDistOut());
        prof_add_bytes("lpf_tile_kernel", double(n) * 12);
        TDC_KCHECK();
        TDC_TRY(partitioned_scatter(c.sortws, st, sc_idx, sc_val, n, lenside, n, true));
    }
    // ---- 3. chain ----
    TDC_LAUNCH(chain_exit_kernel, ntiles, CH_THREADS, 0, st, lenside, n, exitp);
    TDC_LAUNCH(fill_u32_kernel, u32(div_up(u64(ntiles), 256)), 256, 0, st, entry, u64(ntiles), CH_NONE);
    {
        // regions ~ sqrt(ntiles / 2): the walkers' hops (tiles per region) balance the stitcher's hops (~2 per region)
        u32 regions = 1;
        while (u64(regions) * regions * 2 < ntiles) regions *= 2;
        const u32 tiles_per_region = u32(div_up(u64(ntiles), regions));
        regions = u32(div_up(u64(ntiles), tiles_per_region));
        u32* region_exit = c.arena.take<u32>(regions);
        if (!region_exit) { set_error("lzss_lcp: scratch arena too small"); return -2; }
        TDC_LAUNCH(chain_entries_spec_kernel, u32(div_up(u64(regions), 128)), 128, 0, st, exitp, n, tiles_per_region, regions, entry, region_exit);
        TDC_LAUNCH(chain_entries_stitch_kernel, 1, 32, 0, st, exitp, n, tiles_per_region, entry, region_exit, 0u, (u32*)nullptr);
    }
    TDC_CUDA(cudaMemsetAsync(fmask, 0, sizeof(u32) * u64(ntiles) * (CH_TILE / 32), st));
    TDC_LAUNCH(chain_mark_kernel, u32(div_up(u64(ntiles), 128)), 128, 0, st, lenside, n, ntiles, entry, fmask, tile_cnt);
    u32* d_total = c.d_scalars + 0;
    u32* d_minmax = c.d_scalars + 2;
    TDC_LAUNCH(scan_counts_kernel, 1, 1024, 0, st, tile_cnt, ntiles, d_total);
    TDC_KCHECK();
    c.h_scalars[2] = 0xffffffffu;
    c.h_scalars[3] = 0;
    TDC_CUDA(cudaMemcpyAsync(d_minmax, c.h_scalars + 2, 2 * sizeof(u32), cudaMemcpyHostToDevice, st));
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, d_total, sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    const u64 z = c.h_scalars[0];
    if (z > c.factors_cap) {
        if (c.d_factors) TDC_CUDA(cudaFree(c.d_factors));
        c.d_factors = nullptr;
        c.factors_cap = 0;
        const u64 cap = z + z / 8 + 1024;
        TDC_CUDA(cudaMalloc(&c.d_factors, cap * sizeof(Factor)));
        c.factors_cap = cap;
    }
    // ---- 4. emit ----
    if (z > 0) {
        auto emit_factors = emit_factors_kernel<false>;
        TDC_LAUNCH(emit_factors, ntiles, CH_TILE / 32, 0, st, T, c.d_isa, lenside, fmask, tile_cnt, threshold, c.d_factors, d_minmax,
                   (const u32*)nullptr, 0u);
        TDC_KCHECK();
    }
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 2, d_minmax, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    c.num_factors = z;
    c.flen_min = c.h_scalars[2];
    c.flen_max = c.h_scalars[3];
    return 0;
}

}  // namespace tdc
