// Suffix array + inverse suffix array by prefix doubling on the GPU (sm_100a).
//
// Replaces, with identical results, the reference's
//   SADivSufSort  (/root/reference/include/tudocomp/ds/SADivSufSort.hpp:28-51 -> util/divsufsort.hpp:253-279) and
//   ISAFromSA     (/root/reference/include/tudocomp/ds/ISAFromSA.hpp:37-39).
// The SA of a text with a unique smallest sentinel is a function of the text (unsigned byte order,
// util/divsufsort/divsufsort_def.hpp:12), so a different algorithm is bit-exact as long as it sorts correctly.
//
// Algorithm
//   0. byte histogram -> dense symbol codes (sentinel 0 -> code 0), sigma = #distinct bytes.
//   1. key(i) = the first k symbols of suffix i, b bits each (zero padded past the end).  Real symbols are coded
//      1..s with 0 for the sentinel; when s is a power of two they are coded 0..s-1 instead (one bit less per symbol)
//      and a trailing length field min(k, n-1-i) keeps suffixes that run into the sentinel apart ("shorter first").
//      k comes from a cost model (radix passes over n keys vs. suffixes left to the doubling rounds).  One radix sort
//      of (key, i) orders all suffixes by their k-prefix; adjacent distinct keys also give that pair's LCP (count of
//      equal leading symbols), so only pairs inside a group need character comparisons later.
//   2. rerank: group = run of equal keys; rank = SA slot of the group head; singleton groups are final and leave the
//      working set ("active" list), the rest keep (slot, suffix, dense group id).
//   3. doubling round with offset h: key = (group id << rbits) | rank[suffix + h]; sort the active list; rerank; h *= 2.
//      Active suffixes always satisfy suffix + h <= n-1 (a suffix whose h-prefix reaches the sentinel is unique).
//   When the active list is empty, sa[] is the suffix array and rank[] is the inverse suffix array.
#include <algorithm>

#include "sa_kernels.cuh"

namespace tdc {

// sort workspace
int sort_workspace_init(SortWorkspace& ws, u64 max_elems, int sm_count) {
    sort_workspace_free(ws);
    ws.sm_count = sm_count;
    // tiles of the smallest tile configuration bound the descriptor count
    const u64 t64 = rs_tiles<u64>(max_elems), t32 = rs_tiles<u32>(max_elems), tk = rs_tiles_keys(max_elems);
    ws.max_tiles = std::max(std::max(t64, t32), tk) + 1;
    TDC_CUDA(cudaMalloc(&ws.hist, sizeof(u32) * RS_MAX_PASSES * RS_RADIX));
    TDC_CUDA(cudaMalloc(&ws.uniform, sizeof(u32) * RS_MAX_PASSES));
    TDC_CUDA(cudaMalloc(&ws.desc, sizeof(ull) * ws.max_tiles * RS_RADIX));
    TDC_CUDA(cudaMemset(ws.desc, 0, sizeof(ull) * ws.max_tiles * RS_RADIX));
    TDC_CUDA(cudaMallocHost(&ws.h_uniform, sizeof(u32) * RS_MAX_PASSES));
    ws.epoch = 0;
    return 0;
}

void sort_workspace_free(SortWorkspace& ws) {
    if (ws.hist) cudaFree(ws.hist);
    if (ws.uniform) cudaFree(ws.uniform);
    if (ws.desc) cudaFree(ws.desc);
    if (ws.h_uniform) cudaFreeHost(ws.h_uniform);
    ws.desc = nullptr;
    ws.h_uniform = nullptr;
    ws.max_tiles = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
template <class K, bool FIRST>
static int rerank(Ctx& c, const K* keys, const u32* vals, const u32* pos_in, u64 m, u32* agg_lasthead, ull* agg_cnt,
                  u32* pos_out, u32* idx_out, u32* gid_out, u32* sc_idx[2], u32* sc_val[2], u32* lcp_out, PackParams pp,
                  u64* m_out, u64* g_out) {
    const u32 ntiles = u32(div_up(m, RR_TILE));
    auto rerank_reduce = rerank_reduce_kernel<K>;
    TDC_LAUNCH(rerank_reduce, ntiles, RR_THREADS, 0, c.stream, keys, m, agg_lasthead, agg_cnt, FIRST ? pp.idx_bits : 0u);
    prof_add_bytes("rerank_reduce", double(m) * sizeof(K));
    TDC_LAUNCH(rerank_scan_kernel, 1, 1024, 0, c.stream, agg_lasthead, agg_cnt, ntiles, c.d_scalars);
    auto rerank_apply = rerank_apply_kernel<K, FIRST>;
    TDC_LAUNCH(rerank_apply, ntiles, RR_THREADS, 0, c.stream, keys, vals, pos_in, m, agg_lasthead, agg_cnt, c.d_sa, sc_idx[0],
               sc_val[0], pos_out, idx_out, gid_out, lcp_out, pp, 0u, u32(c.n));
    prof_add_bytes("rerank_apply", double(m) * (sizeof(K) + ((FIRST && pp.idx_bits) ? 0 : 4) + (FIRST ? 8 + (lcp_out ? 4 : 0) : 12)));
    TDC_KCHECK();
    // first round: the pairs are (vals[t], head slot) and vals is a permutation of 0..n-1
    TDC_TRY(partitioned_scatter(c.sortws, c.stream, sc_idx, sc_val, m, c.d_isa, c.n, FIRST));
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 2 * sizeof(u32), cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    *m_out = c.h_scalars[0];
    *g_out = c.h_scalars[1];
    return 0;
}

#ifdef TDC_CUSIM
static const u32 SA_SAMPLES = 1u << 10;  // small, so that the CPU tests exercise the sampled model
static const u64 SA_SAMPLE_MIN_N = 3000;
#else
static const u32 SA_SAMPLES = 1u << 16;           // sampled suffix prefixes for the key-length cost model
static const u64 SA_SAMPLE_MIN_N = u64(1) << 22;  // smaller texts: the sampling would cost more than it can save
#endif

int build_suffix_array(Ctx& c, bool want_lcp) {
    const u64 n = c.n;
    cudaStream_t st = c.stream;
    c.sa_rounds = 0;
    c.sa_active_sum = 0;
    c.sa_prefix_work = 0;
    c.sa_lcp_seeded = false;
    c.sa_packed = false;
    c.sa_tail_span = 0;
    c.sa_key_bits = 0;
    c.sortws.stat_passes = 0;
    c.sortws.stat_elems = 0;
    if (n == 1) {  // text == "\0"
        TDC_CUDA(cudaMemsetAsync(c.d_sa, 0, 4, st));
        TDC_CUDA(cudaMemsetAsync(c.d_isa, 0, 4, st));
        return 0;
    }
    // ---- alphabet ----
    u32* d_hist = c.d_scalars + 16;  // 256 counters
    TDC_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(u32), st));
    {
        const u32 grid = u32(min(u64(c.sm_count) * 8, div_up(div_up(n, 16), 256)));
        TDC_LAUNCH(byte_histogram_kernel, grid, 256, 0, st, c.d_text, n, d_hist);
        TDC_KCHECK();
    }
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 16, d_hist, 256 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    c.h_scalars[16 + 256] = 0xff;
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 16 + 256, c.d_text + (n - 1), 1, cudaMemcpyDeviceToHost, st));  // the last text byte
    TDC_CUDA(cudaStreamSynchronize(st));
    const u32* hist = c.h_scalars + 16;
    if (hist[0] != 1 || (c.h_scalars[16 + 256] & 0xffu) != 0) {
        // mirrors TextDS's sentinel requirement (/root/reference/include/tudocomp/ds/TextDS.hpp:132-138: the LAST byte is
        // the 0) and the escape-0 input restriction of SADivSufSort (ds/SADivSufSort.hpp:21-26: no other 0)
        set_error("text must contain exactly one 0 byte, at its end (found %u zero bytes, last byte 0x%02x)", hist[0],
                  c.h_scalars[16 + 256] & 0xffu);
        return -3;
    }
    PackParams pp;
    u32 sigbits;
    double log2_collide[9];
    bool have_collide = false;
    if (n >= SA_SAMPLE_MIN_N && !getenv("TDCGPU_SA_SYMBOLS")) {
        // how fast do k-byte prefixes separate suffixes of THIS text?  sort a regular sample of 8-byte prefixes
        c.arena.reset();
        u64* sk[2] = {c.arena.take<u64>(SA_SAMPLES), c.arena.take<u64>(SA_SAMPLES)};
        u32* sv[2] = {c.arena.take<u32>(SA_SAMPLES), c.arena.take<u32>(SA_SAMPLES)};
        if (sk[0] && sk[1] && sv[0] && sv[1]) {
            TDC_LAUNCH(sample_prefix_kernel, u32(div_up(SA_SAMPLES, 256)), 256, 0, st, c.d_text, n, SA_SAMPLES, sk[0]);
            TDC_KCHECK();
            int sres = 0;
            TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, sk, sv, SA_SAMPLES, 0, 64, true, &sres));
            c.h_samples.resize(SA_SAMPLES);
            TDC_CUDA(cudaMemcpyAsync(c.h_samples.data(), sk[sres], sizeof(u64) * SA_SAMPLES, cudaMemcpyDeviceToHost, st));
            TDC_CUDA(cudaStreamSynchronize(st));
            prefix_collision_table(c.h_samples.data(), SA_SAMPLES, log2_collide);
            have_collide = true;
        }
    }
    // packed records need the suffix index next to the key bits in one 64-bit word
    choose_key_layout(hist, n, &pp, &sigbits, have_collide ? log2_collide : nullptr, bits_for_host(n - 1));
    const bool packed = pp.idx_bits != 0;
    uint8_t code_map[256];
    u32 sigma = 1;
    {
        u32 next = pp.pow2 ? 0 : 1;
        code_map[0] = 0;
        for (int b = 1; b < 256; b++) {
            code_map[b] = uint8_t(next);
            if (hist[b]) { next++; sigma++; }
        }
    }
    c.alphabet = sigma;
    c.symbols_per_key = pp.known;
    c.sa_packed = packed;
    c.sa_tail_span = (packed && pp.pow2) ? pp.k : 0;
    c.sa_key_bits = sigbits;

    // ---- scratch ----
    c.arena.reset();
    u64* keys[2] = {c.arena.take<u64>(n), c.arena.take<u64>(n)};
    u32* vals[2] = {c.arena.take<u32>(n), c.arena.take<u32>(n)};
    u32* pos[2] = {c.arena.take<u32>(n), c.arena.take<u32>(n)};
    u32* gid = c.arena.take<u32>(n);
    u32* sc_a = c.arena.take<u32>(n);  // second buffer pair of the partitioned ISA scatter
    u32* sc_b = c.arena.take<u32>(n);
    const u64 rr_tiles = div_up(n, RR_TILE);
    u32* agg_lasthead = c.arena.take<u32>(rr_tiles);
    ull* agg_cnt = c.arena.take<ull>(rr_tiles);
    uint8_t* d_code_map = c.arena.take<uint8_t>(256);
    if (!keys[0] || !keys[1] || !vals[0] || !vals[1] || !pos[0] || !pos[1] || !gid || !sc_a || !sc_b || !agg_lasthead || !agg_cnt || !d_code_map) {
        set_error("suffix array: scratch arena too small");
        return -2;
    }
    TDC_CUDA(cudaMemcpyAsync(d_code_map, code_map, 256, cudaMemcpyHostToDevice, st));

    // ---- initial sort by k-symbol prefix ----
    // EXPERIMENT (TDCGPU_SA_FUSED_HIST=1): the digit histograms of the sort counted by the pack kernel.  Measured at
    // dna 2^30 (profiles/r2_summary.md): the fused kernel takes 10.4 ms, pack 4.05 + histogram 5.77 ms apart — both are
    // bound by the shared-memory atomics (6 per key), not by the 8 B/key read that the fusion saves.  Off by default.
    const bool fused_hist = !packed && getenv("TDCGPU_SA_FUSED_HIST") != nullptr;
    if (fused_hist) {
        TDC_CUDA(cudaMemsetAsync(c.sortws.hist, 0, sizeof(u32) * RS_MAX_PASSES * RS_RADIX, st));
        auto pack_keys_kernel = tdc::pack_keys_kernel<true>;
        TDC_LAUNCH(pack_keys_kernel, u32(div_up(n, PK_TILE)), PK_THREADS, 0, st, c.d_text, n, d_code_map, pp, keys[0], u64(0), n,
                   make_pass_plan(0, int(sigbits)), c.sortws.hist);
    } else {
        auto pack_keys_kernel = tdc::pack_keys_kernel<false>;
        TDC_LAUNCH(pack_keys_kernel, u32(div_up(n, PK_TILE)), PK_THREADS, 0, st, c.d_text, n, d_code_map, pp, keys[0], u64(0), n, PassPlan(), (u32*)nullptr);
    }
    prof_add_bytes("pack_keys_kernel", double(n) * 9);
    TDC_KCHECK();
    int res = 0;
    if (packed) TDC_TRY(radix_sort_keys(c.sortws, st, keys, n, int(pp.idx_bits), int(pp.idx_bits + sigbits), &res));
    else TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, keys, vals, n, 0, int(sigbits), true, &res, false, fused_hist));
    u64 m = 0, g = 0;
    int pcur = 0;
    // the sorted (key, suffix) pairs are in slot `res`; compacted survivors go to the other slot's value buffer
    {
        // ISA updates of this round: index side = the sorted suffix list itself (read only), value side and the
        // partition scratch live in the key buffer that does not hold the sorted keys plus sc_a/sc_b
        // (packed records: there is no value array; the sorted suffix list is what rerank_apply writes to sa[])
        u32* sc_idx[2] = {packed ? c.d_sa : vals[res], sc_a};
        u32* sc_val[2] = {reinterpret_cast<u32*>(keys[res ^ 1]), sc_b};
        u32* lcp_out = (want_lcp && c.d_lcp) ? c.d_lcp : nullptr;
        TDC_TRY((rerank<u64, true>(c, keys[res], packed ? nullptr : vals[res], nullptr, n, agg_lasthead, agg_cnt, pos[pcur], vals[res ^ 1], gid, sc_idx, sc_val, lcp_out, pp, &m, &g)));
        c.sa_lcp_seeded = lcp_out != nullptr;
    }
    c.sa_rounds = 1;
    c.sa_active_sum = n;
    c.sa_first_residue = m;

    // keys without a length field over a power-of-two alphabet: suffixes that ran into the sentinel are ranked by their
    // overshoot, below all real ranks (build_keys_kernel), which takes one more bit
    const u32 overshoot = (packed && pp.pow2) ? 1u : 0u;
    const u32 rbits = overshoot ? bits_for_host(2 * n - 1) : bits_for_host(n - 1);
    u64 h = min(u64(pp.known), n - 1);  // (a shorter known prefix is always valid; the overshoot ranks need h <= n - 1)
    while (m > 0) {
        // active list: slots pos[pcur][0..m), suffixes vals[res^1][0..m), group ids gid[0..m)
        u64* k2[2] = {keys[0], keys[1]};
        u32* v2[2] = {vals[res ^ 1], vals[res]};
        TDC_LAUNCH(build_keys_kernel, u32(div_up(m, 256)), 256, 0, st, v2[0], gid, c.d_isa, m, h, n, rbits, k2[0], overshoot);
        prof_add_bytes("build_keys_kernel", double(m) * 20);
        TDC_KCHECK();
        const int gbits = g > 1 ? int(bits_for_host(g - 1)) : 0;
        int r2 = 0;
        TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, k2, v2, m, 0, int(rbits) + gbits, false, &r2));
        c.sa_rounds++;
        c.sa_active_sum += m;
        c.sa_prefix_work += double(m) * double(h);
        u64 m_new = 0, g_new = 0;
        // survivors are written to the value buffer that does not hold the sorted input
        u32* sc_idx[2] = {reinterpret_cast<u32*>(k2[r2 ^ 1]), sc_a};
        u32* sc_val[2] = {reinterpret_cast<u32*>(k2[r2 ^ 1]) + n, sc_b};
        TDC_TRY((rerank<u64, false>(c, k2[r2], v2[r2], pos[pcur], m, agg_lasthead, agg_cnt, pos[pcur ^ 1], v2[r2 ^ 1], gid, sc_idx, sc_val, nullptr, pp, &m_new, &g_new)));
        // re-point: next round's suffix list lives in v2[r2 ^ 1]
        if (v2[r2 ^ 1] == vals[res ^ 1]) {
            // already where the loop expects it
        } else {
            res ^= 1;
        }
        pcur ^= 1;
        m = m_new;
        g = g_new;
        h *= 2;
        if (h > 2 * n && m > 0) {
            set_error("suffix array: doubling did not converge (m=%llu)", (unsigned long long)m);
            return -4;
        }
    }
    return 0;
}

}  // namespace tdc
