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
//   1. key(i) = base-sigma number of the first k symbols of suffix i, k maximal with sigma^k <= 2^64 (zero padded past
//      the end; safe because the sentinel is unique).  One radix sort of (key, i) orders all suffixes by their k-prefix.
//   2. rerank: group = run of equal keys; rank = SA slot of the group head; singleton groups are final and leave the
//      working set ("active" list), the rest keep (slot, suffix, dense group id).
//   3. doubling round with offset h: key = (group id << rbits) | rank[suffix + h]; sort the active list; rerank; h *= 2.
//      Active suffixes always satisfy suffix + h <= n-1 (a suffix whose h-prefix reaches the sentinel is unique).
//   When the active list is empty, sa[] is the suffix array and rank[] is the inverse suffix array.
#include <cmath>
#include <cstdlib>

#include "tdc_ctx.h"

namespace tdc {

// ---------------------------------------------------------------------------------------------------------------
// 0. byte histogram
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) byte_histogram_kernel(const uint8_t* __restrict__ text, u64 n, u32* __restrict__ ghist) {
    __shared__ u32 sh[8 * 256];  // one private histogram per warp: repeated bytes are the norm in text
    for (u32 i = threadIdx.x; i < 8 * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    u32* my = sh + warp_id() * 256;
    const u64 nvec = (n + 15) / 16;  // text is zero padded to a multiple of 16
    const uint4* tv = reinterpret_cast<const uint4*>(text);
    for (u64 v = u64(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += u64(gridDim.x) * blockDim.x) {
        const uint4 q = __ldg(tv + v);
        const u32 wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (v * 16 + j * 4 + b < n) atomicAdd(&my[(wds[j] >> (8 * b)) & 0xff], 1u);
            }
        }
    }
    __syncthreads();
    for (u32 d = threadIdx.x; d < 256; d += blockDim.x) {
        u32 s = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += sh[w * 256 + d];
        if (s) atomicAdd(&ghist[d], s);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 1. packed initial keys
// ---------------------------------------------------------------------------------------------------------------
struct PackParams {
    u32 sigma;  // alphabet size incl. sentinel
    u32 k;      // symbols per key
    u64 top;    // sigma^(k-1)
};

static const int PK_THREADS = 256;
static const int PK_IPT = 8;
static const int PK_TILE = PK_THREADS * PK_IPT;
static const int PK_HALO = 64;  // k <= 64 (sigma >= 2)
static const int SA_RESIDUE_LOG2 = 10;  // initial sort aims to leave ~n / 2^10 suffixes to the doubling rounds

__global__ void __launch_bounds__(PK_THREADS)
pack_keys_kernel(const uint8_t* __restrict__ text, u64 n, const uint8_t* __restrict__ code_map, PackParams pp,
                 u64* __restrict__ keys) {
    __shared__ uint8_t codes[PK_TILE + PK_HALO];
    __shared__ uint8_t cmap[256];
    cmap[threadIdx.x] = code_map[threadIdx.x];
    __syncthreads();
    const u64 base = u64(blockIdx.x) * PK_TILE;
    for (u32 j = threadIdx.x; j < PK_TILE + PK_HALO; j += PK_THREADS) {
        const u64 p = base + j;
        codes[j] = p < n ? cmap[text[p]] : uint8_t(0);
    }
    __syncthreads();
    const u32 l0 = threadIdx.x * PK_IPT;
    u64 key = 0;
    for (u32 j = 0; j < pp.k; j++) key = key * pp.sigma + codes[l0 + j];
    u64 out[PK_IPT];
#pragma unroll
    for (int q = 0; q < PK_IPT; q++) {
        out[q] = key;
        key = (key - u64(codes[l0 + q]) * pp.top) * pp.sigma + codes[l0 + q + pp.k];  // roll one symbol
    }
#pragma unroll
    for (int q = 0; q < PK_IPT; q++) {
        const u64 p = base + l0 + q;
        if (p < n) keys[p] = out[q];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. rerank (reduce -> scan of tile aggregates -> apply)
// ---------------------------------------------------------------------------------------------------------------
static const int RR_THREADS = 256;
static const int RR_IPT = 8;
static const int RR_TILE = RR_THREADS * RR_IPT;

// flags for the RR_IPT consecutive elements owned by this thread.
// head bit q: element t0+q starts a group; ns bit q: its group has more than one member.
template <class K>
__device__ __forceinline__ void rr_flags(const K* __restrict__ keys, u64 m, u64 t0, u32* head_bits, u32* ns_bits) {
    K kv[RR_IPT + 2];
#pragma unroll
    for (int q = 0; q < RR_IPT + 2; q++) {
        const u64 t = t0 + q;  // kv[q] = key[t0 - 1 + q]
        kv[q] = (t >= 1 && t - 1 < m) ? keys[t - 1] : K(0);
    }
    u32 hb = 0;  // bit q (0..RR_IPT) = head(t0+q), with head(m) := 1 and head(0) := 1
#pragma unroll
    for (int q = 0; q <= RR_IPT; q++) {
        const u64 t = t0 + q;
        const bool h = (t == 0) || (t >= m) || (kv[q] != kv[q + 1]);
        hb |= u32(h) << q;
    }
    u32 nb = 0;
#pragma unroll
    for (int q = 0; q < RR_IPT; q++) {
        const bool single = ((hb >> q) & 1u) && ((hb >> (q + 1)) & 1u);
        if (t0 + q < m && !single) nb |= 1u << q;
    }
    *head_bits = hb & ((1u << RR_IPT) - 1u);
    *ns_bits = nb;
}

template <class K>
__global__ void __launch_bounds__(RR_THREADS)
rerank_reduce_kernel(const K* __restrict__ keys, u64 m, u32* __restrict__ agg_lasthead, ull* __restrict__ agg_cnt) {
    __shared__ u32 s_max[RR_THREADS / 32];
    __shared__ ull s_sum[RR_THREADS / 32];
    const u64 t0 = u64(blockIdx.x) * RR_TILE + u64(threadIdx.x) * RR_IPT;
    u32 hb, nb;
    rr_flags<K>(keys, m, t0, &hb, &nb);
    u32 lasthead = 0;  // (index + 1) of the last head owned by this thread, 0 if none
#pragma unroll
    for (int q = 0; q < RR_IPT; q++)
        if (((hb >> q) & 1u) && t0 + q < m) lasthead = u32(t0 + q) + 1u;
    ull cnt = ull(__popc(nb)) | (ull(__popc(nb & hb)) << 32);
    lasthead = warp_max(lasthead);
    cnt = warp_sum<ull>(cnt);
    if (lane_id() == 0) { s_max[warp_id()] = lasthead; s_sum[warp_id()] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 mx = 0; ull sm = 0;
        for (int w = 0; w < RR_THREADS / 32; w++) { mx = max(mx, s_max[w]); sm += s_sum[w]; }
        agg_lasthead[blockIdx.x] = mx;
        agg_cnt[blockIdx.x] = sm;
    }
}

// single CTA: exclusive scan of the per-tile aggregates; totals[0] = #non-singleton elements, totals[1] = #their groups
__global__ void __launch_bounds__(1024)
rerank_scan_kernel(u32* __restrict__ agg_lasthead, ull* __restrict__ agg_cnt, u32 ntiles, u32* __restrict__ totals) {
    __shared__ ull scratch_s[33];
    __shared__ u32 scratch_m[33];
    ull carry_s = 0;
    u32 carry_m = 0;
    for (u32 b = 0; b < ntiles; b += 1024) {
        const u32 i = b + threadIdx.x;
        const ull c = i < ntiles ? agg_cnt[i] : 0;
        const u32 h = i < ntiles ? agg_lasthead[i] : 0;
        ull tot_s; u32 tot_m;
        const ull ex_s = block_exclusive_sum<ull>(c, scratch_s, &tot_s);
        const u32 in_m = block_inclusive_max(h, scratch_m, &tot_m);
        // exclusive max = inclusive max of the previous element
        u32 ex_m = __shfl_up_sync(kFull, in_m, 1);
        __shared__ u32 warp_last[32];
        if (lane_id() == 31) warp_last[warp_id()] = in_m;
        __syncthreads();
        if (lane_id() == 0) ex_m = warp_id() ? warp_last[warp_id() - 1] : 0u;
        __syncthreads();
        if (i < ntiles) {
            agg_cnt[i] = carry_s + ex_s;
            agg_lasthead[i] = max(carry_m, ex_m);
        }
        carry_s += tot_s;
        carry_m = max(carry_m, tot_m);
    }
    if (threadIdx.x == 0) {
        totals[0] = u32(carry_s);
        totals[1] = u32(carry_s >> 32);
    }
}

template <class K>
__global__ void __launch_bounds__(RR_THREADS)
rerank_apply_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ pos_in, u64 m,
                    const u32* __restrict__ pre_lasthead, const ull* __restrict__ pre_cnt, u32* __restrict__ sa,
                    u32* __restrict__ rank_idx, u32* __restrict__ rank_val, u32* __restrict__ pos_out,
                    u32* __restrict__ idx_out, u32* __restrict__ gid_out) {
    __shared__ ull scratch_s[33];
    __shared__ u32 scratch_m[33];
    const u64 t0 = u64(blockIdx.x) * RR_TILE + u64(threadIdx.x) * RR_IPT;
    u32 hb, nb;
    rr_flags<K>(keys, m, t0, &hb, &nb);
    u32 lasthead = 0;
#pragma unroll
    for (int q = 0; q < RR_IPT; q++)
        if (((hb >> q) & 1u) && t0 + q < m) lasthead = u32(t0 + q) + 1u;
    const ull cnt = ull(__popc(nb)) | (ull(__popc(nb & hb)) << 32);
    ull tot_s; u32 tot_m;
    ull run_s = pre_cnt[blockIdx.x] + block_exclusive_sum<ull>(cnt, scratch_s, &tot_s);
    // head index (+1) in force before this thread's first element
    u32 in_m = block_inclusive_max(lasthead, scratch_m, &tot_m);
    u32 ex_m = __shfl_up_sync(kFull, in_m, 1);
    __shared__ u32 warp_last[32];
    if (lane_id() == 31) warp_last[warp_id()] = in_m;
    __syncthreads();
    if (lane_id() == 0) ex_m = warp_id() ? warp_last[warp_id() - 1] : 0u;
    u32 cur_head = max(pre_lasthead[blockIdx.x], ex_m);
#pragma unroll
    for (int q = 0; q < RR_IPT; q++) {
        const u64 t = t0 + q;
        if (t >= m) break;
        const bool h = (hb >> q) & 1u, ns = (nb >> q) & 1u;
        if (h) cur_head = u32(t) + 1u;
        const u32 hidx = cur_head - 1u;
        const u32 slot = pos_in ? pos_in[t] : u32(t);
        const u32 headslot = pos_in ? pos_in[hidx] : hidx;
        const u32 sfx = vals[t];
        rank_idx[t] = sfx;  // rank[sfx] = headslot, applied by the partitioned scatter that follows
        rank_val[t] = headslot;
        if (ns) {
            if (h) run_s += ull(1) << 32;
            const u32 o = u32(run_s);
            pos_out[o] = slot;
            idx_out[o] = sfx;
            gid_out[o] = u32(run_s >> 32) - 1u;
            run_s += 1;
        } else {
            sa[slot] = sfx;  // singleton group: final position
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. doubling keys
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_keys_kernel(const u32* __restrict__ idx, const u32* __restrict__ gid, const u32* __restrict__ rank, u64 m, u64 h,
                  u64 n, u32 rbits, u64* __restrict__ keys) {
    const u64 o = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o >= m) return;
    const u64 j = u64(idx[o]) + h;
    const u32 r2 = j < n ? rank[j] : 0u;  // j < n always holds for active suffixes; the guard is defensive
    keys[o] = (u64(gid[o]) << rbits) | r2;
}

// sort workspace (kept in this translation unit so that the shared-memory opt-in applies to the kernels launched here)
int sort_workspace_init(SortWorkspace& ws, u64 max_elems, int sm_count) {
    sort_workspace_free(ws);
    ws.sm_count = sm_count;
    // tiles of the smallest tile configuration bound the descriptor count
    const u64 t64 = rs_tiles<u64>(max_elems), t32 = rs_tiles<u32>(max_elems);
    ws.max_tiles = (t64 > t32 ? t64 : t32) + 1;
    TDC_CUDA(cudaMalloc(&ws.hist, sizeof(u32) * RS_MAX_PASSES * RS_RADIX));
    TDC_CUDA(cudaMalloc(&ws.uniform, sizeof(u32) * RS_MAX_PASSES));
    TDC_CUDA(cudaMalloc(&ws.desc, sizeof(ull) * ws.max_tiles * RS_RADIX));
    TDC_CUDA(cudaMemset(ws.desc, 0, sizeof(ull) * ws.max_tiles * RS_RADIX));
    TDC_CUDA(cudaMallocHost(&ws.h_uniform, sizeof(u32) * RS_MAX_PASSES));
    ws.epoch = 0;
    {
        auto k1 = rs_onesweep_kernel<u64, true>;
        auto k2 = rs_onesweep_kernel<u64, false>;
        TDC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<u64>())));
        TDC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<u64>())));
        auto k3 = rs_onesweep_kernel<u32, true>;
        auto k4 = rs_onesweep_kernel<u32, false>;
        TDC_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<u32>())));
        TDC_CUDA(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<u32>())));
    }
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
template <class K>
static int rerank(Ctx& c, const K* keys, const u32* vals, const u32* pos_in, u64 m, u32* agg_lasthead, ull* agg_cnt,
                  u32* pos_out, u32* idx_out, u32* gid_out, u32* sc_idx[2], u32* sc_val[2], u64* m_out, u64* g_out) {
    const u32 ntiles = u32(div_up(m, RR_TILE));
    auto rerank_reduce = rerank_reduce_kernel<K>;
    TDC_LAUNCH(rerank_reduce, ntiles, RR_THREADS, 0, c.stream, keys, m, agg_lasthead, agg_cnt);
    TDC_LAUNCH(rerank_scan_kernel, 1, 1024, 0, c.stream, agg_lasthead, agg_cnt, ntiles, c.d_scalars);
    auto k3 = rerank_apply_kernel<K>;
    auto rerank_apply = k3;
    TDC_LAUNCH(rerank_apply, ntiles, RR_THREADS, 0, c.stream, keys, vals, pos_in, m, agg_lasthead, agg_cnt, c.d_sa, sc_idx[0],
               sc_val[0], pos_out, idx_out, gid_out);
    TDC_KCHECK();
    TDC_TRY(partitioned_scatter(c.sortws, c.stream, sc_idx, sc_val, m, c.d_isa, c.n));
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 2 * sizeof(u32), cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    *m_out = c.h_scalars[0];
    *g_out = c.h_scalars[1];
    return 0;
}

int build_suffix_array(Ctx& c) {
    const u64 n = c.n;
    cudaStream_t st = c.stream;
    c.sa_rounds = 0;
    c.sa_active_sum = 0;
    c.sa_prefix_work = 0;
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
    TDC_CUDA(cudaStreamSynchronize(st));
    const u32* hist = c.h_scalars + 16;
    if (hist[0] != 1) {
        // mirrors TextDS's sentinel requirement (/root/reference/include/tudocomp/ds/TextDS.hpp:132-138) and the
        // escape-0 input restriction of SADivSufSort (ds/SADivSufSort.hpp:21-26)
        set_error("text must contain exactly one 0 byte, at its end (found %u)", hist[0]);
        return -3;
    }
    uint8_t code_map[256];
    u32 sigma = 0;
    for (int b = 0; b < 256; b++) {
        code_map[b] = uint8_t(sigma);
        if (hist[b]) sigma++;
    }
    PackParams pp;
    pp.sigma = sigma;
    pp.k = 1;
    pp.top = 1;
    {
        // largest k with sigma^k - 1 representable in 64 bits
        unsigned __int128 pw = sigma;
        while (pw * sigma <= (((unsigned __int128)1) << 64) && pp.k < u32(PK_HALO)) {
            pw *= sigma;
            pp.top *= sigma;
            pp.k++;
        }
    }
    // Fewer symbols per key = fewer radix passes over all n suffixes; the price is a larger active set in the doubling
    // rounds.  For a memoryless source a k-symbol prefix is shared with another suffix with probability about
    // n * 2^(-H0 k) (H0 = empirical order-0 entropy from the byte histogram), so k is chosen to leave ~2^-10 of the
    // suffixes unresolved, then rounded up to the most symbols that fit the same number of 8-bit passes.  A wrong guess
    // (text with memory) only moves work into the doubling rounds; the result is the same.
    {
        double h0 = 0;
        for (int b = 0; b < 256; b++)
            if (hist[b]) { const double pr = double(hist[b]) / double(n); h0 -= pr * log2(pr); }
        u32 k_need = pp.k;
        if (h0 > 1e-6) {
            const double want = ceil((log2(double(n)) + double(SA_RESIDUE_LOG2)) / h0);
            if (want < double(pp.k)) k_need = u32(want < 1 ? 1 : want);
        }
        if (const char* e = getenv("TDCGPU_SA_SYMBOLS")) {  // tuning/debug override
            const long v = atol(e);
            if (v >= 1 && v <= long(pp.k)) k_need = u32(v);
            else k_need = pp.k;
        } else {
            const double l2s = log2(double(sigma));
            const u32 passes = u32(ceil(ceil(k_need * l2s) / 8.0));
            while (k_need < pp.k && u32(ceil(ceil((k_need + 1) * l2s + 1e-9) / 8.0)) <= passes) k_need++;
        }
        if (k_need < pp.k) {
            pp.k = k_need;
            pp.top = 1;
            for (u32 j = 1; j < pp.k; j++) pp.top *= sigma;
        }
    }
    u32 sigbits;
    {
        unsigned __int128 pw = 1;
        for (u32 j = 0; j < pp.k; j++) pw *= sigma;
        pw -= 1;
        sigbits = 0;
        while (pw) { sigbits++; pw >>= 1; }
        if (sigbits == 0) sigbits = 1;
    }
    c.alphabet = sigma;
    c.symbols_per_key = pp.k;

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
    TDC_LAUNCH(pack_keys_kernel, u32(div_up(n, PK_TILE)), PK_THREADS, 0, st, c.d_text, n, d_code_map, pp, keys[0]);
    TDC_KCHECK();
    int res = 0;
    TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, keys, vals, n, 0, int(sigbits), true, &res));
    u64 m = 0, g = 0;
    int pcur = 0;
    // the sorted (key, suffix) pairs are in slot `res`; compacted survivors go to the other slot's value buffer
    {
        // the key buffer that does not hold the sorted keys is free: it carries the first (idx, val) pair buffer
        u32* sc_idx[2] = {reinterpret_cast<u32*>(keys[res ^ 1]), sc_a};
        u32* sc_val[2] = {reinterpret_cast<u32*>(keys[res ^ 1]) + n, sc_b};
        TDC_TRY(rerank<u64>(c, keys[res], vals[res], nullptr, n, agg_lasthead, agg_cnt, pos[pcur], vals[res ^ 1], gid, sc_idx, sc_val, &m, &g));
    }
    c.sa_rounds = 1;
    c.sa_active_sum = n;

    const u32 rbits = bits_for_host(n - 1);
    u64 h = pp.k;
    while (m > 0) {
        // active list: slots pos[pcur][0..m), suffixes vals[res^1][0..m), group ids gid[0..m)
        u64* k2[2] = {keys[0], keys[1]};
        u32* v2[2] = {vals[res ^ 1], vals[res]};
        TDC_LAUNCH(build_keys_kernel, u32(div_up(m, 256)), 256, 0, st, v2[0], gid, c.d_isa, m, h, n, rbits, k2[0]);
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
        TDC_TRY(rerank<u64>(c, k2[r2], v2[r2], pos[pcur], m, agg_lasthead, agg_cnt, pos[pcur ^ 1], v2[r2 ^ 1], gid, sc_idx, sc_val, &m_new, &g_new));
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
