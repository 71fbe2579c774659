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
    u32 b;        // bits per symbol
    u32 k;        // symbols per key
    u32 lenbits;  // width of the trailing length field (0: sentinel has its own code 0, no field needed)
};

static const int PK_THREADS = 256;
static const int PK_IPT = 8;
static const int PK_TILE = PK_THREADS * PK_IPT;
static const int PK_HALO = 64;  // k <= 64 (b >= 1)

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
    const u32 width = pp.b * pp.k;  // <= 64
    const u64 mask = width >= 64 ? ~u64(0) : ((u64(1) << width) - 1);
    u64 packed = 0;
    for (u32 j = 0; j < pp.k; j++) packed = (packed << pp.b) | codes[l0 + j];
    u64 out[PK_IPT];
#pragma unroll
    for (int q = 0; q < PK_IPT; q++) {
        const u64 p = base + l0 + q;
        u64 key = packed;
        if (pp.lenbits) {
            const u64 len = p + 1 < n ? min(u64(pp.k), n - 1 - p) : u64(0);
            key = (packed << pp.lenbits) | len;
        }
        out[q] = key;
        packed = ((packed << pp.b) & mask) | codes[l0 + q + pp.k];  // roll one symbol
    }
    if (base + l0 + PK_IPT <= n) {
        ulonglong2* o2 = reinterpret_cast<ulonglong2*>(keys + base + l0);
#pragma unroll
        for (int q = 0; q < PK_IPT / 2; q++) o2[q] = make_ulonglong2(out[2 * q], out[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < PK_IPT; q++) {
            const u64 p = base + l0 + q;
            if (p < n) keys[p] = out[q];
        }
    }
}

// number of equal leading symbols of two different keys = LCP of the two suffixes when it is < k
__device__ __forceinline__ u32 key_common_symbols(u64 a, u64 c, PackParams pp) {
    const u64 pa = a >> pp.lenbits, pc = c >> pp.lenbits;
    u32 common = pp.k;
    if (pa != pc) common = (u32(__clzll((long long)(pa ^ pc))) - (64u - pp.b * pp.k)) / pp.b;
    if (pp.lenbits) {
        const u32 lm = (1u << pp.lenbits) - 1u;
        common = min(common, min(u32(a) & lm, u32(c) & lm));  // a suffix ends where its sentinel stands
    }
    return common;
}

// ---------------------------------------------------------------------------------------------------------------
// 2. rerank (reduce -> scan of tile aggregates -> apply)
// ---------------------------------------------------------------------------------------------------------------
static const int RR_THREADS = 256;
static const int RR_IPT = 8;
static const int RR_TILE = RR_THREADS * RR_IPT;

// RR_IPT consecutive keys of this thread plus one neighbour on either side: kv[q] = key[t0 - 1 + q] (0 outside [0, m)).
// Full threads use 16-byte loads (t0 is a multiple of RR_IPT and the arrays are 256-byte aligned).
template <class K>
__device__ __forceinline__ void rr_load_keys(const K* __restrict__ keys, u64 m, u64 t0, K* kv) {
    if (t0 + RR_IPT <= m) {
        if (sizeof(K) == 8) {
            const ulonglong2* v2 = reinterpret_cast<const ulonglong2*>(keys + t0);
#pragma unroll
            for (int q = 0; q < RR_IPT / 2; q++) {
                const ulonglong2 x = v2[q];
                kv[1 + 2 * q] = K(x.x);
                kv[2 + 2 * q] = K(x.y);
            }
        } else {
            const uint4* v4 = reinterpret_cast<const uint4*>(keys + t0);
#pragma unroll
            for (int q = 0; q < RR_IPT / 4; q++) {
                const uint4 x = v4[q];
                kv[1 + 4 * q] = K(x.x);
                kv[2 + 4 * q] = K(x.y);
                kv[3 + 4 * q] = K(x.z);
                kv[4 + 4 * q] = K(x.w);
            }
        }
        kv[0] = t0 >= 1 ? keys[t0 - 1] : K(0);
        kv[RR_IPT + 1] = t0 + RR_IPT < m ? keys[t0 + RR_IPT] : K(0);
    } else {
#pragma unroll
        for (int q = 0; q < RR_IPT + 2; q++) {
            const u64 t = t0 + q;
            kv[q] = (t >= 1 && t - 1 < m) ? keys[t - 1] : K(0);
        }
    }
}

// flags for the RR_IPT consecutive elements owned by this thread.
// head bit q: element t0+q starts a group; ns bit q: its group has more than one member.
template <class K>
__device__ __forceinline__ void rr_flags(const K* kv, u64 m, u64 t0, u32* head_bits, u32* ns_bits) {
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
    K kv[RR_IPT + 2];
    rr_load_keys<K>(keys, m, t0, kv);
    rr_flags<K>(kv, m, t0, &hb, &nb);
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

// FIRST = initial sort: slot == t, every suffix is written to sa[] (non-final ones are overwritten later), the ISA
// update pairs are (vals[t], head slot) so only the value side is materialised, and the LCP of adjacent distinct
// keys is taken from the keys themselves.  Later rounds: slots come from pos_in, pairs are materialised in full.
template <class K, bool FIRST>
__global__ void __launch_bounds__(RR_THREADS)
rerank_apply_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ pos_in, u64 m,
                    const u32* __restrict__ pre_lasthead, const ull* __restrict__ pre_cnt, u32* __restrict__ sa,
                    u32* __restrict__ rank_idx, u32* __restrict__ rank_val, u32* __restrict__ pos_out,
                    u32* __restrict__ idx_out, u32* __restrict__ gid_out, u32* __restrict__ lcp_out, PackParams pp) {
    __shared__ ull scratch_s[33];
    __shared__ u32 scratch_m[33];
    const u64 t0 = u64(blockIdx.x) * RR_TILE + u64(threadIdx.x) * RR_IPT;
    u32 hb, nb;
    K kv[RR_IPT + 2];
    rr_load_keys<K>(keys, m, t0, kv);
    rr_flags<K>(kv, m, t0, &hb, &nb);
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
    const bool full = t0 + RR_IPT <= m;
    u32 sfxv[RR_IPT], headv[RR_IPT], lcpv[RR_IPT];
    if (full) {
        const uint4* v4 = reinterpret_cast<const uint4*>(vals + t0);
#pragma unroll
        for (int q = 0; q < RR_IPT / 4; q++) {
            const uint4 x = v4[q];
            sfxv[4 * q] = x.x; sfxv[4 * q + 1] = x.y; sfxv[4 * q + 2] = x.z; sfxv[4 * q + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < RR_IPT; q++) sfxv[q] = t0 + q < m ? vals[t0 + q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < RR_IPT; q++) {
        const u64 t = t0 + q;
        headv[q] = 0;
        lcpv[q] = 0;
        if (t >= m) continue;
        const bool h = (hb >> q) & 1u, ns = (nb >> q) & 1u;
        if (h) cur_head = u32(t) + 1u;
        const u32 hidx = cur_head - 1u;
        const u32 slot = FIRST ? u32(t) : pos_in[t];
        const u32 headslot = FIRST ? hidx : pos_in[hidx];
        const u32 sfx = sfxv[q];
        headv[q] = headslot;  // rank[sfx] = headslot, applied by the partitioned scatter that follows
        if (FIRST) {
            if (sizeof(K) == 8 && lcp_out) lcpv[q] = t == 0 ? 0u : (h ? key_common_symbols(u64(kv[q]), u64(kv[q + 1]), pp) : LCP_UNKNOWN);
        } else {
            rank_idx[t] = sfx;
            rank_val[t] = headslot;
        }
        if (ns) {
            if (h) run_s += ull(1) << 32;
            const u32 o = u32(run_s);
            pos_out[o] = slot;
            idx_out[o] = sfx;
            gid_out[o] = u32(run_s >> 32) - 1u;
            run_s += 1;
        } else if (!FIRST) {
            sa[slot] = sfx;  // singleton group: final position
        }
    }
    if (FIRST) {
        if (full) {
#pragma unroll
            for (int q = 0; q < RR_IPT / 4; q++) {
                reinterpret_cast<uint4*>(rank_val + t0)[q] = make_uint4(headv[4 * q], headv[4 * q + 1], headv[4 * q + 2], headv[4 * q + 3]);
                reinterpret_cast<uint4*>(sa + t0)[q] = make_uint4(sfxv[4 * q], sfxv[4 * q + 1], sfxv[4 * q + 2], sfxv[4 * q + 3]);
                if (lcp_out) reinterpret_cast<uint4*>(lcp_out + t0)[q] = make_uint4(lcpv[4 * q], lcpv[4 * q + 1], lcpv[4 * q + 2], lcpv[4 * q + 3]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < RR_IPT; q++) {
                if (t0 + q < m) {
                    rank_val[t0 + q] = headv[q];
                    sa[t0 + q] = sfxv[q];
                    if (lcp_out) lcp_out[t0 + q] = lcpv[q];
                }
            }
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
template <class K, bool FIRST>
static int rerank(Ctx& c, const K* keys, const u32* vals, const u32* pos_in, u64 m, u32* agg_lasthead, ull* agg_cnt,
                  u32* pos_out, u32* idx_out, u32* gid_out, u32* sc_idx[2], u32* sc_val[2], u32* lcp_out, PackParams pp,
                  u64* m_out, u64* g_out) {
    const u32 ntiles = u32(div_up(m, RR_TILE));
    auto rerank_reduce = rerank_reduce_kernel<K>;
    TDC_LAUNCH(rerank_reduce, ntiles, RR_THREADS, 0, c.stream, keys, m, agg_lasthead, agg_cnt);
    prof_add_bytes("rerank_reduce", double(m) * sizeof(K));
    TDC_LAUNCH(rerank_scan_kernel, 1, 1024, 0, c.stream, agg_lasthead, agg_cnt, ntiles, c.d_scalars);
    auto rerank_apply = rerank_apply_kernel<K, FIRST>;
    TDC_LAUNCH(rerank_apply, ntiles, RR_THREADS, 0, c.stream, keys, vals, pos_in, m, agg_lasthead, agg_cnt, c.d_sa, sc_idx[0],
               sc_val[0], pos_out, idx_out, gid_out, lcp_out, pp);
    prof_add_bytes("rerank_apply", double(m) * (sizeof(K) + 4 + (FIRST ? 8 + (lcp_out ? 4 : 0) : 12)));
    TDC_KCHECK();
    // first round: the pairs are (vals[t], head slot) and vals is a permutation of 0..n-1
    TDC_TRY(partitioned_scatter(c.sortws, c.stream, sc_idx, sc_val, m, c.d_isa, c.n, FIRST));
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 2 * sizeof(u32), cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    *m_out = c.h_scalars[0];
    *g_out = c.h_scalars[1];
    return 0;
}

// Cost model for the number of symbols per initial key.  A radix pass moves 24 B per suffix; a suffix the initial sort
// leaves in a group costs about SA_ACTIVE_COST bytes of traffic in the doubling rounds (random rank gather, key
// build, up to 8 passes, re-rank, scatter).  For a memoryless source a k-symbol prefix is shared with another suffix
// with probability about n * 2^(-H0 k), H0 = order-0 entropy of the byte histogram.  A wrong guess (text with
// memory) only moves work between the two phases; the result is the same.
static const double SA_ACTIVE_COST = 400.0;

static u32 bits_of_lenfield(u32 k) { return bits_for_host(k); }

static void choose_key_layout(const u32* hist, u64 n, PackParams* pp, u32* sigbits) {
    u32 real = 0;
    double h0 = 0;
    for (int b = 1; b < 256; b++)
        if (hist[b]) {
            real++;
            const double pr = double(hist[b]) / double(n - 1);
            h0 -= pr * log2(pr);
        }
    const bool pow2 = (real & (real - 1)) == 0;  // real >= 1 here
    // power-of-two alphabets are coded 0..real-1 (the sentinel shares code 0) and need the length field
    u32 b = 1;
    while ((1u << b) < (pow2 ? real : real + 1)) b++;
    u32 kmax = 1;
    for (u32 k = 1; k <= u32(PK_HALO); k++)
        if (b * k + (pow2 ? bits_of_lenfield(k) : 0) <= 64) kmax = k;
    u32 best = kmax;
    if (const char* e = getenv("TDCGPU_SA_SYMBOLS")) {  // tuning/debug override
        const long v = atol(e);
        if (v >= 1 && v <= long(kmax)) best = u32(v);
    } else {
        double best_cost = 1e300;
        for (u32 k = 1; k <= kmax; k++) {
            const u32 bits = b * k + (pow2 ? bits_of_lenfield(k) : 0);
            const double passes = double((bits + 7) / 8);
            double residue = h0 > 1e-9 ? exp2(log2(double(n)) - h0 * double(k)) : 1.0;
            if (residue > 1.0) residue = 1.0;
            const double cost = passes * 24.0 + residue * SA_ACTIVE_COST;
            if (cost <= best_cost) { best_cost = cost; best = k; }  // ties: more symbols
        }
    }
    pp->b = b;
    pp->k = best;
    pp->lenbits = pow2 ? bits_of_lenfield(best) : 0;
    *sigbits = b * best + pp->lenbits;
}

int build_suffix_array(Ctx& c, bool want_lcp) {
    const u64 n = c.n;
    cudaStream_t st = c.stream;
    c.sa_rounds = 0;
    c.sa_active_sum = 0;
    c.sa_prefix_work = 0;
    c.sa_lcp_seeded = false;
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
    PackParams pp;
    u32 sigbits;
    choose_key_layout(hist, n, &pp, &sigbits);
    uint8_t code_map[256];
    u32 sigma = 1;
    {
        u32 next = pp.lenbits ? 0 : 1;
        code_map[0] = 0;
        for (int b = 1; b < 256; b++) {
            code_map[b] = uint8_t(next);
            if (hist[b]) { next++; sigma++; }
        }
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
    prof_add_bytes("pack_keys_kernel", double(n) * 9);
    TDC_KCHECK();
    int res = 0;
    TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, keys, vals, n, 0, int(sigbits), true, &res));
    u64 m = 0, g = 0;
    int pcur = 0;
    // the sorted (key, suffix) pairs are in slot `res`; compacted survivors go to the other slot's value buffer
    {
        // ISA updates of this round: index side = the sorted suffix list itself (read only), value side and the
        // partition scratch live in the key buffer that does not hold the sorted keys plus sc_a/sc_b
        u32* sc_idx[2] = {vals[res], sc_a};
        u32* sc_val[2] = {reinterpret_cast<u32*>(keys[res ^ 1]), sc_b};
        u32* lcp_out = (want_lcp && c.d_lcp) ? c.d_lcp : nullptr;
        TDC_TRY((rerank<u64, true>(c, keys[res], vals[res], nullptr, n, agg_lasthead, agg_cnt, pos[pcur], vals[res ^ 1], gid, sc_idx, sc_val, lcp_out, pp, &m, &g)));
        c.sa_lcp_seeded = lcp_out != nullptr;
    }
    c.sa_rounds = 1;
    c.sa_active_sum = n;
    c.sa_first_residue = m;

    const u32 rbits = bits_for_host(n - 1);
    u64 h = pp.k;
    while (m > 0) {
        // active list: slots pos[pcur][0..m), suffixes vals[res^1][0..m), group ids gid[0..m)
        u64* k2[2] = {keys[0], keys[1]};
        u32* v2[2] = {vals[res ^ 1], vals[res]};
        TDC_LAUNCH(build_keys_kernel, u32(div_up(m, 256)), 256, 0, st, v2[0], gid, c.d_isa, m, h, n, rbits, k2[0]);
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
