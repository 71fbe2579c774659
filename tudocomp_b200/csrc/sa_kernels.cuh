// Kernels of the prefix-doubling suffix-array builder (see suffix_array.cu for the algorithm); shared by the
// single-GPU driver (suffix_array.cu) and the sharded multi-GPU driver (dist_textds.cu).
#pragma once
#include <cmath>
#include <cstdlib>

#include "tdc_ctx.h"

namespace tdc {

// ---------------------------------------------------------------------------------------------------------------
// 0. byte histogram
// ---------------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) byte_histogram_kernel(const uint8_t* __restrict__ text, u64 n, u32* __restrict__ ghist) {
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
    u32 k;        // symbols per key (the last one may be cut by `drop`)
    u32 lenbits;  // width of the trailing length field (0: sentinel has its own code 0, no field needed)
    // packed records (single-GPU initial sort, suffix_array.cu): key = the b*k-bit symbol string without its lowest `drop`
    // bits, record = key << idx_bits | suffix.  idx_bits == 0: plain keys (values travel separately).
    u32 drop;
    u32 idx_bits;
    u32 known;    // whole symbols a key decides: (b*k - drop) / b
    u32 pow2;     // power-of-two alphabet: symbols are coded 0..s-1 and code 0 is shared with the sentinel / the padding
};

#ifndef PK_THREADS_CFG
#define PK_THREADS_CFG 128  // 3.55 ms at dna 2^30 against 4.01 with 256 (profiles/r2ae_variants_2p30.txt)
#endif
static const int PK_THREADS = PK_THREADS_CFG;
#ifndef PK_IPT_CFG
#define PK_IPT_CFG 8
#endif
static const int PK_IPT = PK_IPT_CFG;
static const int PK_TILE = PK_THREADS * PK_IPT;
static const int PK_HALO = 64;  // k <= 64 (b >= 1)

// HIST: the digits of every planned radix pass are counted here, while the key is in a register, into ghist[pass][256]
// (zeroed by the caller) — the sort then skips its own histogram kernel, i.e. one full read of the keys
// (radix_sort_pairs(..., hist_ready = true); 5.8 ms of the 2^30 B step).
template <bool HIST>
static __global__ void __launch_bounds__(PK_THREADS)
pack_keys_kernel(const uint8_t* __restrict__ text, u64 n, const uint8_t* __restrict__ code_map, PackParams pp,
                 u64* __restrict__ keys, u64 pos0, u64 cnt,  // keys[j] = key of suffix pos0 + j, j < cnt
                 PassPlan plan, u32* __restrict__ ghist) {
    __shared__ uint8_t codes[PK_TILE + PK_HALO];
    __shared__ uint8_t cmap[256];
    __shared__ u32 sh_hist[HIST ? RS_MAX_PASSES * RS_RADIX : 1];
    for (u32 i = threadIdx.x; i < 256; i += PK_THREADS) cmap[i] = code_map[i];
    if (HIST)
        for (u32 i = threadIdx.x; i < u32(RS_MAX_PASSES * RS_RADIX); i += PK_THREADS) sh_hist[i] = 0;
    __syncthreads();
    const u64 base = pos0 + u64(blockIdx.x) * PK_TILE;
    const u64 end = pos0 + cnt;
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
        if (pp.idx_bits) key = ((packed >> pp.drop) << pp.idx_bits) | p;  // packed record: the suffix rides in the low bits
        out[q] = key;
        packed = ((packed << pp.b) & mask) | codes[l0 + q + pp.k];  // roll one symbol
    }
    if (HIST) {
#pragma unroll
        for (int q = 0; q < PK_IPT; q++) rs_count_digits<u64>(sh_hist, plan, out[q], base + l0 + q < end);
    }
    if (base + l0 + PK_IPT <= end) {
        ulonglong2* o2 = reinterpret_cast<ulonglong2*>(keys + (base - pos0) + l0);
#pragma unroll
        for (int q = 0; q < PK_IPT / 2; q++) o2[q] = make_ulonglong2(out[2 * q], out[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < PK_IPT; q++) {
            const u64 p = base + l0 + q;
            if (p < end) keys[p - pos0] = out[q];
        }
    }
    if (HIST) {
        __syncthreads();
        for (u32 i = threadIdx.x; i < u32(plan.npass) * RS_RADIX; i += PK_THREADS)
            if (sh_hist[i]) atomicAdd(&ghist[i], sh_hist[i]);
    }
}

// number of equal leading symbols of two different keys = LCP of the two suffixes when it is < k
__device__ __forceinline__ u32 key_common_symbols(u64 a, u64 c, PackParams pp) {  // a, c: keys (records >> idx_bits)
    const u64 pa = a >> pp.lenbits, pc = c >> pp.lenbits;
    const u32 width = pp.b * pp.k - pp.drop;
    u32 common = pp.known;
    if (pa != pc) common = min(common, (u32(__clzll((long long)(pa ^ pc))) - (64u - width)) / pp.b);
    if (pp.lenbits) {
        const u32 lm = (1u << pp.lenbits) - 1u;
        common = min(common, min(u32(a) & lm, u32(c) & lm));  // a suffix ends where its sentinel stands
    }
    return common;
}

// ---------------------------------------------------------------------------------------------------------------
// 2. rerank (reduce -> scan of tile aggregates -> apply)
// ---------------------------------------------------------------------------------------------------------------
#ifndef RR_THREADS_CFG
#define RR_THREADS_CFG 512  // tiles of 2048 elements: with 256 threads the single-CTA scan over the tile aggregates doubles (r2z_variants.txt)
#endif
static const int RR_THREADS = RR_THREADS_CFG;
#ifndef RR_IPT_CFG
#define RR_IPT_CFG 4  // measured (profiles/r2y_variants.txt, dna 2^28): rerank_apply 1.79 ms with 4, 2.53 with 8, 2.49 with 16
#endif
static const int RR_IPT = RR_IPT_CFG;  // elements per thread (multiple of 4, < 32)
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
__device__ __forceinline__ void rr_flags(const K* kv, u64 m, u64 t0, u32* head_bits, u32* ns_bits, u32 idx_bits = 0) {
    u32 hb = 0;  // bit q (0..RR_IPT) = head(t0+q), with head(m) := 1 and head(0) := 1
#pragma unroll
    for (int q = 0; q <= RR_IPT; q++) {
        const u64 t = t0 + q;
        const bool h = (t == 0) || (t >= m) || ((kv[q] >> idx_bits) != (kv[q + 1] >> idx_bits));
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
static __global__ void __launch_bounds__(RR_THREADS)
rerank_reduce_kernel(const K* __restrict__ keys, u64 m, u32* __restrict__ agg_lasthead, ull* __restrict__ agg_cnt, u32 idx_bits) {
    __shared__ u32 s_max[RR_THREADS / 32];
    __shared__ ull s_sum[RR_THREADS / 32];
    const u64 t0 = u64(blockIdx.x) * RR_TILE + u64(threadIdx.x) * RR_IPT;
    u32 hb, nb;
    K kv[RR_IPT + 2];
    rr_load_keys<K>(keys, m, t0, kv);
    rr_flags<K>(kv, m, t0, &hb, &nb, idx_bits);
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
static __global__ void __launch_bounds__(1024)
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
static __global__ void __launch_bounds__(RR_THREADS)
rerank_apply_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ pos_in, u64 m,
                    const u32* __restrict__ pre_lasthead, const ull* __restrict__ pre_cnt, u32* __restrict__ sa,
                    u32* __restrict__ rank_idx, u32* __restrict__ rank_val, u32* __restrict__ pos_out,
                    u32* __restrict__ idx_out, u32* __restrict__ gid_out, u32* __restrict__ lcp_out, PackParams pp,
                    u32 slot_base,  // slot_base: global SA slot of this rank's first slot (0 on a single GPU)
                    u32 n_text) {   // packed records (pp.idx_bits != 0, FIRST only): text length; `vals` is unused
    __shared__ ull scratch_s[33];
    __shared__ u32 scratch_m[33];
    const u64 t0 = u64(blockIdx.x) * RR_TILE + u64(threadIdx.x) * RR_IPT;
    u32 hb, nb;
    K kv[RR_IPT + 2];
    const u32 ib = FIRST ? pp.idx_bits : 0u;
    rr_load_keys<K>(keys, m, t0, kv);
    rr_flags<K>(kv, m, t0, &hb, &nb, ib);
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
    if (FIRST && ib) {
        const u64 im = (u64(1) << ib) - 1;
#pragma unroll
        for (int q = 0; q < RR_IPT; q++) sfxv[q] = t0 + q < m ? u32(u64(kv[q + 1]) & im) : 0u;
    } else if (full) {
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
        const u32 slot = FIRST ? slot_base + u32(t) : pos_in[t];
        const u32 headslot = FIRST ? slot_base + hidx : pos_in[hidx];
        const u32 sfx = sfxv[q];
        headv[q] = headslot;  // rank[sfx] = headslot, applied by the partitioned scatter that follows
        if (FIRST) {
            if (sizeof(K) == 8 && lcp_out) {
                // (packed keys over a power-of-two alphabet have no length field: next to a suffix that runs into the sentinel
                // inside the key this value can be too large; lcp_tail_fix_kernel recomputes those few slots afterwards)
                lcpv[q] = t == 0 ? 0u : (h ? key_common_symbols(u64(kv[q]) >> ib, u64(kv[q + 1]) >> ib, pp) : LCP_UNKNOWN);
            }
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
            sa[slot - slot_base] = sfx;  // singleton group: final position
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
static __global__ void __launch_bounds__(256)
build_keys_kernel(const u32* __restrict__ idx, const u32* __restrict__ gid, const u32* __restrict__ rank, u64 m, u64 h,
                  u64 n, u32 rbits, u64* __restrict__ keys, u32 overshoot) {
    const u64 o = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o >= m) return;
    const u64 j = u64(idx[o]) + h;
    u64 r2;
    if (!overshoot) {
        r2 = j < n ? rank[j] : 0u;  // j < n always holds for active suffixes here; the guard is defensive
    } else {
        // Initial keys without a length field (packed records over a power-of-two alphabet): a suffix that ran into the
        // sentinel inside its known prefix is grouped with suffixes that continue with code-0 symbols.  It is the smaller
        // one, and of two such suffixes the shorter: rank them below every real rank, by how far they overshoot the end.
        r2 = j < n ? n + u64(rank[j]) : n - 1 - min(j - (n - 1), n - 1);
    }
    keys[o] = (u64(gid[o]) << rbits) | r2;
}

// Cost model for the number of symbols per initial key.  A radix pass moves 24 B per suffix; a suffix the initial sort
// leaves in a group costs about SA_ACTIVE_COST bytes of traffic in the doubling rounds (random rank gather, key
// build, up to 8 passes, re-rank, scatter).  For a memoryless source a k-symbol prefix is shared with another suffix
// with probability about n * 2^(-H0 k), H0 = order-0 entropy of the byte histogram.  A wrong guess (text with
// memory) only moves work between the two phases; the result is the same.
static const double SA_ACTIVE_COST = 400.0;
static const double SA_PACKED_ACTIVE_COST = 540.0;  // packed-vs-pairs decision, see choose_key_layout

static u32 bits_of_lenfield(u32 k) { return bits_for_host(k); }

// log2_collide: optional measured table, log2_collide[k] = log2 of the probability that two random suffixes share their
// first k bytes (k = 1..8, from sample_prefix_collisions; > 0 means "not measured").  It replaces the memoryless
// estimate 2^(-H0 k), which is far too optimistic for text with context (order-3 Markov English: H0 = 4.1 bit/symbol but
// ~2.3 bit/symbol of discrimination), and is extrapolated geometrically beyond k = 8.
// idx_bits_avail > 0 (single-GPU builder): the initial sort may also run on PACKED records, key << idx_bits | suffix in one
// 64-bit word, which moves 16 B per suffix and pass instead of 24 but leaves only 64 - idx_bits key bits (a shorter known
// prefix, hence more suffixes for the doubling rounds).  The model prices both layouts and takes the cheaper one.
static void choose_key_layout(const u32* hist, u64 n, PackParams* pp, u32* sigbits, const double* log2_collide = nullptr,
                              u32 idx_bits_avail = 0) {
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
    // expected share of suffixes that a prefix of kf symbols (fractional: a cut symbol counts for its share of bits)
    // leaves in groups, and the doubling rounds a grouped suffix then still has to go through
    int k2 = 0, k1 = 0;
    double rate = 0;
    if (log2_collide) {
        // last two measured points with enough collisions give the decay rate for the extrapolation
        for (int q = 8; q >= 2; q--)
            if (log2_collide[q] <= 0.0 && log2_collide[q - 1] <= 0.0) { k2 = q; break; }
        if (k2) {
            k1 = k2 >= 3 && log2_collide[k2 - 2] <= 0.0 ? k2 - 2 : k2 - 1;
            rate = (log2_collide[k1] - log2_collide[k2]) / double(k2 - k1);  // bits per symbol
        }
    }
    auto residue_of = [&](double kf, double* rounds_left) {
        double residue = h0 > 1e-9 ? exp2(log2(double(n)) - h0 * kf) : 1.0;
        *rounds_left = 0.0;
        if (k2) {
            double lp;
            if (kf <= double(k2)) {
                const int lo = kf < 1.0 ? 1 : int(kf);
                const int hi = lo < 8 && log2_collide[lo + 1] <= 0.0 ? lo + 1 : lo;
                const double f = kf < 1.0 ? 0.0 : kf - double(lo);
                lp = log2_collide[lo] + (log2_collide[hi] - log2_collide[lo]) * f;
            } else {
                lp = log2_collide[k2] - rate * (kf - double(k2));
            }
            const double expected_twins = exp2(log2(double(n)) + lp);
            residue = 1.0 - exp(-expected_twins);
            // a text that keeps most suffixes in groups whatever k is (repeats) pays one doubling round per factor of two
            // that the initial key is shorter than the longest possible one
            *rounds_left = kf < double(kmax) ? 0.5 * log2(double(kmax) / kf) : 0.0;
        }
        return residue > 1.0 ? 1.0 : residue;
    };
    u32 best = kmax;
    double best_cost = 1e300;
    const char* mode_env = getenv("TDCGPU_SA_MODE");  // tuning/debug: "wide" | "packed"
    const bool force_wide = mode_env && mode_env[0] == 'w', force_packed = mode_env && mode_env[0] == 'p';
    if (const char* e = getenv("TDCGPU_SA_SYMBOLS")) {  // tuning/debug override (wide layout)
        const long v = atol(e);
        if (v >= 1 && v <= long(kmax)) best = u32(v);
        best_cost = -1.0;
    } else {
        for (u32 k = 1; k <= kmax; k++) {
            const u32 bits = b * k + (pow2 ? bits_of_lenfield(k) : 0);
            const double passes = double((bits + 7) / 8);
            double rounds_left;
            const double residue = residue_of(double(k), &rounds_left);
            const double cost = passes * 24.0 + residue * SA_ACTIVE_COST * (1.0 + rounds_left);
            if (cost <= best_cost) { best_cost = cost; best = k; }  // ties: more symbols
        }
    }
    pp->b = b;
    pp->k = best;
    pp->lenbits = pow2 ? bits_of_lenfield(best) : 0;
    pp->drop = 0;
    pp->idx_bits = 0;
    pp->known = best;
    pp->pow2 = pow2 ? 1u : 0u;
    *sigbits = b * best + pp->lenbits;
    if (!idx_bits_avail || idx_bits_avail >= 64 - b || force_wide || (best_cost < 0 && !force_packed)) return;
    // packed records: kb key bits, no length field (suffixes that run into the sentinel are handled by the doubling keys)
    const u32 kb_max = min(64u - idx_bits_avail, b * u32(PK_HALO));
    u32 best_kb = 0;
    double best_pcost = 1e300;
    if (const char* e = getenv("TDCGPU_SA_KEYBITS")) {
        const long v = atol(e);
        if (v >= long(b) && v <= long(kb_max)) { best_kb = u32(v); best_pcost = -1.0; }
    }
    if (!best_kb) {
        for (u32 kb = b; kb <= kb_max; kb++) {
            const double passes = double((kb + 7) / 8);
            double rounds_left;
            const double residue = residue_of(double(kb) / double(b), &rounds_left);
            // measured (profiles/r2h_sortbench.txt, 2^28 elements): a keys-only pass takes 0.79 of a pair pass (1.544 vs 1.956 ms)
            // although it moves 0.67 of the bytes (both are bound by the shared-memory pipe, not by DRAM), and a suffix left
            // to the doubling rounds costs ~540 B of pair-pass time (round-2 sort, rank gather, re-rank, rank scatter,
            // direct LCP of its slot)
            const double cost = passes * 19.0 - 8.0 + residue * SA_PACKED_ACTIVE_COST * (1.0 + rounds_left);
            if (cost <= best_pcost) { best_pcost = cost; best_kb = kb; }
        }
    }
    if (!best_kb || (!force_packed && best_pcost >= best_cost)) return;
    pp->k = (best_kb + b - 1) / b;
    pp->lenbits = 0;
    pp->drop = b * pp->k - best_kb;
    pp->idx_bits = idx_bits_avail;
    pp->known = best_kb / b;
    *sigbits = best_kb;
}


// ---------------------------------------------------------------------------------------------------------------
// collision statistics of a regular sample of suffix prefixes (input of the key-length cost model)
// ---------------------------------------------------------------------------------------------------------------
static __global__ void sample_prefix_kernel(const uint8_t* __restrict__ text, u64 n, u32 samples, u64* __restrict__ out) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= samples) return;
    const u64 p = (u64(j) * n) / samples;
    u64 key = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) key = (key << 8) | (p + q < n ? text[p + q] : 0);  // big-endian: integer order = byte order
    out[j] = key;
}

// sorted[0..S): sorted 8-byte prefixes.  log2_collide[k] (k = 1..8) = log2 P(two samples share k bytes), +1 if fewer
// than 32 colliding pairs were seen (unreliable).
static void prefix_collision_table(const u64* sorted, u32 S, double log2_collide[9]) {
    double pairs[9] = {0};
    u64 run[9];
    for (int k = 1; k <= 8; k++) run[k] = 1;
    for (u32 i = 1; i <= S; i++) {
        int common = 0;
        if (i < S) {
            const u64 x = sorted[i] ^ sorted[i - 1];
            common = x ? int(__builtin_clzll(x) >> 3) : 8;
        }
        for (int k = 1; k <= 8; k++) {
            if (k <= common) {
                run[k]++;
            } else {
                pairs[k] += double(run[k]) * double(run[k] - 1) * 0.5;
                run[k] = 1;
            }
        }
    }
    const double all = double(S) * double(S - 1) * 0.5;
    log2_collide[0] = 0.0;
    for (int k = 1; k <= 8; k++) log2_collide[k] = pairs[k] >= 32.0 ? log2(pairs[k] / all) : 1.0;
}

}  // namespace tdc
