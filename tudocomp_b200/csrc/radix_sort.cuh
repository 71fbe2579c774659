// Hand-written LSD radix sort of (key, u32 value) pairs for sm_100a — the sorting engine of the prefix-doubling
// suffix-array builder (replaces sssort/trsort of the reference's divsufsort,
// /root/reference/include/tudocomp/util/divsufsort/divsufsort_ssort.hpp:691, divsufsort_trsort.hpp:506).
//
// Structure ("onesweep"): ONE histogram kernel counts every digit position of the key range in a single read; each
// digit pass is then ONE kernel that reads a tile, ranks it with warp-level match/ballot, stages the tile in shared
// memory in digit order and writes each digit's run coalesced.  The tile's global offsets come from a decoupled
// look-back over per-tile digit counts (one 64-bit descriptor per (tile, digit) carrying epoch|status|count in a single
// word, so a reader never sees a count without its flag).  Per pass: keys and values are read once and written once.
#pragma once
#include "tdc_common.cuh"

namespace tdc {

#ifndef RS_THREADS_CFG
#define RS_THREADS_CFG 256
#endif
#ifndef RS_IPT64_CFG
#define RS_IPT64_CFG 16
#endif
#ifndef RS_LOOKBACK_W
#define RS_LOOKBACK_W 4
#endif
#ifndef RS_MIN_CTAS_CFG
#define RS_MIN_CTAS_CFG 3
#endif
static const int RS_THREADS = RS_THREADS_CFG;  // >= 256: one thread per digit in the look-back
static const int RS_MIN_CTAS = RS_MIN_CTAS_CFG;
static const int RS_WARPS = RS_THREADS / 32;
static const int RS_RADIX = 256;
#ifndef RS_HIST_GRID_MULT_CFG
#define RS_HIST_GRID_MULT_CFG 16  // (dna 2^30: 5.95 ms with 2, 5.75 with 4, 5.12 with 8, 4.97 with 16, 4.89 with 32; profiles/r2ad/r2ae_variants_2p30.txt)
#endif
static const int RS_HIST_GRID_MULT = RS_HIST_GRID_MULT_CFG;  // histogram CTAs (of 512 threads) per SM
#ifndef RS_HIST_EPT_CFG
#define RS_HIST_EPT_CFG 4
#endif
static const int RS_HIST_EPT = RS_HIST_EPT_CFG;
static const int RS_MAX_PASSES = 8;

struct PassPlan {
    int npass;
    u32 shift[RS_MAX_PASSES];
    u32 mask[RS_MAX_PASSES];
};

// digits of the key bits [begin_bit, end_bit): as few passes as 8-bit digits allow, widths balanced
static inline PassPlan make_pass_plan(int begin_bit, int end_bit) {
    PassPlan plan;
    const int bits = end_bit - begin_bit;
    plan.npass = bits <= 0 ? 0 : (bits + 7) / 8;
    const int np = plan.npass <= RS_MAX_PASSES ? plan.npass : RS_MAX_PASSES;
    const int base = np ? bits / plan.npass : 0, extra = np ? bits % plan.npass : 0;
    int sh = begin_bit;
    for (int p = 0; p < np; p++) {
        const int wd = base + (p < extra ? 1 : 0);
        plan.shift[p] = u32(sh);
        plan.mask[p] = (1u << wd) - 1u;
        sh += wd;
    }
    return plan;
}

// One key's digits of every planned pass into the CTA's shared-memory histograms sh[pass][256]; must be called by all 32
// lanes of a warp (`valid` = this lane has a key).  A warp-uniform digit (constant high bits, runs) costs one shared
// atomic instead of a 32-way conflict.
template <class K>
__device__ __forceinline__ void rs_count_digits(u32* sh, const PassPlan& plan, K k, bool valid) {
#pragma unroll
    for (int p = 0; p < RS_MAX_PASSES; p++) {
        if (p < plan.npass) {
            const u32 d = u32(k >> plan.shift[p]) & plan.mask[p];
            const u32 d0 = __shfl_sync(kFull, d, 0);
            if (__all_sync(kFull, valid && d == d0)) {
                if (lane_id() == 0) atomicAdd(&sh[p * RS_RADIX + d0], 32u);
            } else if (valid) {
                atomicAdd(&sh[p * RS_RADIX + d], 1u);
            }
        }
    }
}

struct SortWorkspace {
    u32* hist = nullptr;          // device [RS_MAX_PASSES][256]: digit counts, then exclusive bucket starts
    u32* uniform = nullptr;       // device [RS_MAX_PASSES]: 1 if one bin holds every key (pass can be skipped)
    ull* desc = nullptr;          // device [max_tiles][256] look-back descriptors
    u32* h_uniform = nullptr;     // pinned host mirror of `uniform`
    u64 max_tiles = 0;
    u32 epoch = 0;                // bumped once per executed pass; stale descriptors never match
    int sm_count = 148;
    // statistics for the work model: executed passes and elements moved
    u64 stat_passes = 0, stat_elems = 0;
};

template <class K> struct RsCfg;
template <> struct RsCfg<u64> { static const int IPT = RS_IPT64_CFG; static const int MIN_CTAS = RS_MIN_CTAS_CFG; };
#ifndef RS_IPT32_CFG
#define RS_IPT32_CFG 24  // u32 pairs (scatter partitions): 24 x 256 @ 3 CTAs is 8 % faster than 16 @ 4 (profiles/r2aa_variants.txt)
#endif
#ifndef RS_MIN_CTAS32_CFG
#define RS_MIN_CTAS32_CFG 3
#endif
template <> struct RsCfg<u32> { static const int IPT = RS_IPT32_CFG; static const int MIN_CTAS = RS_MIN_CTAS32_CFG; };
#ifndef RS_MIN_CTAS_KEYS_CFG
#define RS_MIN_CTAS_KEYS_CFG 3  // 24 keys x 256 threads @ 3 CTAs: 4.3 % faster than 16 @ 4 (profiles/r2ab_variants.txt)
#endif
// KEYSONLY passes (packed records: the value lives in the low bits of the key) move 16 B per element instead of 24 and
// need neither the value registers nor the value half of the staging buffer
#ifndef RS_IPT_KEYS_CFG
#define RS_IPT_KEYS_CFG 24
#endif
template <class K, bool KEYSONLY> struct RsOcc { static const int MIN_CTAS = RsCfg<K>::MIN_CTAS; static const int IPT = RsCfg<K>::IPT; };
template <> struct RsOcc<u64, true> { static const int MIN_CTAS = RS_MIN_CTAS_KEYS_CFG; static const int IPT = RS_IPT_KEYS_CFG; };

static const ull RS_STATUS_AGG = 1, RS_STATUS_PREFIX = 2;

__device__ __forceinline__ void desc_store(ull* p, ull v) {
#ifdef TDC_CUSIM
    *p = v;
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ ull desc_load(const ull* p) {
#ifdef TDC_CUSIM
    return *p;
#else
    ull v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// all digit histograms in one read of the keys
// ---------------------------------------------------------------------------------------------------------------
template <class K>
__global__ void __launch_bounds__(512) rs_histogram_kernel(const K* __restrict__ keys, u64 m, PassPlan plan,
                                                           u32* __restrict__ ghist) {
    __shared__ u32 sh[RS_MAX_PASSES * RS_RADIX];
    for (u32 i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    // RS_HIST_EPT keys per thread and iteration, all loads issued before the first counter update (one key per iteration
    // left the kernel waiting on a single 8-byte load per thread: 1.5 TB/s)
    constexpr int EPT = RS_HIST_EPT;
    const u64 chunk = u64(32) * EPT;                       // keys per warp and iteration, lane-strided inside the chunk
    const u64 m_round = (m + chunk - 1) / chunk * chunk;
    const u64 warps = (u64(gridDim.x) * blockDim.x) >> 5;
    for (u64 base = ((u64(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * chunk; base < m_round; base += warps * chunk) {
        K kk[EPT];
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const u64 i = base + u64(e) * 32 + lane_id();
            kk[e] = i < m ? keys[i] : K(0);
        }
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            rs_count_digits<K>(sh, plan, kk[e], base + u64(e) * 32 + lane_id() < m);
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < u32(plan.npass) * RS_RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// one CTA of 256 threads per pass: counts -> exclusive bucket starts; flag single-bin passes
static __global__ void __launch_bounds__(256) rs_scan_kernel(u32* __restrict__ ghist, u32* __restrict__ uniform, u64 m) {
    __shared__ u32 scratch[33];
    u32* h = ghist + blockIdx.x * RS_RADIX;
    const u32 c = h[threadIdx.x];
    u32 total;
    const u32 ex = block_exclusive_sum<u32>(c, scratch, &total);
    h[threadIdx.x] = ex;
    if (u64(c) == m) uniform[blockIdx.x] = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// one digit pass
// ---------------------------------------------------------------------------------------------------------------
// peers of this lane = lanes holding the same digit.  Eight ballots instead of match.any: MATCH is a slow-path
// instruction on sm_100 (ncu: 43 % of the kernel's stall samples sat on its result), VOTE is full rate.
// All eight bits are always voted on: bits above a narrow digit are 0 in every lane and leave `peers` unchanged.  The
// bit is tested ONCE into a predicate that feeds both the vote and the select — written as `(d >> b) & 1` in two places
// the compiler emitted 8 instructions per bit (shift, and, compare for the vote; test, select, combine for the mask,
// plus a uniform branch on the digit width): 64 of the kernel's ~143 instructions per key.  Now 4 per bit.
__device__ __forceinline__ u32 match_digit(u32 d) {
    u32 peers = kFull;
#pragma unroll
    for (int b = 0; b < 8; b++) {
#ifdef TDC_CUSIM
        const bool bit = (d & (1u << b)) != 0u;
        const u32 bal = __ballot_sync(kFull, bit);
        peers &= bit ? bal : ~bal;
#else
        // test -> predicate, vote, select, combine: the C++ form of this compiled to 6-8 instructions per bit
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 t, bal, m;\n\t"
            "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
            "selp.b32 m, 0, 0xffffffff, p;\n\t"
            "xor.b32 bal, bal, m;\n\tand.b32 %0, %0, bal;\n\t}"
            : "+r"(peers) : "r"(d), "r"(1u << b));
#endif
    }
    return peers;
}

// Tile ids are blockIdx.x: CTAs of a 1-D grid are dispatched in index order, so every predecessor a tile looks back
// at is resident or finished (the same assumption CUB's decoupled look-back scan makes).
template <class K, bool IOTA, bool KEYSONLY = false>
__global__ void __launch_bounds__(RS_THREADS, (RsOcc<K, KEYSONLY>::MIN_CTAS))
rs_onesweep_kernel(const K* __restrict__ kin, K* __restrict__ kout, const u32* __restrict__ vin, u32* __restrict__ vout,
                   u64 m, u32 shift, u32 mask, const u32* __restrict__ bucket_start, ull* __restrict__ desc, u32 epoch) {
    constexpr int IPT = RsOcc<K, KEYSONLY>::IPT;
    constexpr int TILE = RS_THREADS * IPT;
    static_assert(TILE <= 65536, "tile-local ranks are kept in 16 bits");
    TDC_DYN_SMEM(smem_raw);
    K* skeys = reinterpret_cast<K*>(smem_raw);                                  // TILE keys
    u32* svals = reinterpret_cast<u32*>(smem_raw + sizeof(K) * TILE);           // TILE values (absent when KEYSONLY)
    u32* warp_cnt = svals + (KEYSONLY ? 0 : TILE);                              // [RS_WARPS][256]
    u32* gbase = warp_cnt + RS_WARPS * RS_RADIX;                                // [256] (+ 256 spare words)
    u32* misc = gbase + RS_RADIX;                                               // [40]: scan scratch

    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const u32 tile = blockIdx.x;
    const u64 tile_base = u64(tile) * TILE;
    const u32 count = u32(min(u64(TILE), m - tile_base));

    // ---- load (warp-striped: coalesced, and index order == (k, lane) order inside a warp) ----
    K key[IPT];
    u32 rank2[(IPT + 1) / 2];  // tile-local ranks (< TILE <= 65536), two per register: 8 registers less across the look-back
#pragma unroll
    for (int k = 0; k < (IPT + 1) / 2; k++) rank2[k] = 0;
    // tile-relative 32-bit indices; a full tile (all but the last) skips the per-element bounds checks, which were 6 of
    // the 7 instructions per key of this load (64-bit compare + select)
    const bool full = count == u32(TILE);
    const K* __restrict__ kin_t = kin + tile_base;
    const u32 wloc = w * (32 * IPT) + lane;  // this thread's first element inside the tile
    if (full) {
#pragma unroll
        for (int k = 0; k < IPT; k++) key[k] = kin_t[wloc + u32(k) * 32];
    } else {
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const u32 loc = wloc + u32(k) * 32;
            key[k] = loc < count ? kin_t[loc] : ~K(0);
        }
    }
    u32* my_cnt = warp_cnt + w * RS_RADIX;
#pragma unroll
    for (int j = 0; j < RS_RADIX / 32; j++) my_cnt[j * 32 + lane] = 0;
    __syncwarp();
    // ---- stable rank inside the warp ----
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const u32 d = u32(key[k] >> shift) & mask;
        const u32 peers = match_digit(d);
        const u32 before = __popc(peers & lanemask_lt());
        const u32 c = my_cnt[d];
        __syncwarp();
        if (before == 0) my_cnt[d] = c + __popc(peers);
        __syncwarp();
        rank2[k >> 1] |= (c + before) << (16 * (k & 1));
    }
    __syncthreads();

    // ---- per-digit totals, tile-local digit starts, warp offsets ----
    u32 total = 0, pub = 0;
    if (tid < RS_RADIX) {
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) total += warp_cnt[ww * RS_RADIX + tid];
        // publish the tile's count of this digit as early as possible (successors are waiting on it)
        pub = total;
        if (tid == mask) pub -= (u32(TILE) - count);  // padding keys (~0) sit in the top bin and are never written
        if (tile != 0) desc_store(desc + u64(tile) * RS_RADIX + tid, (ull(epoch) << 34) | (RS_STATUS_AGG << 32) | pub);
    }
    u32 blk_total;
    const u32 ex = block_exclusive_sum<u32>(total, misc, &blk_total);  // threads >= 256 contribute 0
    if (tid < RS_RADIX) {
        // warp_cnt[w][d] := tile-local position of warp w's first key with digit d (digit start + keys of earlier warps):
        // the staging loop then needs ONE random shared-memory read per key instead of two (ncu: 16 of the kernel's 28
        // shared-memory wavefronts per 32 keys were in that loop, l1tex was the busiest unit at 66 %)
        u32 run = ex;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            const u32 t = warp_cnt[ww * RS_RADIX + tid];
            warp_cnt[ww * RS_RADIX + tid] = run;
            run += t;
        }
    }
    // ---- decoupled look-back for digit `tid` ----
    auto lookback = [&]() {
        if (tid >= RS_RADIX) return;
        const ull tag = ull(epoch) << 34;
        ull* my_desc = desc + u64(tile) * RS_RADIX + tid;
        u32 exclusive = 0;
        if (tile != 0) {
            // A window of RS_LOOKBACK_W predecessors is fetched with independent loads and then consumed in order: the
            // serial chain "load, test, step back" was 40 % of the kernel's stall samples (~16 dependent L2 round trips
            // per tile, profiles/r1g_ncu_summary.md).  Entries behind the first PREFIX are ignored; a window that meets an
            // unpublished predecessor is re-polled from there.  Indices below tile 0 are clamped to tile 0, whose PREFIX
            // always ends the walk.
            u32 t = tile - 1;  // nearest predecessor not yet consumed
            bool done = false;
            while (!done) {
                ull v[RS_LOOKBACK_W];
#pragma unroll
                for (int w = 0; w < RS_LOOKBACK_W; w++) {
                    const u32 tw = t >= u32(w) ? t - u32(w) : 0u;
                    v[w] = desc_load(desc + u64(tw) * RS_RADIX + tid);
                }
#pragma unroll
                for (int w = 0; w < RS_LOOKBACK_W; w++) {
                    if (done) break;
                    const u32 status = u32(v[w] >> 32) & 3u;
                    if ((v[w] >> 34) != ull(epoch) || status == 0) break;  // not published yet: poll again from here
                    exclusive += u32(v[w]);
                    if (status == RS_STATUS_PREFIX) done = true; else t--;
                }
            }
        }
        desc_store(my_desc, tag | (RS_STATUS_PREFIX << 32) | ull(exclusive + pub));
        gbase[tid] = bucket_start[tid] + exclusive - ex;
    };
#ifdef RS_EARLY_LOOKBACK
    lookback();  // (the round-1 order: A/B only)
#endif
    __syncthreads();

    // ---- stage in shared memory in digit order ----
    const u32* __restrict__ vin_t = (IOTA || KEYSONLY) ? nullptr : vin + tile_base;
    u32 val[KEYSONLY ? 1 : IPT];
    if (!KEYSONLY) {  // all value loads of the thread are issued before the first dependent shared-memory store
                      // (issuing them before the per-digit section instead measured the same: 1.947 vs 1.956 ms per pass)
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const u32 loc = wloc + u32(k) * 32;
            val[k] = IOTA ? u32(tile_base) + loc : ((full || loc < count) ? vin_t[loc] : 0u);
        }
    }
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const u32 d = u32(key[k] >> shift) & mask;
        const u32 p = warp_cnt[w * RS_RADIX + d] + ((rank2[k >> 1] >> (16 * (k & 1))) & 0xffffu);
        skeys[p] = key[k];  // (padding keys of the partial last tile land behind `count` in the top bin and are never written out)
        if (!KEYSONLY) svals[p] = val[k];
    }
#ifndef RS_EARLY_LOOKBACK
    // The look-back runs AFTER the staging: the predecessors have had the time of this tile's value loads and
    // shared-memory scatter to publish, so far fewer polls (look-back first: 28 % of the kernel's stall samples and 23 % of
    // its executed instructions were descriptor polling, profiles/r2e_ncu_sortpass_summary.md; 2.09 -> 1.96 ms per pass at
    // 2^28 pairs).  The global offsets are only needed by the write-out below.
    lookback();
#endif
    __syncthreads();

    // ---- coalesced runs out ----
    if (full) {
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const u32 j = tid + u32(k) * RS_THREADS;
            const K kk = skeys[j];
            const u32 o = gbase[u32(kk >> shift) & mask] + j;
            kout[o] = kk;
            if (!KEYSONLY) vout[o] = svals[j];
        }
    } else {
        for (u32 j = tid; j < count; j += RS_THREADS) {
            const K kk = skeys[j];
            const u32 o = gbase[u32(kk >> shift) & mask] + j;
            kout[o] = kk;
            if (!KEYSONLY) vout[o] = svals[j];
        }
    }
}

// bucket starts of a single pass over keys that are a permutation of 0..m-1: no need to read the keys
static __global__ void __launch_bounds__(256) rs_perm_starts_kernel(u32* __restrict__ ghist, u64 m, u32 shift) {
    const u64 s = u64(threadIdx.x) << shift;
    ghist[threadIdx.x] = u32(s < m ? s : m);
}

static __global__ void rs_iota_kernel(u32* v, u64 m) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < m) v[i] = u32(i);
}

template <class K>
static inline size_t rs_smem_bytes(bool keysonly = false) {
    const int TILE = RS_THREADS * (keysonly ? RsOcc<u64, true>::IPT : RsCfg<K>::IPT);
    size_t bytes = sizeof(K) * TILE + (keysonly ? 0 : 4 * TILE) + 4 * (RS_WARPS * RS_RADIX + 2 * RS_RADIX + 40);
    return bytes;
}
template <class K>
static inline u64 rs_tiles(u64 m) {
    return div_up(m, u64(RS_THREADS) * RsCfg<K>::IPT);
}
static inline u64 rs_tiles_keys(u64 m) { return div_up(m, u64(RS_THREADS) * RsOcc<u64, true>::IPT); }

int sort_workspace_init(SortWorkspace& ws, u64 max_elems, int sm_count);
void sort_workspace_free(SortWorkspace& ws);

// Sorts m pairs by key bits [begin_bit, end_bit).  Buffers ping-pong between (k[0], v[0]) and (k[1], v[1]); the input
// is in slot 0 and *result receives the slot holding the output.  With iota=true the input values are implicitly
// 0..m-1 (v[0] is not read).  keys_are_perm: the keys are a permutation of 0..m-1 and [begin_bit, end_bit) is one
// digit reaching the top key bit, so the bucket starts are known without a histogram.  Stable.
template <class K>
static int radix_sort_pairs(SortWorkspace& ws, cudaStream_t st, K* k[2], u32* v[2], u64 m, int begin_bit, int end_bit,
                            bool iota, int* result, bool keys_are_perm = false, bool hist_ready = false) {
    // hist_ready: the producer of the keys has already counted the digits of make_pass_plan(begin_bit, end_bit) into
    // ws.hist (rs_count_digits; pack_keys_kernel does, which saves one full read of the keys)
    *result = 0;
    if (m == 0) return 0;
    {
        // > 48 KB of dynamic shared memory needs an opt-in per kernel, per device and per translation unit (internal
        // linkage: every TU that sorts has its own copy of this function and of the kernels it launches); a few us
        auto k1 = rs_onesweep_kernel<K, true>;
        auto k2 = rs_onesweep_kernel<K, false>;
        TDC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<K>())));
        TDC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rs_smem_bytes<K>())));
    }
    const int bits = end_bit - begin_bit;
    const PassPlan plan = make_pass_plan(begin_bit, end_bit);
    if (plan.npass > RS_MAX_PASSES) { set_error("radix_sort_pairs: %d bits need more than %d passes", bits, RS_MAX_PASSES); return -1; }
    if (rs_tiles<K>(m) > ws.max_tiles) { set_error("radix_sort_pairs: workspace too small"); return -1; }
    int cur = 0;
    bool need_iota = iota;
    if (plan.npass == 1 && keys_are_perm) {
        TDC_LAUNCH(rs_perm_starts_kernel, 1, 256, 0, st, ws.hist, m, plan.shift[0]);
        ws.h_uniform[0] = 0;
    } else if (plan.npass > 0) {
        TDC_CUDA(cudaMemsetAsync(ws.uniform, 0, sizeof(u32) * RS_MAX_PASSES, st));
        if (!hist_ready) {
            TDC_CUDA(cudaMemsetAsync(ws.hist, 0, sizeof(u32) * RS_MAX_PASSES * RS_RADIX, st));
            const u32 hgrid = u32(min(u64(ws.sm_count) * RS_HIST_GRID_MULT, div_up(m, 512 * 8)));
            auto rs_histogram = rs_histogram_kernel<K>;
            TDC_LAUNCH(rs_histogram, hgrid, 512, 0, st, k[0], m, plan, ws.hist);
            prof_add_bytes("rs_histogram", double(m) * sizeof(K));
        }
        TDC_LAUNCH(rs_scan_kernel, plan.npass, 256, 0, st, ws.hist, ws.uniform, m);
        TDC_KCHECK();
        TDC_CUDA(cudaMemcpyAsync(ws.h_uniform, ws.uniform, sizeof(u32) * RS_MAX_PASSES, cudaMemcpyDeviceToHost, st));
        TDC_CUDA(cudaStreamSynchronize(st));
    }
    if (plan.npass > 0) {
        const u32 grid = u32(rs_tiles<K>(m));
        const size_t smem = rs_smem_bytes<K>();
        for (int p = 0; p < plan.npass; p++) {
            if (ws.h_uniform[p]) continue;  // every key has the same digit here: the pass would be the identity
            ws.epoch++;
            // algorithmic bytes of one pass: keys and values read once, written once
            if (need_iota) {
                auto rs_onesweep_iota = rs_onesweep_kernel<K, true>;
                TDC_LAUNCH(rs_onesweep_iota, grid, RS_THREADS, smem, st, k[cur], k[cur ^ 1], v[cur], v[cur ^ 1], m,
                           plan.shift[p], plan.mask[p], ws.hist + p * RS_RADIX, ws.desc, ws.epoch);
                prof_add_bytes("rs_onesweep_iota", double(m) * (2 * sizeof(K) + 4));
            } else {
                auto rs_onesweep = rs_onesweep_kernel<K, false>;
                TDC_LAUNCH(rs_onesweep, grid, RS_THREADS, smem, st, k[cur], k[cur ^ 1], v[cur], v[cur ^ 1], m,
                           plan.shift[p], plan.mask[p], ws.hist + p * RS_RADIX, ws.desc, ws.epoch);
                prof_add_bytes("rs_onesweep", double(m) * (2 * sizeof(K) + 8));
            }
            TDC_KCHECK();
            need_iota = false;
            cur ^= 1;
            ws.stat_passes++;
            ws.stat_elems += m;
        }
    }
    if (need_iota) {
        TDC_LAUNCH(rs_iota_kernel, u32(div_up(m, 256)), 256, 0, st, v[cur], m);
        TDC_KCHECK();
    }
    *result = cur;
    return 0;
}

// Sorts m 64-bit RECORDS by their bits [begin_bit, end_bit); the bits below begin_bit are payload (the suffix index of
// the packed initial sort, suffix_array.cu) and travel inside the record: 16 B per element and pass instead of 24.
// Stable.  Input in k[0]; *result = slot holding the output.
static int radix_sort_keys(SortWorkspace& ws, cudaStream_t st, u64* k[2], u64 m, int begin_bit, int end_bit, int* result) {
    *result = 0;
    if (m == 0) return 0;
    auto kern = rs_onesweep_kernel<u64, false, true>;
    const size_t smem = rs_smem_bytes<u64>(true);
    TDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int bits = end_bit - begin_bit;
    PassPlan plan;
    plan.npass = bits <= 0 ? 0 : (bits + 7) / 8;
    if (plan.npass > RS_MAX_PASSES) { set_error("radix_sort_keys: %d bits need more than %d passes", bits, RS_MAX_PASSES); return -1; }
    if (rs_tiles_keys(m) > ws.max_tiles) { set_error("radix_sort_keys: workspace too small"); return -1; }
    if (plan.npass == 0) return 0;
    {
        int base = bits / plan.npass, extra = bits % plan.npass, sh = begin_bit;
        for (int p = 0; p < plan.npass; p++) {
            int wd = base + (p < extra ? 1 : 0);
            plan.shift[p] = u32(sh);
            plan.mask[p] = (1u << wd) - 1u;
            sh += wd;
        }
    }
    TDC_CUDA(cudaMemsetAsync(ws.hist, 0, sizeof(u32) * RS_MAX_PASSES * RS_RADIX, st));
    TDC_CUDA(cudaMemsetAsync(ws.uniform, 0, sizeof(u32) * RS_MAX_PASSES, st));
    const u32 hgrid = u32(min(u64(ws.sm_count) * RS_HIST_GRID_MULT, div_up(m, 512 * 8)));
    auto rs_histogram = rs_histogram_kernel<u64>;
    TDC_LAUNCH(rs_histogram, hgrid, 512, 0, st, k[0], m, plan, ws.hist);
    prof_add_bytes("rs_histogram", double(m) * 8);
    TDC_LAUNCH(rs_scan_kernel, plan.npass, 256, 0, st, ws.hist, ws.uniform, m);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(ws.h_uniform, ws.uniform, sizeof(u32) * RS_MAX_PASSES, cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    int cur = 0;
    const u32 grid = u32(rs_tiles_keys(m));
    for (int p = 0; p < plan.npass; p++) {
        if (ws.h_uniform[p]) continue;
        ws.epoch++;
        auto rs_onesweep_keys = kern;
        TDC_LAUNCH(rs_onesweep_keys, grid, RS_THREADS, smem, st, k[cur], k[cur ^ 1], (const u32*)nullptr, (u32*)nullptr, m, plan.shift[p],
                   plan.mask[p], ws.hist + p * RS_RADIX, ws.desc, ws.epoch);
        prof_add_bytes("rs_onesweep_keys", double(m) * 16);
        TDC_KCHECK();
        cur ^= 1;
        ws.stat_passes++;
        ws.stat_elems += m;
    }
    *result = cur;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// partitioned scatter: dst[idx[t]] = val[t] for distinct idx
// ---------------------------------------------------------------------------------------------------------------
// A 4-byte store to a random address costs DRAM a 32 B sector read plus a 32 B sector write (ncu: 64-80 B/element for
// the ISA / Phi / LPF scatters).  Instead: one radix pass partitions the (idx, val) pairs by the top bits of idx into
// <= 256 destination windows (each far smaller than the 126 MB L2), then the pairs are scattered window by window:
// every store of a window lands in L2, sectors are completed there and written back once.
// One element per thread, CTAs in index order: the in-order CTA dispatch keeps all resident CTAs inside the same window
// (measured with tools/scatterbench.cu: 160 Gelem/s for 16 MiB windows vs 23 Gelem/s unpartitioned; a grid-stride
// loop loses the locality and most of the gain).
#ifndef SP_EPT_CFG
#define SP_EPT_CFG 4
#endif
static const int SP_EPT = SP_EPT_CFG;  // pairs per thread (two 16-byte loads); a CTA still covers one contiguous run of the pairs
#ifndef SP_THREADS_CFG
#define SP_THREADS_CFG 256
#endif
static const int SP_THREADS = SP_THREADS_CFG;
static __global__ void __launch_bounds__(SP_THREADS)
scatter_pairs_kernel(const u32* __restrict__ idx, const u32* __restrict__ val, u64 m, u32* __restrict__ dst) {
    const u64 t0 = (u64(blockIdx.x) * blockDim.x + threadIdx.x) * SP_EPT;
    const bool aligned = ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(val)) & 15u) == 0;  // kernel-uniform
    if (aligned && t0 + SP_EPT <= m) {
        const uint4 i4 = *reinterpret_cast<const uint4*>(idx + t0);
        const uint4 v4 = *reinterpret_cast<const uint4*>(val + t0);
        dst[i4.x] = v4.x;
        dst[i4.y] = v4.y;
        dst[i4.z] = v4.z;
        dst[i4.w] = v4.w;
    } else {
        for (u64 t = t0; t < min(t0 + u64(SP_EPT), m); t++) dst[idx[t]] = val[t];
    }
}

#ifdef TDC_CUSIM
static const u64 PS_DIRECT_BELOW = 1000;  // tiny thresholds so that the CPU tests exercise the partition pass
static const int PS_WINDOW_BITS = 8;
#else
static const u64 PS_DIRECT_BELOW = u64(1) << 22;  // small batches stay L2-resident anyway
#ifndef PS_WINDOW_BITS_CFG
#define PS_WINDOW_BITS_CFG 22
#endif
static const int PS_WINDOW_BITS = PS_WINDOW_BITS_CFG;  // window = 2^22 elements = 16 MiB
#endif

// idx[0]/val[0] hold the pairs; idx[1]/val[1] are scratch of the same size.  n_dst = size of dst (bounds the idx bits).
// full_perm: idx is a permutation of 0..n_dst-1 (m == n_dst), which makes the partition pass histogram-free.
static inline int partitioned_scatter(SortWorkspace& ws, cudaStream_t st, u32* idx[2], u32* val[2], u64 m, u32* dst, u64 n_dst,
                                      bool full_perm = false) {
    if (m == 0) return 0;
    int res = 0;
    const int bits = int(bits_for_host(n_dst > 1 ? n_dst - 1 : 1));
    if (m >= PS_DIRECT_BELOW && bits > PS_WINDOW_BITS) {
        const int wbits = bits - PS_WINDOW_BITS > 8 ? 8 : bits - PS_WINDOW_BITS;  // at most 256 windows
        TDC_TRY(radix_sort_pairs<u32>(ws, st, idx, val, m, bits - wbits, bits, false, &res, full_perm && m == n_dst));
    }
    TDC_LAUNCH(scatter_pairs_kernel, u32(div_up(m, SP_THREADS * SP_EPT)), SP_THREADS, 0, st, idx[res], val[res], m, dst);
    prof_add_bytes("scatter_pairs_kernel", double(m) * 12);
    TDC_KCHECK();
    return 0;
}

}  // namespace tdc
