// Kernels of the lzss_lcp factoriser (see lzss_factorize.cu); shared by the single-GPU and the sharded multi-GPU driver.
#pragma once
#include "tdc_ctx.h"

namespace tdc {

static const int MT_MAX_LEVELS = 8;
struct MinTree {
    const u32* a[MT_MAX_LEVELS];  // level 0 = SA
    const u32* l[MT_MAX_LEVELS];  // level 0 = LCP
    u32 sz[MT_MAX_LEVELS];
    int nlev;
    __device__ __forceinline__ u32 A(int lvl, u32 i) const { return a[lvl][i]; }
    __device__ __forceinline__ u32 L(int lvl, u32 i) const { return l[lvl][i]; }
    __device__ __forceinline__ u32 size(int lvl) const { return sz[lvl]; }
    __device__ __forceinline__ int levels() const { return nlev; }
};

// one warp per output element: min over 32 inputs
static __global__ void __launch_bounds__(256)
mintree_level_kernel(const u32* __restrict__ a_in, const u32* __restrict__ l_in, u32 sz_in, u32* __restrict__ a_out,
                     u32* __restrict__ l_out, u32 sz_out) {
    const u32 o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= sz_out) return;  // whole warps leave together (sz_out is tested per warp)
    const u64 i = u64(o) * 32 + lane_id();
    u32 av = i < sz_in ? a_in[i] : 0xffffffffu;
    u32 lv = i < sz_in ? l_in[i] : 0xffffffffu;
    av = warp_min(av);
    lv = warp_min(lv);
    if (lane_id() == 0) { a_out[o] = av; l_out[o] = lv; }
}

enum : int { WALK_ABANDONED = 0, WALK_FOUND = 1, WALK_OFF_TREE = 2 };

// Nearest rank q < p with SA[q] < v.  m (in: min LCP over the ranks already passed, LCP[p] at the start) becomes
// min LCP[q+1..p].  WALK_ABANDONED: the minimum fell below thr (this side cannot produce a factor);
// WALK_OFF_TREE: no such rank inside this tree (m = minimum over everything passed, so a caller can continue in an
// enclosing tree from the first rank of this one).
template <class Tree>
__device__ __forceinline__ int walk_psv(const Tree& T, u32 p, u32 v, u32 thr, u32& m, u32& q_out) {
    if (m < thr) return WALK_ABANDONED;
    u32 idx = p, q = 0;
    int lvl = 0;
    bool found = false;
    while (!found) {
        const u32 bs = idx & ~31u;
        for (q = idx; q-- > bs;) {
            if (T.A(lvl, q) < v) { found = true; break; }
            m = min(m, T.L(lvl, q));
            if (m < thr) return WALK_ABANDONED;
        }
        if (found) break;
        if (lvl == T.levels() - 1) return WALK_OFF_TREE;
        idx >>= 5;
        lvl++;
    }
    while (lvl > 0) {
        const u32 lo = q * 32u;
        u32 c = min(lo + 32u, T.size(lvl - 1));
        while (c-- > lo) {
            if (T.A(lvl - 1, c) < v) break;
            m = min(m, T.L(lvl - 1, c));
            if (m < thr) return WALK_ABANDONED;
        }
        q = c;
        lvl--;
    }
    q_out = q;
    return WALK_FOUND;
}

// Nearest rank q > p with SA[q] < v; m (in: minimum so far, 0xffffffff at the start) becomes min LCP[p+1..q].
template <class Tree>
__device__ __forceinline__ int walk_nsv(const Tree& T, u32 p, u32 v, u32 thr, u32& m, u32& q_out) {
    if (m < thr) return WALK_ABANDONED;
    u32 idx = p, q = 0;
    int lvl = 0;
    bool found = false;
    while (!found) {
        const u32 be = min((idx | 31u) + 1u, T.size(lvl));
        for (q = idx + 1; q < be; q++) {
            if (T.A(lvl, q) < v) { found = true; break; }
            m = min(m, T.L(lvl, q));
            if (m < thr) return WALK_ABANDONED;
        }
        if (found) break;
        if (lvl == T.levels() - 1) return WALK_OFF_TREE;
        idx >>= 5;
        lvl++;
    }
    while (lvl > 0) {
        u32 c = q * 32u;
        while (true) {
            if (T.A(lvl - 1, c) < v) break;
            m = min(m, T.L(lvl - 1, c));
            if (m < thr) return WALK_ABANDONED;
            c++;
        }
        q = c;
        lvl--;
    }
    m = min(m, T.L(0, q));
    if (m < thr) return WALK_ABANDONED;
    q_out = q;
    return WALK_FOUND;
}

// Per rank: longest previous factor length and winning side.  A tile of LPF_TILE consecutive ranks of SA and LCP is
// staged in shared memory together with two local min-tree levels; almost every PSV/NSV walk ends inside its tile at
// shared-memory latency (a walk is a chain of dependent loads), the few that leave it continue in the global tree.
// Output in rank order; the partitioned scatter that follows moves it to text order (index side = SA itself).
#ifdef TDC_CUSIM
static const int LPF_THREADS = 128;  // small tiles so that the CPU tests leave their tile often
static const int LPF_TILE = 1024;
#else
static const int LPF_THREADS = 512;
static const int LPF_TILE = 4096;
#endif
static const int LPF_L1 = LPF_TILE / 32;  // 128
static const int LPF_L2 = LPF_L1 / 32;    // 4

// the tile's three levels lie back to back in shared memory: no pointer table, no dynamic indexing
struct TileTree {
    const u32* sA;
    const u32* sL;
    __device__ __forceinline__ static u32 off(int lvl) { return lvl == 0 ? 0u : (lvl == 1 ? u32(LPF_TILE) : u32(LPF_TILE + LPF_L1)); }
    __device__ __forceinline__ u32 A(int lvl, u32 i) const { return sA[off(lvl) + i]; }
    __device__ __forceinline__ u32 L(int lvl, u32 i) const { return sL[off(lvl) + i]; }
    __device__ __forceinline__ u32 size(int lvl) const { return lvl == 0 ? u32(LPF_TILE) : (lvl == 1 ? u32(LPF_L1) : u32(LPF_L2)); }
    __device__ __forceinline__ int levels() const { return 3; }
};

static __global__ void __launch_bounds__(LPF_THREADS)
lpf_tile_kernel(MinTree T, u32 n, u32 thr, u32* __restrict__ out_lenside) {
    __shared__ u32 sA[LPF_TILE + LPF_L1 + LPF_L2];
    __shared__ u32 sL[LPF_TILE + LPF_L1 + LPF_L2];
    const u32 base = blockIdx.x * LPF_TILE;
    for (u32 j = threadIdx.x; j < LPF_TILE; j += LPF_THREADS) {
        const u32 i = base + j;
        sA[j] = i < n ? T.a[0][i] : 0xffffffffu;
        sL[j] = i < n ? T.l[0][i] : 0xffffffffu;
    }
    __syncthreads();
    for (u32 g = warp_id(); g < LPF_L1; g += LPF_THREADS / 32) {
        const u32 av = warp_min(sA[g * 32 + lane_id()]);
        const u32 lv = warp_min(sL[g * 32 + lane_id()]);
        if (lane_id() == 0) { sA[LPF_TILE + g] = av; sL[LPF_TILE + g] = lv; }
    }
    __syncthreads();
    if (warp_id() < LPF_L2) {
        const u32 av = warp_min(sA[LPF_TILE + warp_id() * 32 + lane_id()]);
        const u32 lv = warp_min(sL[LPF_TILE + warp_id() * 32 + lane_id()]);
        if (lane_id() == 0) { sA[LPF_TILE + LPF_L1 + warp_id()] = av; sL[LPF_TILE + LPF_L1 + warp_id()] = lv; }
    }
    __syncthreads();
    TileTree S;
    S.sA = sA;
    S.sL = sL;
    const u32 last = min(base + u32(LPF_TILE), n) - 1u;  // last rank of this tile
    for (u32 j = threadIdx.x; j < LPF_TILE; j += LPF_THREADS) {
        const u32 p = base + j;
        if (p >= n) break;
        const u32 v = sA[j];
        u32 q;
        u32 mu = sL[j];
        int r = walk_psv(S, j, v, thr, mu, q);
        if (r == WALK_OFF_TREE) r = walk_psv(T, base, v, thr, mu, q);
        const u32 lu = r == WALK_FOUND ? mu : 0u;
        u32 md = 0xffffffffu;
        r = walk_nsv(S, j, v, thr, md, q);
        if (r == WALK_OFF_TREE) r = walk_nsv(T, last, v, thr, md, q);
        const u32 ld = r == WALK_FOUND ? md : 0u;
        const u32 len = max(lu, ld);
        out_lenside[p] = len >= thr ? ((len << 1) | (lu >= ld ? 0u : 1u)) : 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// greedy chain
// ---------------------------------------------------------------------------------------------------------------
static const int CH_THREADS = 512;
#ifdef TDC_CUSIM
static const int CH_IPT = 2;  // small tiles so that the CPU tests cross many tile/region boundaries
#else
static const int CH_IPT = 16;
#endif
static const int CH_TILE = CH_THREADS * CH_IPT;  // text positions per tile
static const u32 CH_NONE = 0xffffffffu;

__device__ __forceinline__ u32 next_of(u32 i, u32 ls) {
    const u32 len = ls >> 1;
    return i + (len ? len : 1u);
}

// where does each position leave its tile?  nodes are positions < n-1; anything >= n-1 is terminal.
static __global__ void __launch_bounds__(CH_THREADS)
chain_exit_kernel(const u32* __restrict__ lenside, u32 n, u32* __restrict__ exitp) {
    __shared__ u32 J[CH_TILE];
    __shared__ u32 changed;
    const u32 base = blockIdx.x * CH_TILE;
    const u32 tile_end = min(base + u32(CH_TILE), n - 1);
    for (u32 j = threadIdx.x; j < CH_TILE; j += CH_THREADS) {
        const u32 i = base + j;
        J[j] = i < tile_end ? next_of(i, lenside[i]) : CH_NONE;
    }
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        bool any = false;
        for (u32 j = threadIdx.x; j < CH_TILE; j += CH_THREADS) {
            const u32 t = J[j];
            if (t < tile_end) {  // still inside: hop through the target's current pointer (always a node on j's path)
                J[j] = J[t - base];
                any = true;
            }
        }
        if (any) changed = 1;
        __syncthreads();
        const bool again = changed != 0;
        __syncthreads();
        if (!again) break;
    }
    for (u32 j = threadIdx.x; j < CH_TILE; j += CH_THREADS) {
        const u32 i = base + j;
        if (i < tile_end) exitp[i] = J[j];
    }
}

// The first visited position ("entry") of every tile the chain touches.  The chain is a dependent pointer walk over
// tile exits, so it is split: `regions` walkers start speculatively at their region's first position (as if it were
// visited) and record the entries of their own path; a scalar stitcher then follows the TRUE chain and, in each
// region, only walks until it lands on a node the region's walker also visited — from there on the two paths are
// identical, so the walker's remaining entries are already right and the stitcher jumps to the walker's exit.
// Entries the stitcher skips over (speculative but not on the true chain) are erased.
static __global__ void __launch_bounds__(128)
chain_entries_spec_kernel(const u32* __restrict__ exitp, u32 n, u32 tiles_per_region, u32 regions,
                          u32* __restrict__ entry, u32* __restrict__ region_exit) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= regions) return;
    const u64 start = u64(r) * tiles_per_region * CH_TILE;
    const u64 end = min(start + u64(tiles_per_region) * CH_TILE, u64(n - 1));
    u64 x = start;
    while (x < end) {
        entry[x / CH_TILE] = u32(x);
        x = exitp[x];
    }
    region_exit[r] = u32(min(x, u64(0xffffffffu)));
}

static __global__ void chain_entries_stitch_kernel(const u32* __restrict__ exitp, u32 n, u32 tiles_per_region,
                                            u32* __restrict__ entry, const u32* __restrict__ region_exit, u32 first,
                                            u32* __restrict__ last_out) {
    // first: where the chain enters this position range (0 on a single GPU); *last_out: where it leaves it
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 x = first, clear_from = 0;
    while (x < n - 1) {
        const u32 t = x / CH_TILE;
        for (u32 tt = clear_from; tt < t; tt++) entry[tt] = CH_NONE;
        if (entry[t] == x) {  // merged with the region walker's path
            const u32 r = t / tiles_per_region;
            x = region_exit[r];
            clear_from = (r + 1) * tiles_per_region;
        } else {
            entry[t] = x;
            x = exitp[x];
            clear_from = t + 1;
        }
    }
    const u32 ntiles = (n + CH_TILE - 1) / CH_TILE;
    for (u32 tt = clear_from; tt < ntiles; tt++) entry[tt] = CH_NONE;
    if (last_out) *last_out = x;
}

// Mark the visited positions of every tile: with the tile entries known the tiles are independent, so ONE THREAD walks
// one tile's chain (a few hundred dependent, mostly L1-resident loads) while hundreds of thousands of tiles are in
// flight.  Output: one bit per position that starts a factor (fmask is pre-zeroed) and the per-tile factor count.
static __global__ void __launch_bounds__(128)
chain_mark_kernel(const u32* __restrict__ lenside, u32 n, u32 ntiles, const u32* __restrict__ entry,
                  u32* __restrict__ fmask, u32* __restrict__ tile_count) {
    const u32 tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    u32 x = entry[tile];
    u32 cnt = 0;
    if (x != CH_NONE) {
        const u32 tile_end = u32(min(u64(tile) * CH_TILE + CH_TILE, u64(n - 1)));
        u32 word = x >> 5, bits = 0;
        while (x < tile_end) {
            const u32 ls = lenside[x];
            if (ls) { bits |= 1u << (x & 31); cnt++; }
            x = next_of(x, ls);
            if ((x >> 5) != word) {
                if (bits) fmask[word] = bits;
                word = x >> 5;
                bits = 0;
            }
        }
        if (bits) fmask[word] = bits;  // unreachable (the word changes when x leaves it); kept for clarity
    }
    tile_count[tile] = cnt;
}

// single CTA: exclusive scan of per-tile counts; *total = sum
static __global__ void __launch_bounds__(1024) scan_counts_kernel(u32* __restrict__ cnt, u32 ntiles, u32* __restrict__ total) {
    __shared__ u32 scratch[33];
    u32 carry = 0;
    for (u32 b = 0; b < ntiles; b += 1024) {
        const u32 i = b + threadIdx.x;
        const u32 c = i < ntiles ? cnt[i] : 0;
        u32 tot;
        const u32 ex = block_exclusive_sum<u32>(c, scratch, &tot);
        if (i < ntiles) cnt[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}

// emit (pos, src, len) in position order; one thread per 32-bit mask word.  HAVE_SRC: the source positions were
// carried along (src_arr, sharded multi-GPU path) instead of being recovered by a walk; pos_base = first text position
// of this rank's range.
template <bool HAVE_SRC>
static __global__ void __launch_bounds__(CH_TILE / 32)
emit_factors_kernel(MinTree T, const u32* __restrict__ isa, const u32* __restrict__ lenside, const u32* __restrict__ fmask,
                    const u32* __restrict__ tile_off, u32 thr, Factor* __restrict__ out, u32* __restrict__ minmax,
                    const u32* __restrict__ src_arr, u32 pos_base) {
    __shared__ u32 scratch[33];
    __shared__ u32 s_min[CH_TILE / 32 / 32], s_max[CH_TILE / 32 / 32];
    const u32 base = blockIdx.x * CH_TILE;
    u32 word = fmask[u64(blockIdx.x) * (CH_TILE / 32) + threadIdx.x];
    u32 tot;
    u32 o = tile_off[blockIdx.x] + block_exclusive_sum<u32>(u32(__popc(word)), scratch, &tot);
    u32 mn = 0xffffffffu, mx = 0;
    while (word) {
        const u32 b = __ffs(int(word)) - 1;
        word &= word - 1;
        const u32 i = base + threadIdx.x * 32 + b;
        const u32 ls = lenside[i];
        const u32 len = ls >> 1;
        Factor f;
        if (HAVE_SRC) {
            f.src = src_arr[i];
        } else {
            const u32 p = isa[i];
            u32 q = 0, m = 0xffffffffu;
            if (ls & 1u) {
                walk_nsv(T, p, i, thr, m, q);
            } else {
                m = T.l[0][p];
                walk_psv(T, p, i, thr, m, q);
            }
            f.src = T.a[0][q];
        }
        f.pos = pos_base + i;
        f.len = len;
        out[o++] = f;
        mn = min(mn, len);
        mx = max(mx, len);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane_id() == 0) { s_min[warp_id()] = mn; s_max[warp_id()] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (u32 w = 0; w < CH_TILE / 32 / 32; w++) { mn = min(mn, s_min[w]); mx = max(mx, s_max[w]); }
        if (mx) { atomicMin(&minmax[0], mn); atomicMax(&minmax[1], mx); }
    }
}

static __global__ void fill_u32_kernel(u32* p, u64 count, u32 v) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

}  // namespace tdc
