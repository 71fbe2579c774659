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
__global__ void __launch_bounds__(256)
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

// ---------------------------------------------------------------------------------------------------------------
// Per rank: longest previous factor length and winning side.
//
// A tile of LPF_TILE consecutive ranks of SA and LCP is staged in shared memory; ONE THREAD owns a chunk of 32
// consecutive ranks and runs the sequential all-nearest-smaller-values recurrence over it: "pop" = follow the PSV
// pointer of the current candidate, carrying the LCP minimum of the skipped range.  That is amortised O(1) per rank
// whatever the distribution of distances, so a warp never waits for one lane's long linear walk (the first version, one
// independent walk per rank, spent 50 instructions per rank on exactly that divergence).  Only a chunk's prefix minima
// (PSV side) / suffix minima (NSV side) are unresolved; their answers are nested, so ONE continuing tree walk per chunk
// and side resolves all of them: three tile-local levels in shared memory first, the global tree for the few that leave
// the tile.  Output in rank order; the partitioned scatter that follows moves it to text order.
// ---------------------------------------------------------------------------------------------------------------
#ifdef TDC_CUSIM
static const int LPF_THREADS = 32;  // small tiles so that the CPU tests leave their tile often
#else
static const int LPF_THREADS = 128;
#endif
static const int LPF_TILE = LPF_THREADS * 32;
static const int LPF_L1 = LPF_TILE / 32;           // one entry per chunk
static const int LPF_L2 = (LPF_L1 + 31) / 32;
static const u32 LPF_INF = 0xffffffffu;

// level 0 is XOR-swizzled: a thread walking its own chunk (index t*32 + s) and a warp reading 32 consecutive ranks
// both touch 32 different banks
__device__ __forceinline__ u32 lpf_phys(u32 i) { return i ^ ((i >> 5) & 31u); }

struct TileTree {
    const u32* sA;  // [LPF_TILE] swizzled, then LPF_L1, then LPF_L2 plain
    const u32* sL;
    __device__ __forceinline__ u32 A(int lvl, u32 i) const {
        return lvl == 0 ? sA[lpf_phys(i)] : (lvl == 1 ? sA[LPF_TILE + i] : sA[LPF_TILE + LPF_L1 + i]);
    }
    __device__ __forceinline__ u32 L(int lvl, u32 i) const {
        return lvl == 0 ? sL[lpf_phys(i)] : (lvl == 1 ? sL[LPF_TILE + i] : sL[LPF_TILE + LPF_L1 + i]);
    }
    __device__ __forceinline__ u32 size(int lvl) const { return lvl == 0 ? u32(LPF_TILE) : (lvl == 1 ? u32(LPF_L1) : u32(LPF_L2)); }
    __device__ __forceinline__ int levels() const { return 3; }
};

static inline size_t lpf_smem_bytes() {
    return sizeof(u32) * (2 * (LPF_TILE + LPF_L1 + LPF_L2) + 2 * LPF_TILE) + LPF_TILE;
}

__global__ void __launch_bounds__(LPF_THREADS)
lpf_tile_kernel(MinTree T, u32 n, u32 thr, u32* __restrict__ out_lenside) {
    TDC_DYN_SMEM(smem_raw);
    u32* sA = reinterpret_cast<u32*>(smem_raw);
    u32* sL = sA + (LPF_TILE + LPF_L1 + LPF_L2);
    u32* sU = sL + (LPF_TILE + LPF_L1 + LPF_L2);  // l_up per rank, later the combined result (swizzled like level 0)
    u32* sD = sU + LPF_TILE;                       // l_dn per rank
    uint8_t* sP = reinterpret_cast<uint8_t*>(sD + LPF_TILE);  // chunk-local PSV / NSV pointer (+1, 0 = none), swizzled
    const u32 base = blockIdx.x * LPF_TILE;
    for (u32 j = threadIdx.x; j < LPF_TILE; j += LPF_THREADS) {
        const u32 i = base + j;
        sA[lpf_phys(j)] = i < n ? T.a[0][i] : LPF_INF;
        sL[lpf_phys(j)] = i < n ? T.l[0][i] : LPF_INF;
    }
    __syncthreads();
    const u32 cs = threadIdx.x * 32u;  // this thread's chunk [cs, cs + 32)
    const u32 sw = threadIdx.x & 31u;  // swizzle of the chunk: phys(cs + s) = cs + (s ^ sw)

    // ---- pass 1: PSV inside the chunk, left to right; flat loop (every iteration either pops or finishes a rank) ----
    u32 unres_up = 0;  // bit s: rank cs+s has no smaller value to its left inside the chunk
    {
        u32 s = 0, v = sA[cs + (0 ^ sw)], m = sL[cs + (0 ^ sw)];
        int j = -1;  // candidate (chunk-local), -1 = none left
        u32 amin = v, lmin = m;
        while (true) {
            if (j >= 0 && sA[cs + (u32(j) ^ sw)] > v) {
                m = min(m, sU[cs + (u32(j) ^ sw)]);
                j = int(sP[cs + (u32(j) ^ sw)]) - 1;
            } else {
                sU[cs + (s ^ sw)] = m;
                sP[cs + (s ^ sw)] = uint8_t(j + 1);
                if (j < 0) unres_up |= 1u << s;
                if (++s == 32) break;
                v = sA[cs + (s ^ sw)];
                m = sL[cs + (s ^ sw)];
                amin = min(amin, v);
                lmin = min(lmin, m);
                j = int(s) - 1;
            }
        }
        sA[LPF_TILE + threadIdx.x] = amin;
        sL[LPF_TILE + threadIdx.x] = lmin;
    }
    __syncthreads();
    if (threadIdx.x < LPF_L2) {
        u32 av = LPF_INF, lv = LPF_INF;
        for (u32 k = threadIdx.x * 32; k < min(threadIdx.x * 32 + 32, u32(LPF_L1)); k++) {
            av = min(av, sA[LPF_TILE + k]);
            lv = min(lv, sL[LPF_TILE + k]);
        }
        sA[LPF_TILE + LPF_L1 + threadIdx.x] = av;
        sL[LPF_TILE + LPF_L1 + threadIdx.x] = lv;
    }
    __syncthreads();
    TileTree S;
    S.sA = sA;
    S.sL = sL;
    // ---- PSV of the chunk's prefix minima: one continuing walk to the left (values fall, answers move left) ----
    {
        u32 m = LPF_INF, pos = cs;
        int mode = 0;  // 0: inside the tile, 1: global tree, 2: no smaller value further left / below the threshold
        u32 bits = unres_up;
        while (bits) {
            const u32 s = __ffs(int(bits)) - 1;
            bits &= bits - 1;
            const u32 v = sA[cs + (s ^ sw)];
            u32 lu = 0;
            if (v != LPF_INF && mode != 2) {
                m = min(m, sU[cs + (s ^ sw)]);  // + min LCP[cs..cs+s]
                u32 q = 0;
                int r = WALK_OFF_TREE;
                if (mode == 0) {
                    r = walk_psv(S, pos, v, thr, m, q);
                    if (r == WALK_FOUND) pos = q + 1;
                    if (r == WALK_OFF_TREE) { mode = 1; pos = base; }
                }
                if (mode == 1 && r == WALK_OFF_TREE) {
                    r = walk_psv(T, pos, v, thr, m, q);
                    if (r == WALK_FOUND) pos = q + 1;
                }
                if (r == WALK_FOUND) lu = m; else mode = 2;
            }
            sU[cs + (s ^ sw)] = lu;
        }
    }
    // ---- pass 2: NSV inside the chunk, right to left ----
    u32 unres_dn = 0;
    {
        int s = 31;
        u32 v = sA[cs + (31u ^ sw)], m = LPF_INF;
        u32 j = 32;  // candidate (chunk-local), 32 = none left
        bool fresh = true;  // candidate j has not been examined yet (its own LCP is not in m)
        while (true) {
            bool finish = j >= 32;
            if (!finish) {
                const u32 pj = cs + (j ^ sw);
                if (fresh) m = min(m, sL[pj]);
                if (sA[pj] < v) {
                    finish = true;
                } else {
                    m = min(m, sD[pj]);
                    const u32 nx = sP[pj];
                    j = nx ? nx - 1 : 32u;
                    fresh = false;  // LCP[N[j]] is part of l_dn[j]
                }
            }
            if (finish) {
                sD[cs + (u32(s) ^ sw)] = m;
                sP[cs + (u32(s) ^ sw)] = uint8_t(j < 32 ? j + 1 : 0);
                if (j >= 32) unres_dn |= 1u << s;
                if (--s < 0) break;
                v = sA[cs + (u32(s) ^ sw)];
                m = LPF_INF;
                j = u32(s) + 1;
                fresh = true;
            }
        }
    }
    // ---- NSV of the chunk's suffix minima: one continuing walk to the right ----
    {
        const u32 last = min(base + u32(LPF_TILE), n) - 1u;  // last rank of this tile
        u32 m = LPF_INF, pos = cs + 31;
        int mode = 0;
        u32 bits = unres_dn;
        while (bits) {
            const u32 s = 31u - u32(__clz(int(bits)));
            bits &= ~(1u << s);
            const u32 v = sA[cs + (s ^ sw)];
            u32 ld = 0;
            if (v != LPF_INF && mode != 2) {
                m = min(m, sD[cs + (s ^ sw)]);  // + min LCP[cs+s+1..cs+31]
                u32 q = 0;
                int r = WALK_OFF_TREE;
                if (mode == 0) {
                    r = walk_nsv(S, pos, v, thr, m, q);
                    if (r == WALK_FOUND) pos = q - 1;
                    if (r == WALK_OFF_TREE) { mode = 1; pos = last; }
                }
                if (mode == 1 && r == WALK_OFF_TREE) {
                    r = walk_nsv(T, pos, v, thr, m, q);
                    if (r == WALK_FOUND) pos = q - 1;
                }
                if (r == WALK_FOUND) ld = m; else mode = 2;
            }
            sD[cs + (s ^ sw)] = ld;
        }
    }
    // ---- combine (PSV wins ties, LZSSLCPCompressor.hpp:101) ----
#pragma unroll 4
    for (u32 s = 0; s < 32; s++) {
        const u32 lu = sU[cs + (s ^ sw)], ld = sD[cs + (s ^ sw)];
        const u32 len = max(lu, ld);
        sU[cs + (s ^ sw)] = len >= thr ? ((len << 1) | (lu >= ld ? 0u : 1u)) : 0u;
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < LPF_TILE; j += LPF_THREADS) {
        const u32 p = base + j;
        if (p < n) out_lenside[p] = sU[lpf_phys(j)];
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
__global__ void __launch_bounds__(CH_THREADS)
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
__global__ void __launch_bounds__(128)
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

__global__ void chain_entries_stitch_kernel(const u32* __restrict__ exitp, u32 n, u32 tiles_per_region,
                                            u32* __restrict__ entry, const u32* __restrict__ region_exit) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 x = 0, clear_from = 0;
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
}

// Mark the visited positions of every tile: with the tile entries known the tiles are independent, so ONE THREAD walks
// one tile's chain (a few hundred dependent, mostly L1-resident loads) while hundreds of thousands of tiles are in
// flight.  Output: one bit per position that starts a factor (fmask is pre-zeroed) and the per-tile factor count.
__global__ void __launch_bounds__(128)
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
__global__ void __launch_bounds__(1024) scan_counts_kernel(u32* __restrict__ cnt, u32 ntiles, u32* __restrict__ total) {
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

// emit (pos, src, len) in position order; one thread per 32-bit mask word
__global__ void __launch_bounds__(CH_TILE / 32)
emit_factors_kernel(MinTree T, const u32* __restrict__ isa, const u32* __restrict__ lenside, const u32* __restrict__ fmask,
                    const u32* __restrict__ tile_off, u32 thr, Factor* __restrict__ out, u32* __restrict__ minmax) {
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
        const u32 p = isa[i];
        u32 q = 0, m = 0xffffffffu;
        if (ls & 1u) {
            walk_nsv(T, p, i, thr, m, q);
        } else {
            m = T.l[0][p];
            walk_psv(T, p, i, thr, m, q);
        }
        Factor f;
        f.pos = i;
        f.src = T.a[0][q];
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

__global__ void fill_u32_kernel(u32* p, u64 count, u32 v) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

int factorize_lzss_lcp(Ctx& c, u32 threshold) {
    if (threshold < 1) { set_error("lzss_lcp: threshold must be >= 1"); return -5; }
    if ((c.have & (DS_SA | DS_ISA | DS_LCP)) != (DS_SA | DS_ISA | DS_LCP)) {
        set_error("lzss_lcp: SA, ISA and LCP must be built first");
        return -6;
    }
    const u32 n = u32(c.n);
    cudaStream_t st = c.stream;
    c.num_factors = 0;
    c.flen_min = 0xffffffffu;
    c.flen_max = 0;
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
        TDC_LAUNCH(mintree_level_kernel, u32(div_up(u64(szo) * 32, 256)), 256, 0, st, T.a[T.nlev - 1], T.l[T.nlev - 1], szi, a, l, szo);
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
        TDC_CUDA(cudaFuncSetAttribute(lpf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lpf_smem_bytes())));
        TDC_LAUNCH(lpf_tile_kernel, u32(div_up(u64(n), LPF_TILE)), LPF_THREADS, lpf_smem_bytes(), st, T, n, threshold, sc_val[0]);
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
        TDC_LAUNCH(chain_entries_stitch_kernel, 1, 32, 0, st, exitp, n, tiles_per_region, entry, region_exit);
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
        TDC_LAUNCH(emit_factors_kernel, ntiles, CH_TILE / 32, 0, st, T, c.d_isa, lenside, fmask, tile_cnt, threshold, c.d_factors, d_minmax);
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
