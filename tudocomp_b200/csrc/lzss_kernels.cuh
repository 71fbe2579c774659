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

// one warp per 4 output elements (128 inputs): every lane takes 4 consecutive inputs with one 16-byte load per array,
// 8 lanes form one output
static const int MT_OUT_PER_WARP = 4;
static __global__ void __launch_bounds__(256)
mintree_level_kernel(const u32* __restrict__ a_in, const u32* __restrict__ l_in, u32 sz_in, u32* __restrict__ a_out,
                     u32* __restrict__ l_out, u32 sz_out) {
    const u32 wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 o0 = wi * MT_OUT_PER_WARP;
    if (o0 >= sz_out) return;  // whole warps leave together
    const u64 i = u64(o0) * 32 + u64(lane_id()) * 4;
    u32 av = 0xffffffffu, lv = 0xffffffffu;
    if (i + 4 <= sz_in) {
        const uint4 a = *reinterpret_cast<const uint4*>(a_in + i);
        const uint4 l = *reinterpret_cast<const uint4*>(l_in + i);
        av = min(min(a.x, a.y), min(a.z, a.w));
        lv = min(min(l.x, l.y), min(l.z, l.w));
    } else {
        for (u32 q = 0; q < 4; q++)
            if (i + q < sz_in) { av = min(av, a_in[i + q]); lv = min(lv, l_in[i + q]); }
    }
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) {
        av = min(av, __shfl_xor_sync(kFull, av, d));
        lv = min(lv, __shfl_xor_sync(kFull, lv, d));
    }
    const u32 o = o0 + (lane_id() >> 3);
    if ((lane_id() & 7u) == 0 && o < sz_out) { a_out[o] = av; l_out[o] = lv; }
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
// Per rank: longest previous factor length and winning side (all nearest smaller values of SA + LCP range minima).
//
// A tile of LPF_TILE consecutive ranks is staged in shared memory and solved in three steps:
//  1. ONE THREAD per chunk of 32 ranks runs the sequential nearest-smaller-value recurrence ("pop" = follow the pointer
//     of the current candidate, carrying the LCP minimum of the skipped range): amortised O(1) per rank, whatever the
//     distribution of distances.  Written as a flat, predicated loop (one candidate per iteration).
//     Unresolved afterwards: the chunk's prefix minima (PSV side) and suffix minima (NSV side).
//  2. a binary merge tree over the chunks: the suffix minima of the left node and the prefix minima of the right node
//     are two lists of falling values, linked by the very PSV/NSV pointers already computed; one two-pointer merge
//     resolves the PSV of the right list and the NSV of the left list at once (O(list length) per merge), the
//     leftovers form the lists of the merged node.  log2(chunks) levels, half of the threads each time.
//  3. what is still unresolved (the tile's own prefix/suffix minima, ~ln(LPF_TILE) per side) walks the global min-tree.
// Output in rank order; the partitioned scatter that follows moves it to text order.
// (History, profiles/r1d_ncu_summary.md: one independent linear walk per rank spends ~50 instructions per rank on
// divergence; per-chunk recurrences + tree walks for the leftovers spend even more on the walks.)
// ---------------------------------------------------------------------------------------------------------------
#ifndef LPF_CHUNK_CFG
#define LPF_CHUNK_CFG 8
#endif
static const int LPF_CHUNK = LPF_CHUNK_CFG;  // ranks per thread (4, 8, 16 or 32: a chunk must not straddle 32 ranks, see lpf_phys)
#if defined(TDC_CUSIM) && !defined(LPF_THREADS_CFG)
static const int LPF_THREADS = 32;  // small tiles so that the CPU tests leave their tile often
#else
#ifndef LPF_THREADS_CFG
#define LPF_THREADS_CFG 128  // 128 chunks of 8 ranks = tiles of 1024 ranks: measured best (profiles/r2v_lpf_variants.txt:
                             // 5.92 ms at dna 2^28 against 6.33 for 64 x 16, 5.94 for 32 x 16, 6.6 for 256 x 8, 7.5-9.4 for chunks of 4)
#endif
static const int LPF_THREADS = LPF_THREADS_CFG;  // chunks per tile (power of two)
#endif
static const int LPF_TILE = LPF_THREADS * LPF_CHUNK;
static const u32 LPF_INF = 0xffffffffu;
static const u32 LPF_NONE = 0xffffffffu;

// XOR swizzle: the threads of a warp walking their own chunks (index t*LPF_CHUNK + s, same s) and a warp reading 32
// consecutive ranks both touch 32 different banks
__device__ __forceinline__ u32 lpf_phys(u32 i) { return i ^ ((i >> 5) & 31u); }

static inline size_t lpf_smem_bytes() { return sizeof(u32) * 3 * LPF_TILE + sizeof(unsigned short) * 2 * LPF_TILE + sizeof(u32) * (LPF_THREADS + 2 * LPF_TILE / 16 + 8); }

// DIST (sharded multi-GPU path): l_up / l_dn and their source positions are written separately and walks that find
// nothing smaller inside this shard are queued for the neighbouring shards (dist_textds.cu).
struct WalkQuery {  // p: slot index inside the origin shard; v = SA[p]; m = LCP minimum collected so far
    u32 p, v, m;
};
struct LpfDistOut {
    u32 *lu, *su, *ld, *sd;
    WalkQuery *q_up, *q_dn;
    u32* q_cnt;  // [2]
    u32 qcap;
};

template <bool DIST>
__device__ __forceinline__ void lpf_resolve_open(const MinTree& T, u32 base, u32 last, u32 thr, u32 e, u32 side, u32* sA, u32* sU, u32* sD,
                                                 const LpfDistOut& D) {
    const u32 pe = lpf_phys(e);
    const u32 v = sA[pe];
    if (v == LPF_INF) return;  // padding past the end of the array
    u32 q = 0;
    if (side == 0) {
        u32 m = sU[pe];  // min LCP[tile start .. e]
        const int r = walk_psv(T, base, v, thr, m, q);
        sU[pe] = r == WALK_FOUND ? m : 0u;
        if (DIST) {
            D.su[base + e] = r == WALK_FOUND ? T.a[0][q] : 0u;
            if (r == WALK_OFF_TREE) {
                const u32 slot = atomicAdd(&D.q_cnt[0], 1u);
                if (slot < D.qcap) { WalkQuery w; w.p = base + e; w.v = v; w.m = m; D.q_up[slot] = w; }
            }
        }
    } else {
        u32 m = sD[pe];  // min LCP[e + 1 .. tile end)
        const int r = walk_nsv(T, last, v, thr, m, q);
        sD[pe] = r == WALK_FOUND ? m : 0u;
        if (DIST) {
            D.sd[base + e] = r == WALK_FOUND ? T.a[0][q] : 0u;
            if (r == WALK_OFF_TREE) {
                const u32 slot = atomicAdd(&D.q_cnt[1], 1u);
                if (slot < D.qcap) { WalkQuery w; w.p = base + e; w.v = v; w.m = m; D.q_dn[slot] = w; }
            }
        }
    }
}

template <bool DIST>
static __global__ void __launch_bounds__(LPF_THREADS)
lpf_tile_kernel(MinTree T, u32 n, u32 thr, u32* __restrict__ out_lenside, LpfDistOut D) {
    TDC_DYN_SMEM(smem_raw);
    u32* sA = reinterpret_cast<u32*>(smem_raw);            // SA values
    u32* sU = sA + LPF_TILE;                                // LCP values, overwritten in place by l_up
    u32* sD = sU + LPF_TILE;                                // l_dn
    unsigned short* sPp = reinterpret_cast<unsigned short*>(sD + LPF_TILE);  // PSV pointer (tile-local index + 1, 0 = open)
    unsigned short* sPn = sPp + LPF_TILE;                                    // NSV pointer
    u32* sNL = reinterpret_cast<u32*>(sPn + LPF_TILE);     // [LPF_THREADS] LCP minimum of the node starting at chunk t
    u32* sQ = sNL + LPF_THREADS;                            // [2 * LPF_TILE / 16] open (index << 1 | side) after the merges
    u32* sQn = sQ + 2 * LPF_TILE / 16;                      // [1] queue length
    const u32 base = blockIdx.x * LPF_TILE;
    constexpr u32 CH = u32(LPF_CHUNK);
    // Every warp loads the 32 chunks its own lanes will work on, so that steps 1 and the first five merge levels only
    // need warp-level synchronisation: with block barriers after every level the barrier was the top stall reason
    // (9.7 stalled warps per issue, profiles/r1i_summary.md) because the upper levels keep 1-2 warps busy.
    {
        const u32 wbase = warp_id() * 32u * CH;
#pragma unroll 4
        for (u32 k = 0; k < CH; k++) {
            const u32 j = wbase + k * 32u + lane_id();
            const u32 i = base + j;
            sA[lpf_phys(j)] = i < n ? T.a[0][i] : LPF_INF;
            sU[lpf_phys(j)] = i < n ? T.l[0][i] : LPF_INF;
        }
    }
    if (threadIdx.x == 0) *sQn = 0;
    __syncwarp();
    const u32 cs = threadIdx.x * CH;  // this thread's chunk [cs, cs + CH)
    const u32 sw = (cs >> 5) & 31u;   // swizzle of the chunk: phys(cs + s) = (cs + s) ^ sw
#define LPF_AT(arr, s) arr[(cs + (s)) ^ sw]

    // ---- 1a. NSV inside the chunk, right to left (reads the raw LCP values) ----
    // One candidate per iteration, written with selects so that lanes that finish a rank and lanes that pop a candidate
    // run the same instructions (a branchy version ran at 15 of 32 active lanes, profiles/r1f_ncu_summary.md).
    {
        int s = int(CH) - 1;
        u32 v = LPF_AT(sA, CH - 1), m = LPF_INF, j = CH;
        do {
            const bool valid = j < CH;
            const u32 jj = valid ? j : CH - 1;
            const u32 aj = LPF_AT(sA, jj), lj = LPF_AT(sU, jj), dj = LPF_AT(sD, jj), nj = LPF_AT(sPn, jj);
            const u32 m1 = valid ? min(m, lj) : m;  // LCP[j] lies inside the range whether or not j is the answer
            const bool found = valid && aj < v;
            const bool fin = !valid || found;
            if (fin) {
                LPF_AT(sD, u32(s)) = m1;
                LPF_AT(sPn, u32(s)) = (unsigned short)(found ? cs + j + 1 : 0);
            }
            j = fin ? u32(s) : (nj ? nj - 1 - cs : CH);  // next rank s-1 starts at candidate s
            m = fin ? LPF_INF : min(m1, dj);
            s -= fin ? 1 : 0;
            v = LPF_AT(sA, u32(max(s, 0)));
        } while (s >= 0);
    }
    // ---- 1b. PSV inside the chunk, left to right; l_up replaces the LCP value in place ----
    {
        u32 s = 0, v = LPF_AT(sA, 0u), m = LPF_AT(sU, 0u), lmin = m;
        int j = -1;
        do {
            const bool valid = j >= 0;
            const u32 jj = valid ? u32(j) : 0u;
            const u32 aj = LPF_AT(sA, jj), uj = LPF_AT(sU, jj), pj = LPF_AT(sPp, jj);
            const bool found = valid && aj < v;
            const bool fin = !valid || found;
            if (fin) {
                LPF_AT(sU, s) = m;
                LPF_AT(sPp, s) = (unsigned short)(found ? cs + u32(j) + 1 : 0);
            }
            j = fin ? int(s) : (pj ? int(pj - 1 - cs) : -1);  // next rank s+1 starts at candidate s
            s += fin ? 1u : 0u;
            const u32 sn = min(s, CH - 1);
            const u32 raw = LPF_AT(sU, sn);  // still the raw LCP value when a new rank starts
            v = LPF_AT(sA, sn);
            lmin = fin ? min(lmin, raw) : lmin;
            m = fin ? raw : min(m, uj);
        } while (s < CH);
        sNL[threadIdx.x] = lmin;
    }
    __syncwarp();
    // ---- 2. merge tree ----
    // The node made of chunks [ca, ca + 2 * half) is merged by thread ca (the thread of its first chunk): nodes of up to
    // 32 chunks live inside one warp and are ordered by __syncwarp, only the last log2(warps) levels need the block.
    for (u32 half = 1; half < u32(LPF_THREADS); half <<= 1) {  // half = chunks per child node
        if (half >= 32) __syncthreads();
        if ((threadIdx.x & (2 * half - 1)) == 0) {
            const u32 ca = threadIdx.x, cb = ca + half;  // first chunks of the left / right child
            u32 a = cb * CH - 1, b = cb * CH;                        // heads: last rank of A, first rank of B
            u32 va = sA[lpf_phys(a)], vb = sA[lpf_phys(b)];
            while (a != LPF_NONE && b != LPF_NONE) {
                const u32 pa = lpf_phys(a), pb = lpf_phys(b);
                if (va > vb) {  // NSV(a) = b
                    sPn[pa] = (unsigned short)(b + 1);
                    sD[pa] = min(sD[pa], sU[pb]);
                    const u32 nx = sPp[pa];
                    a = nx ? nx - 1 : LPF_NONE;
                    if (a != LPF_NONE) va = sA[lpf_phys(a)];
                } else {        // PSV(b) = a
                    sPp[pb] = (unsigned short)(a + 1);
                    sU[pb] = min(sU[pb], sD[pa]);
                    const u32 nx = sPn[pb];
                    b = nx ? nx - 1 : LPF_NONE;
                    if (b != LPF_NONE) vb = sA[lpf_phys(b)];
                }
            }
            const u32 la = sNL[ca], lb = sNL[cb];
            while (a != LPF_NONE) {  // smaller than everything in B: still open to the right
                const u32 pa = lpf_phys(a);
                sD[pa] = min(sD[pa], lb);
                const u32 nx = sPp[pa];
                a = nx ? nx - 1 : LPF_NONE;
            }
            while (b != LPF_NONE) {  // smaller than everything in A: still open to the left
                const u32 pb = lpf_phys(b);
                sU[pb] = min(sU[pb], la);
                const u32 nx = sPn[pb];
                b = nx ? nx - 1 : LPF_NONE;
            }
            sNL[ca] = min(la, lb);
        }
        if (half < 16) __syncwarp();
    }
    __syncthreads();
    // ---- 3. the tile's own prefix / suffix minima continue in the global tree ----
    {
        u32 open_up = 0, open_dn = 0;
#pragma unroll 4
        for (u32 s = 0; s < CH; s++) {
            if (LPF_AT(sPp, s) == 0) open_up |= 1u << s;
            if (LPF_AT(sPn, s) == 0) open_dn |= 1u << s;
        }
        const u32 cnt = __popc(open_up) + __popc(open_dn);
        if (cnt) {
            u32 o = atomicAdd(sQn, cnt);
            while (open_up) {
                const u32 s = __ffs(int(open_up)) - 1;
                open_up &= open_up - 1;
                if (o < u32(2 * LPF_TILE / 16)) sQ[o] = (cs + s) << 1;
                o++;
            }
            while (open_dn) {
                const u32 s = __ffs(int(open_dn)) - 1;
                open_dn &= open_dn - 1;
                if (o < u32(2 * LPF_TILE / 16)) sQ[o] = ((cs + s) << 1) | 1u;
                o++;
            }
        }
    }
    __syncthreads();
    {
        const u32 last = min(base + u32(LPF_TILE), n) - 1u;  // last rank of this tile
        const u32 qn = *sQn;
        if (qn <= u32(2 * LPF_TILE / 16)) {
            for (u32 k = threadIdx.x; k < qn; k += LPF_THREADS) lpf_resolve_open<DIST>(T, base, last, thr, sQ[k] >> 1, sQ[k] & 1u, sA, sU, sD, D);
        } else {
            // pathological tile (e.g. monotone SA): more open ranks than the queue holds; every thread serves its own chunk
            for (u32 s = 0; s < CH; s++) {
                if (LPF_AT(sPp, s) == 0) lpf_resolve_open<DIST>(T, base, last, thr, cs + s, 0u, sA, sU, sD, D);
                if (LPF_AT(sPn, s) == 0) lpf_resolve_open<DIST>(T, base, last, thr, cs + s, 1u, sA, sU, sD, D);
            }
        }
    }
    __syncthreads();
    // ---- combine (PSV wins ties, LZSSLCPCompressor.hpp:101) ----
    for (u32 j = threadIdx.x; j < LPF_TILE; j += LPF_THREADS) {
        const u32 p = base + j;
        if (p >= n) break;
        const u32 lu = sU[lpf_phys(j)], ld = sD[lpf_phys(j)];
        if (DIST) {
            D.lu[p] = lu;
            D.ld[p] = ld;
            const u32 pp = sPp[lpf_phys(j)], pn = sPn[lpf_phys(j)];
            if (pp) D.su[p] = sA[lpf_phys(pp - 1)];  // open ranks were written by lpf_resolve_open
            if (pn) D.sd[p] = sA[lpf_phys(pn - 1)];
        } else {
            const u32 len = max(lu, ld);
            out_lenside[p] = len >= thr ? ((len << 1) | (lu >= ld ? 0u : 1u)) : 0u;
        }
    }
#undef LPF_AT
}

// ---------------------------------------------------------------------------------------------------------------
// greedy chain
// ---------------------------------------------------------------------------------------------------------------
#ifdef TDC_CUSIM
static const int CH_THREADS = 512;
static const int CH_IPT = 2;  // small tiles so that the CPU tests cross many tile/region boundaries
#else
#ifndef CH_THREADS_CFG
#define CH_THREADS_CFG 256
#endif
#ifndef CH_IPT_CFG
#define CH_IPT_CFG 8  // 256 x 8 = tiles of 2048 positions (round 1: 512 x 16): chain_mark 0.49 instead of 0.83 ms, chain_exit 1.17
                      // instead of 1.31 ms at dna 2^28 (profiles/r2y_variants.txt, r2z_variants.txt)
#endif
static const int CH_THREADS = CH_THREADS_CFG;
static const int CH_IPT = CH_IPT_CFG;
#endif
static const int CH_TILE = CH_THREADS * CH_IPT;  // text positions per tile
static_assert(CH_TILE % 1024 == 0, "emit_factors runs CH_TILE / 32 threads per tile and reduces over whole warps");
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
        // read phase / write phase with a barrier in between: every hop reads the pointers of the previous round (the
        // in-place version hopped through pointers other threads were rewriting — same fixed point, but a data race)
        u32 nxt[CH_IPT];
        bool any = false;
#pragma unroll
        for (int q = 0; q < CH_IPT; q++) {
            const u32 t = J[threadIdx.x + u32(q) * CH_THREADS];
            const bool inside = t < tile_end;  // still inside: hop through the target's pointer (a node on j's path)
            nxt[q] = inside ? J[t - base] : t;
            any |= inside;
        }
        __syncthreads();
        if (any) {
            changed = 1;
#pragma unroll
            for (int q = 0; q < CH_IPT; q++) J[threadIdx.x + u32(q) * CH_THREADS] = nxt[q];
        }
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
