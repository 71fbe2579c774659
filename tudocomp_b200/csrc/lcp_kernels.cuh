// Kernels for Phi / PLCP / LCP / BWT (see lcp.cu); shared by the single-GPU and the sharded multi-GPU driver.
#pragma once
#include "tdc_ctx.h"

namespace tdc {

// Phi[SA[j]] = SA[j-1] is emitted as (index, value) pairs for the partitioned scatter; BWT is a gather
static __global__ void __launch_bounds__(256)
phi_bwt_kernel(const u32* __restrict__ sa, const uint8_t* __restrict__ text, u64 n, u32* __restrict__ phi_idx,
               u32* __restrict__ phi_val, uint8_t* __restrict__ bwt) {
    const u64 j = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u32 s = sa[j];
    if (phi_idx) { phi_idx[j] = s; phi_val[j] = sa[j ? j - 1 : n - 1]; }
    if (bwt) bwt[j] = s ? text[s - 1] : text[n - 1];
}

static const u32 PLCP_IRRED = 0x80000000u;  // flag bit in the temporary array (values are < 2^31)
static const u32 PLCP_THREAD_LIMIT = 64;    // bytes compared by a single thread before the warp takes over

// longest common extension of suffixes a and b (a != b), starting from a known common length l0
__device__ __forceinline__ u32 lce_thread(const uint8_t* __restrict__ t, u64 a, u64 b, u32 l0, u32 limit, bool* done) {
    u32 l = l0;
    *done = true;
    while (true) {
        const u64 x = load_text8(t, a + l) ^ load_text8(t, b + l);
        if (x) return l + ((__ffsll((long long)x) - 1) >> 3);
        l += 8;
        if (l >= limit) { *done = false; return l; }
    }
}

// warp-cooperative continuation: lane q compares bytes [l + 8q, l + 8q + 8)
__device__ __forceinline__ u32 lce_warp(const uint8_t* __restrict__ t, u64 a, u64 b, u32 l0) {
    u32 l = l0;
    while (true) {
        const u64 off = u64(l) + 8u * lane_id();
        const u64 x = load_text8(t, a + off) ^ load_text8(t, b + off);
        const u32 bal = __ballot_sync(kFull, x != 0);
        if (bal) {
            const u32 first = __ffs(int(bal)) - 1;
            const u64 xf = __shfl_sync(kFull, x, int(first));
            return l + 8u * first + ((__ffsll((long long)xf) - 1) >> 3);
        }
        l += 256;
    }
}

// Pass 1: irreducible positions get their PLCP (flagged), reducible ones 0.  Long comparisons are queued.
static __global__ void __launch_bounds__(256)
plcp_irreducible_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ phi, u64 n, u32* __restrict__ tmp,
                        u32* __restrict__ queue, u32* __restrict__ queue_len) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i + 1 >= n) {
        if (i + 1 == n) tmp[i] = PLCP_IRRED;  // position n-1: handled by the fill pass (reference quirk)
        return;
    }
    const u32 ph = phi[i];
    const bool irreducible = (i == 0) || (ph == 0) || (text[i - 1] != text[ph - 1]);
    if (!irreducible) { tmp[i] = 0; return; }
    bool done;
    const u32 l = lce_thread(text, i, ph, 0, PLCP_THREAD_LIMIT, &done);
    tmp[i] = l | PLCP_IRRED;
    if (!done) queue[atomicAdd(queue_len, 1u)] = u32(i);
}

static __global__ void __launch_bounds__(256)
plcp_long_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ phi, u32* __restrict__ tmp,
                 const u32* __restrict__ queue, const u32* __restrict__ queue_len, u32* __restrict__ too_long) {
    const u32 nq = *queue_len;
    const u32 warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += warps) {
        const u32 i = queue[q];
        const u32 l0 = tmp[i] & ~PLCP_IRRED;
        const u32 l = lce_warp(text, i, phi[i], l0);
        __syncwarp();
        if (lane_id() == 0) {
            tmp[i] = l | PLCP_IRRED;
            if (l & PLCP_IRRED) *too_long = 1;  // a common prefix of 2^31 bytes or more (only possible beyond n = 2^31): reported by the host
        }
    }
}

static const int PF_THREADS = 256;
static const int PF_IPT = 8;
static const int PF_TILE = PF_THREADS * PF_IPT;

// Pass 2a: per tile, (index + 1) of the last irreducible position
static __global__ void __launch_bounds__(PF_THREADS)
plcp_fill_reduce_kernel(const u32* __restrict__ tmp, u64 n, u32* __restrict__ agg) {
    __shared__ u32 s_max[PF_THREADS / 32];
    const u64 t0 = u64(blockIdx.x) * PF_TILE + u64(threadIdx.x) * PF_IPT;
    u32 last = 0;
#pragma unroll
    for (int q = 0; q < PF_IPT; q++) {
        const u64 i = t0 + q;
        if (i < n && (tmp[i] & PLCP_IRRED)) last = u32(i) + 1u;
    }
    last = warp_max(last);
    if (lane_id() == 0) s_max[warp_id()] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 mx = 0;
        for (int w = 0; w < PF_THREADS / 32; w++) mx = max(mx, s_max[w]);
        agg[blockIdx.x] = mx;
    }
}

// Pass 2b: single CTA exclusive max-scan over the tile aggregates
static __global__ void __launch_bounds__(1024) plcp_fill_scan_kernel(u32* __restrict__ agg, u32 ntiles) {
    __shared__ u32 scratch[33];
    __shared__ u32 warp_last[32];
    u32 carry = 0;
    for (u32 b = 0; b < ntiles; b += 1024) {
        const u32 i = b + threadIdx.x;
        const u32 h = i < ntiles ? agg[i] : 0;
        u32 tot;
        const u32 in_m = block_inclusive_max(h, scratch, &tot);
        u32 ex_m = __shfl_up_sync(kFull, in_m, 1);
        if (lane_id() == 31) warp_last[warp_id()] = in_m;
        __syncthreads();
        if (lane_id() == 0) ex_m = warp_id() ? warp_last[warp_id() - 1] : 0u;
        __syncthreads();
        if (i < ntiles) agg[i] = max(carry, ex_m);
        carry = max(carry, tot);
    }
}

// Pass 2c: PLCP[i] = PLCP[i0] - (i - i0), i0 = last irreducible position <= i; max over i <= n-2
static __global__ void __launch_bounds__(PF_THREADS)
plcp_fill_apply_kernel(const u32* __restrict__ tmp, const u32* __restrict__ phi, u64 n, const u32* __restrict__ pre,
                       u32* __restrict__ plcp, u32* __restrict__ max_lcp) {
    __shared__ u32 scratch[33];
    __shared__ u32 warp_last[32];
    __shared__ u32 s_max[PF_THREADS / 32];
    const u64 t0 = u64(blockIdx.x) * PF_TILE + u64(threadIdx.x) * PF_IPT;
    u32 v[PF_IPT];
    u32 last = 0;
#pragma unroll
    for (int q = 0; q < PF_IPT; q++) {
        const u64 i = t0 + q;
        v[q] = i < n ? tmp[i] : 0u;
        if (i < n && (v[q] & PLCP_IRRED)) last = u32(i) + 1u;
    }
    u32 tot;
    const u32 in_m = block_inclusive_max(last, scratch, &tot);
    u32 ex_m = __shfl_up_sync(kFull, in_m, 1);
    if (lane_id() == 31) warp_last[warp_id()] = in_m;
    __syncthreads();
    if (lane_id() == 0) ex_m = warp_id() ? warp_last[warp_id() - 1] : 0u;
    u32 cur = max(pre[blockIdx.x], ex_m);  // (index + 1) of the irreducible position in force
    u32 cur_val = cur ? (tmp[cur - 1] & ~PLCP_IRRED) : 0u;
    u32 mx = 0;
#pragma unroll
    for (int q = 0; q < PF_IPT; q++) {
        const u64 i = t0 + q;
        if (i >= n) break;
        if (v[q] & PLCP_IRRED) { cur = u32(i) + 1u; cur_val = v[q] & ~PLCP_IRRED; }
        u32 out;
        if (i + 1 == n) {
            out = phi[i];  // PLCPFromPhi.hpp:38 — the loop stops at n-2, the slot keeps Phi[n-1]
        } else {
            out = cur_val - (u32(i) - (cur - 1u));
            mx = max(mx, out);
        }
        plcp[i] = out;
    }
    mx = warp_max(mx);
    if (lane_id() == 0) s_max[warp_id()] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 m2 = 0;
        for (int w = 0; w < PF_THREADS / 32; w++) m2 = max(m2, s_max[w]);
        if (m2) atomicMax(max_lcp, m2);
    }
}

static __global__ void __launch_bounds__(256)
lcp_gather_kernel(const u32* __restrict__ sa, const u32* __restrict__ plcp, u64 n, u32* __restrict__ lcp) {
    const u64 j = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    lcp[j] = j ? plcp[sa[j]] : 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// Direct route (LCP only, texts whose suffixes separate early): LCP[j] = lce(SA[j-1], SA[j]) by character comparison
// in SA order.  Each lane loads the first 32 bytes of its own suffix once and receives its predecessor's from the
// neighbouring lane, so the common case costs one random text access per suffix instead of the Phi scatter, the PLCP
// pass and the LCP gather of the reference's data flow.  Identical values by definition (LCPFromPLCP.hpp:43-47).
// ---------------------------------------------------------------------------------------------------------------
static const u32 LCPD_THREAD_LIMIT = 256;  // bytes compared by one thread before the pair is queued for a whole warp

static __global__ void __launch_bounds__(256)
lcp_direct_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ sa, u64 n, u32* __restrict__ lcp,
                  u32* __restrict__ queue, u32* __restrict__ queue_len, u32* __restrict__ max_lcp) {
    __shared__ u32 s_max[256 / 32];
    const u64 j = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool valid = j < n;
    const u32 own = valid ? sa[j] : 0u;
    u64 w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = valid ? load_text8(text, u64(own) + 8u * k) : 0ull;
    u32 prev = __shfl_up_sync(kFull, own, 1);
    u64 pw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) pw[k] = __shfl_up_sync(kFull, w[k], 1);
    if (lane_id() == 0 && valid && j > 0) {
        prev = sa[j - 1];
#pragma unroll
        for (int k = 0; k < 4; k++) pw[k] = load_text8(text, u64(prev) + 8u * k);
    }
    u32 l = 0;
    if (valid && j > 0) {
        bool found = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!found) {
                const u64 x = w[k] ^ pw[k];
                if (x) { l = 8u * k + ((__ffsll((long long)x) - 1) >> 3); found = true; }
            }
        }
        if (!found) {
            bool done;
            l = lce_thread(text, own, prev, 32, LCPD_THREAD_LIMIT, &done);
            if (!done) queue[atomicAdd(queue_len, 1u)] = u32(j);
        }
        lcp[j] = l;
    } else if (valid) {
        lcp[0] = 0;
    }
    u32 mx = warp_max(l);
    if (lane_id() == 0) s_max[warp_id()] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < 256 / 32; q++) mx = max(mx, s_max[q]);
        if (mx) atomicMax(max_lcp, mx);
    }
}

// Seeded variant: the initial sort already wrote the LCP of every pair its keys separate (suffix_array.cu,
// key_common_symbols); only pairs marked LCP_UNKNOWN — same k-symbol prefix — are compared, starting at offset k.
static const int LCPFIX_EPT = 4;  // slots per thread: one 16-byte load of the seeded values (a 4-byte load per thread ran at 1 TB/s)
static __global__ void __launch_bounds__(256)
lcp_fix_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ sa, u64 n, u32* __restrict__ lcp, u32 l0,
               u32* __restrict__ queue, u32* __restrict__ queue_len, u32* __restrict__ max_lcp) {
    __shared__ u32 s_max[256 / 32];
    const u64 j0 = (u64(blockIdx.x) * blockDim.x + threadIdx.x) * LCPFIX_EPT;
    u32 lv[LCPFIX_EPT];
    if (j0 + LCPFIX_EPT <= n) {
        const uint4 q = *reinterpret_cast<const uint4*>(lcp + j0);
        lv[0] = q.x; lv[1] = q.y; lv[2] = q.z; lv[3] = q.w;
    } else {
#pragma unroll
        for (int q = 0; q < LCPFIX_EPT; q++) lv[q] = j0 + q < n ? lcp[j0 + q] : 0u;
    }
    u32 mx = 0;
#pragma unroll
    for (int q = 0; q < LCPFIX_EPT; q++) {
        const u64 j = j0 + q;
        u32 l = lv[q];
        if (j < n && l == LCP_UNKNOWN) {  // j >= 1: slot 0 is never unknown
            const u32 own = sa[j], prev = sa[j - 1];
            // (keys without a length field: a suffix shorter than l0 shares its padded key, not l0 real symbols)
            const u32 start = min(l0, min(u32(n - 1) - own, u32(n - 1) - prev));
            bool done;
            l = lce_thread(text, own, prev, start, start + LCPD_THREAD_LIMIT, &done);
            if (!done) queue[atomicAdd(queue_len, 1u)] = u32(j);
            lcp[j] = l;
        }
        mx = max(mx, l);
    }
    mx = warp_max(mx);
    if (lane_id() == 0) s_max[warp_id()] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < 256 / 32; q++) mx = max(mx, s_max[q]);
        if (mx) atomicMax(max_lcp, mx);
    }
}

// Packed initial keys without a length field (suffix_array.cu): the `span` suffixes that reach the sentinel inside their
// key were padded with code 0, so the key-derived LCP of a slot next to one of them may count padding.  Recompute the
// two slots around each of them by direct comparison (they are shorter than a key: a few bytes each).
static __global__ void lcp_tail_fix_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ sa, const u32* __restrict__ isa,
                                           u64 n, u32 span, u32* __restrict__ lcp) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > span || u64(j) >= n) return;
    const u64 p = isa[n - 1 - j];
    for (u64 s = p; s <= p + 1 && s < n; s++) {
        if (s == 0) continue;
        const u64 a = sa[s - 1], b = sa[s];
        u32 l = 0;
        while (text[a + l] == text[b + l]) l++;  // two different suffixes: the unique 0 at n-1 ends the loop
        lcp[s] = l;
    }
}

static __global__ void __launch_bounds__(256)
lcp_direct_long_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ sa, u32* __restrict__ lcp,
                       const u32* __restrict__ queue, const u32* __restrict__ queue_len, u32* __restrict__ max_lcp) {
    const u32 nq = *queue_len;
    const u32 warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += warps) {
        const u32 j = queue[q];
        const u32 l = lce_warp(text, sa[j], sa[j - 1], lcp[j]);
        __syncwarp();
        if (lane_id() == 0) { lcp[j] = l; atomicMax(max_lcp, l); }
    }
}

}  // namespace tdc
