// Kernels that exist only in the sharded multi-GPU text index (dist_textds.cu): bucket partition for the exchanges,
// request/reply gathers, the LPF variant that carries source positions and queues walks leaving the shard, and the
// resolution of such walks on the neighbouring shards.
#pragma once
#include "lcp_kernels.cuh"
#include "lzss_kernels.cuh"
#include "sa_kernels.cuh"

namespace tdc {

static const int DIST_MAX_RANKS = 16;

// ---------------------------------------------------------------------------------------------------------------
// bucket functors: which rank does an element go to, and what is sent
// ---------------------------------------------------------------------------------------------------------------
struct SplitterFn {  // initial sort: bucket = number of splitters <= key (equal keys share a bucket)
    u64 spl[DIST_MAX_RANKS - 1];
    int nspl;
    __device__ __forceinline__ u32 bucket(u64 key) const {
        u32 b = 0;
#pragma unroll
        for (int i = 0; i < DIST_MAX_RANKS - 1; i++)
            if (i < nspl && spl[i] <= key) b++;
        return b;
    }
    __device__ __forceinline__ u64 out(u64 key, u32) const { return key; }
};
struct OwnerFn {  // position-sharded arrays: owner = position / block, the owner receives its local index
    u32 block;
    __device__ __forceinline__ u32 bucket(u32 pos) const { return pos / block; }
    __device__ __forceinline__ u32 out(u32 pos, u32 b) const { return pos - b * block; }
};

// lanes of the warp that hold the same bucket (< 16) and the same validity: 5 ballots instead of one per bucket
__device__ __forceinline__ u32 bucket_peers(u32 b, bool valid) {
    u32 peers = kFull;
#pragma unroll
    for (int bit = 0; bit < 4; bit++) {
        const u32 bal = __ballot_sync(kFull, (b >> bit) & 1u);
        peers &= ((b >> bit) & 1u) ? bal : ~bal;
    }
    const u32 balv = __ballot_sync(kFull, valid);
    return peers & (valid ? balv : ~balv);
}

static const int BP_THREADS = 256;
static const int BP_IPT = 8;
static const int BP_TILE = BP_THREADS * BP_IPT;

template <class K, class F>
static __global__ void __launch_bounds__(BP_THREADS)
bucket_count_kernel(const K* __restrict__ keys, u64 m, F f, int nbuckets, ull* __restrict__ gcount) {
    __shared__ u32 cnt[DIST_MAX_RANKS];
    if (threadIdx.x < DIST_MAX_RANKS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 m_round = (m + 31) & ~u64(31);
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < m_round; i += u64(gridDim.x) * blockDim.x) {
        const bool valid = i < m;
        const u32 b = valid ? f.bucket(keys[i]) : 0u;
        const u32 peers = bucket_peers(b, valid);  // one shared atomic per distinct bucket of the warp
        if (valid && lane_id() == u32(__ffs(int(peers)) - 1)) atomicAdd(&cnt[b], u32(__popc(peers)));
    }
    __syncthreads();
    if (threadIdx.x < u32(nbuckets) && cnt[threadIdx.x]) atomicAdd(&gcount[threadIdx.x], ull(cnt[threadIdx.x]));
}

// Where the elements of each bucket go.  With the peer-memory transport k[b] (and v[b], v2[b] when the values travel
// too) point INTO RANK b's receive buffer, at the place reserved for this rank: the partition kernel's stores are the
// all-to-all (NVLink writes), there is no send buffer and no separate exchange.  Otherwise they point at the bucket's
// segment of a local staging buffer.  cursor[b] counts from 0; order inside a bucket is arbitrary.
struct BucketDst {
    void* k[DIST_MAX_RANKS];
    u32* v[DIST_MAX_RANKS];
    u32* v2[DIST_MAX_RANKS];
};

// vals == nullptr: the value of element i is vbase + i.  vals2: optional second value travelling with the first.
// The tile is staged in shared memory in bucket order and written out run by run, so that a warp's stores to a peer are
// whole consecutive lines (with one store per thread straight from registers, a warp hit all P buckets at once and the
// NVLink writes were 32-byte fragments: the push was 40 % of the 8-GPU step, profiles/r1m_summary.md).
template <class K, class F>
static __global__ void __launch_bounds__(BP_THREADS)
bucket_scatter_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, u32 vbase, u64 m, F f, int nbuckets,
                      ull* __restrict__ cursor, BucketDst dst, const u32* __restrict__ vals2) {
    __shared__ u32 cnt[DIST_MAX_RANKS];
    __shared__ u32 start[DIST_MAX_RANKS + 1];
    __shared__ ull gbase[DIST_MAX_RANKS];
    __shared__ K* kb[DIST_MAX_RANKS];
    __shared__ u32* vb[DIST_MAX_RANKS];
    __shared__ u32* v2b[DIST_MAX_RANKS];
    __shared__ K s_key[BP_TILE];
    __shared__ u32 s_val[BP_TILE];
    __shared__ u32 s_val2[BP_TILE];
    if (threadIdx.x < DIST_MAX_RANKS) {
        cnt[threadIdx.x] = 0;
        kb[threadIdx.x] = static_cast<K*>(dst.k[threadIdx.x]);
        vb[threadIdx.x] = dst.v[threadIdx.x];
        v2b[threadIdx.x] = dst.v2[threadIdx.x];
    }
    __syncthreads();
    const u64 t0 = u64(blockIdx.x) * BP_TILE;
    K key[BP_IPT];
    u32 b[BP_IPT], lr[BP_IPT];
#pragma unroll
    for (int q = 0; q < BP_IPT; q++) {
        const u64 i = t0 + u64(q) * BP_THREADS + threadIdx.x;
        const bool valid = i < m;
        key[q] = valid ? keys[i] : K(0);
        b[q] = valid ? f.bucket(key[q]) : 0u;
        // rank inside the tile's bucket: the first lane of every group of equal buckets reserves the group's slots
        const u32 peers = bucket_peers(b[q], valid);
        const u32 leader = u32(__ffs(int(peers)) - 1);
        u32 base = 0;
        if (valid && lane_id() == leader) base = atomicAdd(&cnt[b[q]], u32(__popc(peers)));
        base = __shfl_sync(kFull, base, leader);
        lr[q] = base + __popc(peers & lanemask_lt());
    }
    __syncthreads();
    if (threadIdx.x < u32(nbuckets)) gbase[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], ull(cnt[threadIdx.x])) : 0;
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (int bb = 0; bb < DIST_MAX_RANKS; bb++) {
            start[bb] = run;
            if (bb < nbuckets) run += cnt[bb];
        }
        start[DIST_MAX_RANKS] = run;
    }
    __syncthreads();
    // ---- stage in bucket order ----
#pragma unroll
    for (int q = 0; q < BP_IPT; q++) {
        const u64 i = t0 + u64(q) * BP_THREADS + threadIdx.x;
        if (i < m) {
            const u32 p = start[b[q]] + lr[q];
            s_key[p] = f.out(key[q], b[q]);
            s_val[p] = vals ? vals[i] : vbase + u32(i);
            if (vals2) s_val2[p] = vals2[i];
        }
    }
    __syncthreads();
    // ---- runs out: consecutive threads, consecutive addresses of the same bucket ----
    const u32 total = start[DIST_MAX_RANKS];
    for (u32 j = threadIdx.x; j < total; j += BP_THREADS) {
        u32 bj = 0;
#pragma unroll
        for (int x = 1; x < DIST_MAX_RANKS; x++)
            if (x < nbuckets && start[x] <= j) bj = u32(x);  // starts are non-decreasing: the last one not above j
        const u64 o = gbase[bj] + (j - start[bj]);
        kb[bj][o] = s_key[j];
        vb[bj][o] = s_val[j];
        if (vals2) v2b[bj][o] = s_val2[j];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// small element-wise helpers
// ---------------------------------------------------------------------------------------------------------------
static __global__ void add_offset_kernel(const u32* __restrict__ in, u64 m, u32 off, u32* __restrict__ out) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < m) out[i] = in[i] + off;
}
static __global__ void gather_u32_kernel(const u32* __restrict__ table, const u32* __restrict__ idx, u64 m, u32* __restrict__ out) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < m) out[i] = table[idx[i]];
}
static __global__ void sample_keys_kernel(const u64* __restrict__ keys, u64 m, u32 samples, u64* __restrict__ out) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < samples) out[j] = m ? keys[(u64(j) * m) / samples] : ~u64(0);
}
// doubling key from an already gathered second rank
static __global__ void build_keys_from_kernel(const u32* __restrict__ gid, const u32* __restrict__ r2, u64 m, u32 rbits,
                                              u64* __restrict__ keys) {
    const u64 o = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o < m) keys[o] = (u64(gid[o]) << rbits) | r2[o];
}
// LCP of the first slot of a shard against the last suffix of the previous shard (one warp)
static __global__ void lcp_boundary_kernel(const uint8_t* __restrict__ text, const u32* __restrict__ sa, u32 prev_sa,
                                           u32* __restrict__ lcp, u32* __restrict__ max_lcp) {
    const u32 l = lce_warp(text, sa[0], prev_sa, 0);
    if (lane_id() == 0) { lcp[0] = l; atomicMax(max_lcp, l); }
}

// ---------------------------------------------------------------------------------------------------------------
// LPF per rank with source positions; walks that leave the shard are queued for the neighbouring shards
// ---------------------------------------------------------------------------------------------------------------
struct WalkAnswer {
    u32 p, m, src;
};

// Queries arriving from a neighbouring shard, answered against this shard's tree.  UP: the walk enters at the shard's
// last slot and moves towards slot 0; otherwise it enters at slot 0 and moves up.  Found -> answer for the origin;
// not found -> forwarded (with the minimum over this whole shard folded in); below the threshold -> dropped.
template <bool UP>
static __global__ void __launch_bounds__(128)
resolve_queries_kernel(MinTree T, u32 n, u32 thr, const WalkQuery* __restrict__ in, u32 nq, WalkAnswer* __restrict__ ans,
                       WalkQuery* __restrict__ fwd, u32* __restrict__ cnt /*[0] answers, [1] forwards*/) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    WalkQuery w = in[t];
    u32 q = 0;
    int r = WALK_OFF_TREE;
    if (n > 0) {
        if (UP) {
            r = walk_psv(T, n, w.v, thr, w.m, q);
        } else {
            w.m = min(w.m, T.l[0][0]);
            if (w.m < thr) r = WALK_ABANDONED;
            else if (T.a[0][0] < w.v) { r = WALK_FOUND; q = 0; }
            else r = walk_nsv(T, 0, w.v, thr, w.m, q);
        }
    }
    if (r == WALK_FOUND) {
        WalkAnswer a;
        a.p = w.p;
        a.m = w.m;
        a.src = T.a[0][q];
        ans[atomicAdd(&cnt[0], 1u)] = a;
    } else if (r == WALK_OFF_TREE) {
        fwd[atomicAdd(&cnt[1], 1u)] = w;
    }
}

static __global__ void apply_answers_kernel(const WalkAnswer* __restrict__ ans, u32 na, u32* __restrict__ l_out, u32* __restrict__ s_out) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na) return;
    const WalkAnswer a = ans[t];
    l_out[a.p] = a.m;
    s_out[a.p] = a.src;
}

// (l_up, src_up, l_dn, src_dn) -> (len << 1 | side, src); PSV wins ties (LZSSLCPCompressor.hpp:101)
static __global__ void lpf_combine_kernel(const u32* __restrict__ lu, const u32* __restrict__ su, const u32* __restrict__ ld,
                                          const u32* __restrict__ sd, u64 m, u32 thr, u32* __restrict__ lenside, u32* __restrict__ src) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u32 a = lu[i], b = ld[i];
    const u32 len = max(a, b);
    const bool up = a >= b;
    lenside[i] = len >= thr ? ((len << 1) | (up ? 0u : 1u)) : 0u;
    src[i] = up ? su[i] : sd[i];
}

}  // namespace tdc
