// Sharded multi-GPU text index + lzss_lcp factoriser: ONE text, P ranks (one process and one GPU each).
// No reference counterpart (tudocomp is single-threaded); results are the same SA / ISA / LCP / factor list the
// single-GPU path and the reference produce (SADivSufSort.hpp:28-51, ISAFromSA.hpp:37-39, LCPFromPLCP.hpp:43-47,
// LZSSLCPCompressor.hpp:60-115), cut into shards.
//
// Layout.  The text (n bytes) is replicated on every rank.  Position-indexed arrays (the rank array = ISA, the
// text-order LPF table) are sharded by blocks of `block` positions: rank r owns positions [r*block, (r+1)*block).
// Slot-indexed arrays (SA, LCP) are sharded by the buckets of the initial sort: rank r owns the slots
// [slot_lo(r), slot_lo(r) + slot_cnt(r)), the ranks in increasing key order.
//
// Suffix array by prefix doubling:
//   round 0  every rank packs the k-symbol keys of its positions, P-1 splitters come from an all-gathered regular
//            sample, the (key, suffix) pairs are partitioned by splitter and exchanged (ALL-TO-ALL #1: 12 B per
//            suffix), each rank radix-sorts its bucket.  Equal keys share a bucket, so a group never spans ranks and
//            every later sort is rank-local.  The new ranks (suffix -> head slot) go to the owners of the positions
//            (ALL-TO-ALL #2: 8 B per suffix) and are applied with the partitioned scatter.
//   round r  for the suffixes still in groups: rank[suffix + h] is fetched from its owner (request/reply ALL-TO-ALLs,
//            4 + 4 B per active suffix), local segmented sort by (group, rank[suffix + h]), re-rank, rank updates to
//            their owners.  Stops when no rank has a group left (all-gathered count).
// LCP: compared directly on the replicated text in SA order (seeded by the initial keys); the first slot of a shard is
//      compared with the previous shard's last suffix.
// Factorisation: PSV/NSV + range minima per slot on the local min-tree; walks that leave the shard are queued and
//      travel shard by shard (neighbour exchange, at most P-1 hops; typically a few dozen queries per shard);
//      (len|side, src) go to text order through the position owners; the greedy chain is walked rank after rank
//      (one 8-byte hand-over per rank); factors are emitted per rank, already in position order.
#include <algorithm>
#include <new>

#include "../../include/tdcgpu.h"
#include "dist_comm.h"
#include <cstdlib>

#include "dist_kernels.cuh"

namespace tdc {

struct DistCtx {
    Ctx c;  // device, stream, replicated text, scalars, sort workspace, scratch arena, phase times
    Comm* comm = nullptr;
    int P = 1, rank = 0;
    u64 n = 0;
    u32 block = 0;
    u64 pos_lo = 0, pos_cnt = 0, slot_lo = 0, slot_cnt = 0;
    std::vector<u64> slot_cnts;
    u64 cap = 0, qcap = 0;
    u32 *d_sa = nullptr, *d_rank = nullptr, *d_lcp = nullptr;
    u32 have = 0;
    ull* d_counts = nullptr;  // [2 * DIST_MAX_RANKS] bucket counts / cursors
    ull* h_counts = nullptr;  // pinned, [64]
    u64 total_factors = 0;
    std::vector<u64> xchg;  // scratch for count matrices
    // peer-memory transport: the scratch arena of every rank mapped into this process (nullptr: NCCL send/recv instead)
    bool p2p = false;
    void* peer_arena[DIST_MAX_RANKS] = {nullptr};
    cudaStream_t copy_streams[DIST_MAX_RANKS] = {};
    cudaEvent_t copy_events[DIST_MAX_RANKS + 1] = {};
    u64 p2p_bytes = 0, nccl_bytes = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// exchange helpers
// ---------------------------------------------------------------------------------------------------------------
// matrix[r * P + p] = number of elements rank r sends to rank p
static int exchange_matrix(DistCtx& d, const u64* send_cnt, std::vector<u64>& matrix) {
    matrix.assign(size_t(d.P) * d.P, 0);
    return d.comm->allgather_host(send_cnt, matrix.data(), sizeof(u64) * d.P);
}

// All-to-all of n_arrays arrays that share one count matrix (matrix[r * P + p] = elements rank r sends to rank p): the
// elements for peer p start at sum_{q<p} matrix[rank][q] in every send array and arrive at sum_{r<rank} matrix[r][p] in
// p's receive array.
//   peer-memory transport: the receive buffers live in the scratch arenas, which every rank has mapped; the ranks tell
//   each other the arena offsets of their receive buffers, then one device-to-device copy per peer and array is pushed
//   over NVLink on P-1 side streams; stream sync + host barrier (the data of all peers has landed).
//   otherwise: grouped ncclSend/ncclRecv.
static int a2a_multi(DistCtx& d, int n_arrays, const void* const* sends, void* const* recvs, const size_t* esz, const u64* matrix) {
    const int P = d.P, rank = d.rank;
    u64 scnt[DIST_MAX_RANKS], rcnt[DIST_MAX_RANKS], soff[DIST_MAX_RANKS], roff[DIST_MAX_RANKS], remote_off[DIST_MAX_RANKS];
    u64 so = 0, ro = 0, sent = 0;
    for (int p = 0; p < P; p++) {
        scnt[p] = matrix[size_t(rank) * P + p];
        rcnt[p] = matrix[size_t(p) * P + rank];
        soff[p] = so; roff[p] = ro;
        so += scnt[p]; ro += rcnt[p];
        remote_off[p] = 0;
        for (int r = 0; r < rank; r++) remote_off[p] += matrix[size_t(r) * P + p];
        if (p != rank) sent += scnt[p];
    }
    cudaStream_t st = d.c.stream;
    if (d.p2p) {
#ifndef TDC_CUSIM
        LaunchScope timed("p2p_alltoallv", st, false);
#endif
        // Which of its buffers a rank receives into is rank-specific (the ping-pong side a radix sort ends on depends on
        // the passes it could skip), so the receive offsets inside the arenas are exchanged; this all-gather is also the
        // "ready to receive" barrier: every rank has finished reading what is about to be overwritten.
        u64 my_off[4] = {0, 0, 0, 0};
        if (n_arrays > 4) { set_error("a2a_multi: too many arrays"); return TDCGPU_ERR_INTERNAL; }
        for (int a = 0; a < n_arrays; a++) my_off[a] = u64(static_cast<const uint8_t*>(recvs[a]) - d.c.arena.base);
        std::vector<u64> peer_off(size_t(P) * 4);
        TDC_TRY(d.comm->allgather_host(my_off, peer_off.data(), sizeof(my_off)));
        u64 bytes = 0;
        TDC_CUDA(cudaEventRecord(d.copy_events[DIST_MAX_RANKS], st));
        for (int p = 0; p < P; p++) {
            if (scnt[p] == 0) continue;
            cudaStream_t cs = p == rank ? st : d.copy_streams[p];
            if (p != rank) TDC_CUDA(cudaStreamWaitEvent(cs, d.copy_events[DIST_MAX_RANKS], 0));
            for (int a = 0; a < n_arrays; a++) {
                uint8_t* dst = static_cast<uint8_t*>(d.peer_arena[p]) + peer_off[size_t(p) * 4 + a] + remote_off[p] * esz[a];
                const uint8_t* src = static_cast<const uint8_t*>(sends[a]) + soff[p] * esz[a];
                TDC_CUDA(cudaMemcpyAsync(dst, src, scnt[p] * esz[a], cudaMemcpyDeviceToDevice, cs));
                if (p != rank) bytes += scnt[p] * esz[a];
            }
            if (p != rank) {
                TDC_CUDA(cudaEventRecord(d.copy_events[p], cs));
                TDC_CUDA(cudaStreamWaitEvent(st, d.copy_events[p], 0));
            }
        }
        prof_add_bytes("p2p_alltoallv", double(bytes));
        d.p2p_bytes += bytes;
        TDC_CUDA(cudaStreamSynchronize(st));
        u64 token = 1;
        std::vector<u64> all(P);
        return d.comm->allgather_host(&token, all.data(), sizeof(u64));  // barrier: every rank's pushes are complete
    }
    for (int a = 0; a < n_arrays; a++) {
        u64 sb[DIST_MAX_RANKS], rb[DIST_MAX_RANKS], sob[DIST_MAX_RANKS], rob[DIST_MAX_RANKS];
        for (int p = 0; p < P; p++) { sb[p] = scnt[p] * esz[a]; rb[p] = rcnt[p] * esz[a]; sob[p] = soff[p] * esz[a]; rob[p] = roff[p] * esz[a]; }
#ifndef TDC_CUSIM
        LaunchScope timed("nccl_alltoallv", st, false);  // shows up in the per-kernel profile, not in the launch count
#endif
        prof_add_bytes("nccl_alltoallv", double(sent) * double(esz[a]));
        d.nccl_bytes += sent * esz[a];
        TDC_TRY(d.comm->alltoallv(sends[a], sob, sb, recvs[a], rob, rb, st));
    }
    return 0;
}

static int a2a_elems(DistCtx& d, const void* send, void* recv, size_t esz, const u64* matrix) {
    return a2a_multi(d, 1, &send, &recv, &esz, matrix);
}

// one value per rank
static int allgather_u64(DistCtx& d, const u64* mine, int words, std::vector<u64>& all) {
    all.assign(size_t(d.P) * words, 0);
    return d.comm->allgather_host(mine, all.data(), sizeof(u64) * words);
}

// Partition m (key, value[, value2]) elements by destination rank and deliver them.
//   kin / vin / vin2     the elements (vin == nullptr: value i = vbase + i)
//   kstage / vstage / v2stage   local staging, one segment per bucket (bucket order).  When send_vals is false the first
//                        value does not travel: it stays in vstage, in the order in which the keys were sent.
//   krecv / vrecv / v2recv      receive buffers (in the scratch arena), filled in source-rank order
//   matrix               out: matrix[r * P + p] = elements rank r sent to rank p
// Peer-memory transport: the count matrix and the receive-buffer offsets are all-gathered first, then ONE kernel does the
// partition and the all-to-all: bucket b's elements are stored straight into rank b's receive buffer over NVLink (the
// staging buffers are not touched for what travels).  Otherwise: partition into the staging buffers, then NCCL.
template <class K, class F>
static int exchange_by_bucket(DistCtx& d, const K* kin, const u32* vin, u32 vbase, const u32* vin2, u64 m, F f, K* kstage, u32* vstage,
                              u32* v2stage, K* krecv, u32* vrecv, u32* v2recv, bool send_vals, std::vector<u64>& matrix, u64* R_out) {
    cudaStream_t st = d.c.stream;
    const int P = d.P, rank = d.rank;
    TDC_CUDA(cudaMemsetAsync(d.d_counts, 0, sizeof(ull) * 2 * DIST_MAX_RANKS, st));
    if (m) {
        const u32 grid = u32(std::min<u64>(u64(d.c.sm_count) * 8, div_up(m, BP_THREADS)));
        auto bucket_count = bucket_count_kernel<K, F>;
        TDC_LAUNCH(bucket_count, grid, BP_THREADS, 0, st, kin, m, f, P, d.d_counts);
        TDC_KCHECK();
    }
    TDC_CUDA(cudaMemcpyAsync(d.h_counts, d.d_counts, sizeof(ull) * P, cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    // counts + where my receive buffers are (arena offsets); this all-gather is also the "ready to receive" barrier
    const int W = DIST_MAX_RANKS + 3;
    u64 mine[DIST_MAX_RANKS + 3] = {0};
    for (int p = 0; p < P; p++) mine[p] = d.h_counts[p];
    const uint8_t* ab = d.c.arena.base;
    mine[DIST_MAX_RANKS + 0] = u64(reinterpret_cast<const uint8_t*>(krecv) - ab);
    mine[DIST_MAX_RANKS + 1] = vrecv ? u64(reinterpret_cast<const uint8_t*>(vrecv) - ab) : 0;
    mine[DIST_MAX_RANKS + 2] = v2recv ? u64(reinterpret_cast<const uint8_t*>(v2recv) - ab) : 0;
    std::vector<u64> all(size_t(P) * W);
    TDC_TRY(d.comm->allgather_host(mine, all.data(), sizeof(u64) * W));
    matrix.assign(size_t(P) * P, 0);
    for (int r = 0; r < P; r++)
        for (int p = 0; p < P; p++) matrix[size_t(r) * P + p] = all[size_t(r) * W + p];
    u64 worst = 0;
    for (int p = 0; p < P; p++) {
        u64 R = 0;
        for (int r = 0; r < P; r++) R += matrix[size_t(r) * P + p];
        worst = std::max(worst, R);
        if (p == rank) *R_out = R;
    }
    if (worst > d.cap) {  // every rank sees the same matrix and takes the same decision
        set_error("exchange: a rank would receive %llu elements, capacity %llu (skewed keys)", (unsigned long long)worst, (unsigned long long)d.cap);
        return TDCGPU_ERR_NOMEM;
    }
    BucketDst dst;
    u64 start = 0, sent = 0;
    for (int p = 0; p < DIST_MAX_RANKS; p++) { dst.k[p] = nullptr; dst.v[p] = nullptr; dst.v2[p] = nullptr; }
    for (int p = 0; p < P; p++) {
        u64 remote_off = 0;
        for (int r = 0; r < rank; r++) remote_off += matrix[size_t(r) * P + p];
        if (d.p2p) {
            uint8_t* pa = static_cast<uint8_t*>(d.peer_arena[p]);
            dst.k[p] = pa + all[size_t(p) * W + DIST_MAX_RANKS + 0] + remote_off * sizeof(K);
            dst.v[p] = send_vals ? reinterpret_cast<u32*>(pa + all[size_t(p) * W + DIST_MAX_RANKS + 1]) + remote_off : vstage + start;
            dst.v2[p] = vin2 ? reinterpret_cast<u32*>(pa + all[size_t(p) * W + DIST_MAX_RANKS + 2]) + remote_off : nullptr;
        } else {
            dst.k[p] = kstage + start;
            dst.v[p] = vstage + start;
            dst.v2[p] = vin2 ? v2stage + start : nullptr;
        }
        start += mine[p];
        if (p != rank) sent += mine[p];
    }
    if (m) {
        if (d.p2p) {  // same kernel; the name tells the profile that its stores are the exchange
            auto bucket_scatter_push = bucket_scatter_kernel<K, F>;
            TDC_LAUNCH(bucket_scatter_push, u32(div_up(m, BP_TILE)), BP_THREADS, 0, st, kin, vin, vbase, m, f, P, d.d_counts + DIST_MAX_RANKS, dst, vin2);
            prof_add_bytes("bucket_scatter_push", double(m) * 2 * (sizeof(K) + 4 + (vin2 ? 4 : 0)));
        } else {
            auto bucket_scatter = bucket_scatter_kernel<K, F>;
            TDC_LAUNCH(bucket_scatter, u32(div_up(m, BP_TILE)), BP_THREADS, 0, st, kin, vin, vbase, m, f, P, d.d_counts + DIST_MAX_RANKS, dst, vin2);
            prof_add_bytes("bucket_scatter", double(m) * 2 * (sizeof(K) + 4 + (vin2 ? 4 : 0)));
        }
    }
    TDC_KCHECK();
    const size_t per_elem = sizeof(K) + (send_vals ? 4 : 0) + (vin2 ? 4 : 0);
    if (d.p2p) {
        d.p2p_bytes += sent * per_elem;
        TDC_CUDA(cudaStreamSynchronize(st));
        u64 token = 1;
        std::vector<u64> bar(P);
        return d.comm->allgather_host(&token, bar.data(), sizeof(u64));  // every rank's stores have landed
    }
    const void* sends[3] = {kstage, nullptr, nullptr};
    void* recvs[3] = {krecv, nullptr, nullptr};
    size_t esz[3] = {sizeof(K), 4, 4};
    int na = 1;
    if (send_vals) { sends[na] = vstage; recvs[na] = vrecv; na++; }
    if (vin2) { sends[na] = v2stage; recvs[na] = v2recv; na++; }
    return a2a_multi(d, na, sends, recvs, esz, matrix.data());
}

// dst_shard[idx - owner*block] = val on the owner of idx (and dst2 / val2 alike when given: one partition and one
// exchange of the indices serve both).  bufs: six u32 buffers of d.cap elements, eight with a second value.
static int dist_scatter(DistCtx& d, const u32* idx, const u32* val, u64 m, u32* dst_shard, u32** bufs, const u32* val2 = nullptr,
                        u32* dst2 = nullptr) {
    OwnerFn f;
    f.block = d.block;
    u64 R = 0;
    TDC_TRY((exchange_by_bucket<u32, OwnerFn>(d, idx, val, 0, val2, m, f, bufs[0], bufs[1], val2 ? bufs[6] : nullptr, bufs[2], bufs[3],
                                              val2 ? bufs[7] : nullptr, true, d.xchg, &R)));
    u32* si[2] = {bufs[2], bufs[4]};
    u32* sv[2] = {bufs[3], bufs[5]};
    TDC_TRY(partitioned_scatter(d.c.sortws, d.c.stream, si, sv, R, dst_shard, d.pos_cnt, R == d.pos_cnt));
    if (val2) {
        // bufs[0]/bufs[1] are free again: scratch pair of the second scatter (the received indices are intact)
        u32* si2[2] = {bufs[2], bufs[0]};
        u32* sv2[2] = {bufs[7], bufs[1]};
        TDC_TRY(partitioned_scatter(d.c.sortws, d.c.stream, si2, sv2, R, dst2, d.pos_cnt, R == d.pos_cnt));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// suffix array + ISA (+ key-seeded LCP)
// ---------------------------------------------------------------------------------------------------------------
template <bool FIRST>
static int dist_rerank(DistCtx& d, const u64* keys, const u32* vals, const u32* pos_in, u64 m, u32* agg_lasthead, ull* agg_cnt,
                       u32* rank_idx, u32* rank_val, u32* pos_out, u32* idx_out, u32* gid_out, u32* lcp_out, PackParams pp,
                       u64* m_out, u64* g_out) {
    Ctx& c = d.c;
    *m_out = 0;
    *g_out = 0;
    if (m == 0) return 0;
    const u32 ntiles = u32(div_up(m, RR_TILE));
    auto rerank_reduce = rerank_reduce_kernel<u64>;
    TDC_LAUNCH(rerank_reduce, ntiles, RR_THREADS, 0, c.stream, keys, m, agg_lasthead, agg_cnt, 0u);
    TDC_LAUNCH(rerank_scan_kernel, 1, 1024, 0, c.stream, agg_lasthead, agg_cnt, ntiles, c.d_scalars);
    auto rerank_apply = rerank_apply_kernel<u64, FIRST>;
    TDC_LAUNCH(rerank_apply, ntiles, RR_THREADS, 0, c.stream, keys, vals, pos_in, m, agg_lasthead, agg_cnt, d.d_sa, rank_idx,
               rank_val, pos_out, idx_out, gid_out, lcp_out, pp, u32(d.slot_lo), u32(d.n));
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 2 * sizeof(u32), cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    *m_out = c.h_scalars[0];
    *g_out = c.h_scalars[1];
    return 0;
}

static int dist_build_sa(DistCtx& d) {
    Ctx& c = d.c;
    const u64 n = d.n, cap = d.cap;
    const int P = d.P;
    cudaStream_t st = c.stream;
    c.sa_rounds = 0;
    c.sa_active_sum = 0;
    c.sortws.stat_passes = 0;
    c.sortws.stat_elems = 0;
    // ---- alphabet (every rank reads the whole replicated text: identical key layout everywhere) ----
    u32* d_hist = c.d_scalars + 16;
    TDC_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(u32), st));
    {
        const u32 grid = u32(std::min<u64>(u64(c.sm_count) * 8, div_up(div_up(n, 16), 256)));
        TDC_LAUNCH(byte_histogram_kernel, grid, 256, 0, st, c.d_text, n, d_hist);
        TDC_KCHECK();
    }
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 16, d_hist, 256 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    c.h_scalars[16 + 256] = 0xff;
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 16 + 256, c.d_text + (n - 1), 1, cudaMemcpyDeviceToHost, st));  // the last text byte
    TDC_CUDA(cudaStreamSynchronize(st));
    const u32* hist = c.h_scalars + 16;
    if (hist[0] != 1 || (c.h_scalars[16 + 256] & 0xffu) != 0) {  // same contract as the single-GPU builder (TextDS.hpp:132-138)
        set_error("text must contain exactly one 0 byte, at its end (found %u zero bytes, last byte 0x%02x)", hist[0],
                  c.h_scalars[16 + 256] & 0xffu);
        return TDCGPU_ERR_SENTINEL;
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
    u64* K[2] = {c.arena.take<u64>(cap), c.arena.take<u64>(cap)};
    u32* V[2] = {c.arena.take<u32>(cap), c.arena.take<u32>(cap)};
    u32* Q[2] = {c.arena.take<u32>(cap), c.arena.take<u32>(cap)};
    u32* G = c.arena.take<u32>(cap);
    u32* S[4] = {c.arena.take<u32>(cap), c.arena.take<u32>(cap), c.arena.take<u32>(cap), c.arena.take<u32>(cap)};
    const u64 rr_tiles = div_up(cap, RR_TILE);
    u32* agg_lasthead = c.arena.take<u32>(rr_tiles);
    ull* agg_cnt = c.arena.take<ull>(rr_tiles);
    uint8_t* d_code_map = c.arena.take<uint8_t>(256);
    const u32 NS = 1024;  // samples per rank
    u64* d_samples = c.arena.take<u64>(NS);
    if (!K[0] || !K[1] || !V[0] || !V[1] || !Q[0] || !Q[1] || !G || !S[0] || !S[1] || !S[2] || !S[3] || !agg_lasthead ||
        !agg_cnt || !d_code_map || !d_samples) {
        set_error("dist suffix array: scratch arena too small");
        return TDCGPU_ERR_NOMEM;
    }
    TDC_CUDA(cudaMemcpyAsync(d_code_map, code_map, 256, cudaMemcpyHostToDevice, st));

    // ---- round 0: keys of my positions, splitters, exchange, local sort ----
    if (d.pos_cnt) {
        auto pack_keys_kernel = tdc::pack_keys_kernel<false>;
        TDC_LAUNCH(pack_keys_kernel, u32(div_up(d.pos_cnt, PK_TILE)), PK_THREADS, 0, st, c.d_text, n, d_code_map, pp, K[0], d.pos_lo, d.pos_cnt, PassPlan(), (u32*)nullptr);
        TDC_KCHECK();
    }
    SplitterFn sf;
    sf.nspl = P - 1;
    for (int i = 0; i < DIST_MAX_RANKS - 1; i++) sf.spl[i] = ~u64(0);
    if (P > 1) {
        TDC_LAUNCH(sample_keys_kernel, u32(div_up(NS, 256)), 256, 0, st, K[0], d.pos_cnt, NS, d_samples);
        TDC_KCHECK();
        std::vector<u64> mine(NS), all(size_t(NS) * P);
        TDC_CUDA(cudaMemcpyAsync(mine.data(), d_samples, sizeof(u64) * NS, cudaMemcpyDeviceToHost, st));
        TDC_CUDA(cudaStreamSynchronize(st));
        TDC_TRY(d.comm->allgather_host(mine.data(), all.data(), sizeof(u64) * NS));
        std::sort(all.begin(), all.end());
        for (int i = 0; i + 1 < P; i++) sf.spl[i] = all[size_t(i + 1) * NS];
    }
    // ALL-TO-ALL #1: keys and suffix ids to the rank that owns their bucket.  Peer-memory transport: received straight
    // into K[1] / V[1] (K[0] is still being read by the partition); NCCL: staged in K[1] / V[1], received into K[0] / V[0].
    u64* krecv = d.p2p ? K[1] : K[0];
    u32* vrecv = d.p2p ? V[1] : V[0];
    u64 mr = 0;
    TDC_TRY((exchange_by_bucket<u64, SplitterFn>(d, K[0], nullptr, u32(d.pos_lo), nullptr, d.pos_cnt, sf, K[1], V[1], nullptr, krecv,
                                                 vrecv, nullptr, true, d.xchg, &mr)));
    d.slot_cnts.assign(P, 0);
    for (int r = 0; r < P; r++)
        for (int p = 0; p < P; p++) d.slot_cnts[r] += d.xchg[size_t(p) * P + r];
    d.slot_lo = 0;
    for (int r = 0; r < d.rank; r++) d.slot_lo += d.slot_cnts[r];
    d.slot_cnt = d.slot_cnts[d.rank];
    if (d.p2p) { std::swap(K[0], K[1]); std::swap(V[0], V[1]); }  // the received pairs are the sort's input, slot 0
    int res = 0;
    TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, K, V, mr, 0, int(sigbits), false, &res));
    u64 m = 0, g = 0;
    int pcur = 0;
    TDC_TRY((dist_rerank<true>(d, K[res], V[res], nullptr, mr, agg_lasthead, agg_cnt, nullptr, reinterpret_cast<u32*>(K[res ^ 1]),
                               Q[pcur], V[res ^ 1], G, d.d_lcp, pp, &m, &g)));
    {
        // ALL-TO-ALL #2: rank[suffix] = head slot, to the owner of the position
        u32* bufs[6] = {S[0], S[1], S[2], S[3], reinterpret_cast<u32*>(K[res]), reinterpret_cast<u32*>(K[res]) + cap};
        TDC_TRY(dist_scatter(d, V[res], reinterpret_cast<u32*>(K[res ^ 1]), mr, d.d_rank, bufs));
    }
    c.sa_rounds = 1;
    c.sa_active_sum = mr;
    c.sa_first_residue = m;

    // ---- doubling rounds ----
    const u32 rbits = bits_for_host(n - 1);
    OwnerFn of;
    of.block = d.block;
    u64 h = pp.k;
    while (true) {
        u64 mine[1] = {m};
        std::vector<u64> all;
        TDC_TRY(allgather_u64(d, mine, 1, all));
        u64 m_tot = 0;
        for (u64 x : all) m_tot += x;
        if (m_tot == 0) break;
        if (h > 2 * n) { set_error("dist suffix array: doubling did not converge"); return TDCGPU_ERR_INTERNAL; }
        u32* Vact = V[res ^ 1];
        // rank[suffix + h] from its owner: request / reply
        if (m) TDC_LAUNCH(add_offset_kernel, u32(div_up(m, 256)), 256, 0, st, Vact, m, u32(h), S[0]);
        u64 R = 0;  // requests (owner-local positions) travel to S[3] of the owner; the origin slots stay here, in S[2]
        TDC_TRY((exchange_by_bucket<u32, OwnerFn>(d, S[0], nullptr, 0u, nullptr, m, of, S[1], S[2], nullptr, S[3], nullptr, nullptr, false,
                                                  d.xchg, &R)));
        if (R) TDC_LAUNCH(gather_u32_kernel, u32(div_up(R, 256)), 256, 0, st, d.d_rank, S[3], R, S[0]);
        {
            std::vector<u64> back(size_t(P) * P);  // the replies travel the transposed way
            for (int r = 0; r < P; r++)
                for (int p = 0; p < P; p++) back[size_t(r) * P + p] = d.xchg[size_t(p) * P + r];
            TDC_TRY(a2a_elems(d, S[0], S[1], 4, back.data()));
        }
        if (m) {
            TDC_LAUNCH(scatter_pairs_kernel, u32(div_up(m, SP_THREADS * SP_EPT)), SP_THREADS, 0, st, S[2], S[1], m, S[3]);
            TDC_LAUNCH(build_keys_from_kernel, u32(div_up(m, 256)), 256, 0, st, G, S[3], m, rbits, K[0]);
            TDC_KCHECK();
        }
        const int gbits = g > 1 ? int(bits_for_host(g - 1)) : 0;
        u64* k2[2] = {K[0], K[1]};
        u32* v2[2] = {V[res ^ 1], V[res]};
        int r2 = 0;
        TDC_TRY(radix_sort_pairs<u64>(c.sortws, st, k2, v2, m, 0, int(rbits) + gbits, false, &r2));
        c.sa_rounds++;
        c.sa_active_sum += m;
        u64 m_new = 0, g_new = 0;
        u32* rank_idx = reinterpret_cast<u32*>(k2[r2 ^ 1]);
        u32* rank_val = reinterpret_cast<u32*>(k2[r2 ^ 1]) + cap;
        TDC_TRY((dist_rerank<false>(d, k2[r2], v2[r2], Q[pcur], m, agg_lasthead, agg_cnt, rank_idx, rank_val, Q[pcur ^ 1],
                                    v2[r2 ^ 1], G, nullptr, pp, &m_new, &g_new)));
        {
            u32* bufs[6] = {S[0], S[1], S[2], S[3], reinterpret_cast<u32*>(k2[r2]), reinterpret_cast<u32*>(k2[r2]) + cap};
            TDC_TRY(dist_scatter(d, rank_idx, rank_val, m, d.d_rank, bufs));
        }
        if (v2[r2 ^ 1] != V[res ^ 1]) res ^= 1;
        pcur ^= 1;
        m = m_new;
        g = g_new;
        h *= 2;
    }
    return 0;
}

static int dist_build_lcp(DistCtx& d) {
    Ctx& c = d.c;
    cudaStream_t st = c.stream;
    const u64 mr = d.slot_cnt;
    c.arena.reset();
    u32* queue = c.arena.take<u32>(d.cap);
    if (!queue) { set_error("dist lcp: scratch arena too small"); return TDCGPU_ERR_NOMEM; }
    u32* d_qlen = c.d_scalars + 0;
    u32* d_max = c.d_scalars + 1;
    TDC_CUDA(cudaMemsetAsync(c.d_scalars, 0, 2 * sizeof(u32), st));
    u64 mine[2] = {0, mr};
    if (mr) {
        TDC_LAUNCH(lcp_fix_kernel, u32(div_up(mr, 256 * LCPFIX_EPT)), 256, 0, st, c.d_text, d.d_sa, mr, d.d_lcp, c.symbols_per_key, queue, d_qlen, d_max);
        TDC_KCHECK();
        TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 8, d.d_sa + (mr - 1), sizeof(u32), cudaMemcpyDeviceToHost, st));
        TDC_CUDA(cudaStreamSynchronize(st));
        mine[0] = c.h_scalars[8];
    }
    std::vector<u64> all;
    TDC_TRY(allgather_u64(d, mine, 2, all));
    if (mr && d.slot_lo > 0) {
        int pr = d.rank - 1;
        while (pr >= 0 && all[size_t(pr) * 2 + 1] == 0) pr--;
        if (pr < 0) { set_error("dist lcp: inconsistent slot ranges"); return TDCGPU_ERR_INTERNAL; }
        TDC_LAUNCH(lcp_boundary_kernel, 1, 32, 0, st, c.d_text, d.d_sa, u32(all[size_t(pr) * 2]), d.d_lcp, d_max);
    }
    if (mr) TDC_LAUNCH(lcp_direct_long_kernel, u32(c.sm_count * 4), 256, 0, st, c.d_text, d.d_sa, d.d_lcp, queue, d_qlen, d_max);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    u64 mx[1] = {c.h_scalars[1]};
    TDC_TRY(allgather_u64(d, mx, 1, all));
    c.max_lcp = 0;
    for (u64 x : all) c.max_lcp = std::max<u32>(c.max_lcp, u32(x));
    c.lcp_route = 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// factorisation
// ---------------------------------------------------------------------------------------------------------------
static int dist_factorize(DistCtx& d, u32 threshold) {
    Ctx& c = d.c;
    cudaStream_t st = c.stream;
    const int P = d.P, rank = d.rank;
    const u64 cap = d.cap, qcap = d.qcap;
    const u32 mr = u32(d.slot_cnt);
    c.num_factors = 0;
    d.total_factors = 0;
    c.flen_min = 0xffffffffu;
    c.flen_max = 0;
    c.arena.reset();
    // ---- local min-tree ----
    MinTree T;
    T.a[0] = d.d_sa;
    T.l[0] = d.d_lcp;
    T.sz[0] = mr;
    T.nlev = 1;
    while (T.sz[T.nlev - 1] > 32) {
        if (T.nlev >= MT_MAX_LEVELS) { set_error("min-tree too deep"); return TDCGPU_ERR_INTERNAL; }
        const u32 szi = T.sz[T.nlev - 1], szo = u32(div_up(szi, 32));
        u32* a = c.arena.take<u32>(szo);
        u32* l = c.arena.take<u32>(szo);
        if (!a || !l) { set_error("dist lzss_lcp: scratch arena too small"); return TDCGPU_ERR_NOMEM; }
        TDC_LAUNCH(mintree_level_kernel, u32(div_up(div_up(u64(szo), MT_OUT_PER_WARP) * 32, 256)), 256, 0, st, T.a[T.nlev - 1], T.l[T.nlev - 1], szi, a, l, szo);
        T.a[T.nlev] = a;
        T.l[T.nlev] = l;
        T.sz[T.nlev] = szo;
        T.nlev++;
    }
    for (int i = T.nlev; i < MT_MAX_LEVELS; i++) { T.a[i] = nullptr; T.l[i] = nullptr; T.sz[i] = 0; }
    TDC_KCHECK();

    u32* W[10];  // lu, su, ld, sd, then six more work buffers
    for (int i = 0; i < 10; i++) W[i] = c.arena.take<u32>(cap);
    u32* lenside_t = c.arena.take<u32>(cap + 64);  // text order
    u32* src_t = c.arena.take<u32>(cap + 64);
    WalkQuery* qb[6];  // up A/B, down A/B, received up, received down
    for (int i = 0; i < 6; i++) qb[i] = c.arena.take<WalkQuery>(qcap);
    WalkAnswer* ab[4];  // answers up, answers down, received answers up / down
    for (int i = 0; i < 4; i++) ab[i] = c.arena.take<WalkAnswer>(qcap);
    u32* d_cnt = c.d_scalars + 8;  // [4]
    if (!W[9] || !lenside_t || !src_t || !qb[5] || !ab[3]) { set_error("dist lzss_lcp: scratch arena too small"); return TDCGPU_ERR_NOMEM; }
    u32 *lu = W[0], *su = W[1], *ld = W[2], *sd = W[3];

    // ---- LPF per slot; walks leaving the shard are queued ----
    TDC_CUDA(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(u32), st));
    if (mr) {
        LpfDistOut lo;
        lo.lu = lu; lo.su = su; lo.ld = ld; lo.sd = sd;
        lo.q_up = qb[0]; lo.q_dn = qb[2]; lo.q_cnt = d_cnt; lo.qcap = u32(qcap);
        auto lpf_tile_dist = lpf_tile_kernel<true>;
        TDC_CUDA(cudaFuncSetAttribute(lpf_tile_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lpf_smem_bytes())));
        TDC_LAUNCH(lpf_tile_dist, u32(div_up(u64(mr), LPF_TILE)), LPF_THREADS, lpf_smem_bytes(), st, T, mr, threshold, (u32*)nullptr, lo);
        TDC_KCHECK();
    }
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 8, d_cnt, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    u64 cu = c.h_scalars[8], cd = c.h_scalars[9];
    int cur = 0;  // qb[cur] / qb[2 + cur] hold the queries to send
    std::vector<u64> all;
    for (int hop = 1; hop < P; hop++) {
        u64 mine[2] = {cu, cd};
        TDC_TRY(allgather_u64(d, mine, 2, all));
        u64 outstanding = 0, worst = 0;
        for (u64 x : all) { outstanding += x; worst = std::max(worst, x); }
        if (worst > qcap) { set_error("dist lzss_lcp: %llu PSV/NSV walks leave a shard (capacity %llu)", (unsigned long long)worst, (unsigned long long)qcap); return TDCGPU_ERR_NOMEM; }
        if (outstanding == 0) break;
        // queries: up-queue to rank-1, down-queue to rank+1 (every rank knows every count: full matrices)
        std::vector<u64> mx(size_t(P) * P);
        std::fill(mx.begin(), mx.end(), 0);
        for (int r = 1; r < P; r++) mx[size_t(r) * P + (r - 1)] = all[size_t(r) * 2];
        const u64 nru = rank + 1 < P ? all[size_t(rank + 1) * 2] : 0;
        TDC_TRY(a2a_elems(d, qb[cur], qb[4], sizeof(WalkQuery), mx.data()));
        std::fill(mx.begin(), mx.end(), 0);
        for (int r = 0; r + 1 < P; r++) mx[size_t(r) * P + (r + 1)] = all[size_t(r) * 2 + 1];
        const u64 nrd = rank > 0 ? all[size_t(rank - 1) * 2 + 1] : 0;
        TDC_TRY(a2a_elems(d, qb[2 + cur], qb[5], sizeof(WalkQuery), mx.data()));
        // resolve against my tree
        TDC_CUDA(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(u32), st));
        if (nru) {
            auto resolve_up = resolve_queries_kernel<true>;
            TDC_LAUNCH(resolve_up, u32(div_up(nru, 128)), 128, 0, st, T, mr, threshold, qb[4], u32(nru), ab[0], qb[cur ^ 1], d_cnt);
        }
        if (nrd) {
            auto resolve_dn = resolve_queries_kernel<false>;
            TDC_LAUNCH(resolve_dn, u32(div_up(nrd, 128)), 128, 0, st, T, mr, threshold, qb[5], u32(nrd), ab[1], qb[2 + (cur ^ 1)], d_cnt + 2);
        }
        TDC_KCHECK();
        TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 8, d_cnt, 4 * sizeof(u32), cudaMemcpyDeviceToHost, st));
        TDC_CUDA(cudaStreamSynchronize(st));
        const u64 au = c.h_scalars[8], fu = c.h_scalars[9], ad = c.h_scalars[10], fd = c.h_scalars[11];
        // answers straight back to the origin: up-queries seen at this hop came from rank + hop, down from rank - hop
        u64 am[2] = {au, ad};
        TDC_TRY(allgather_u64(d, am, 2, all));
        std::fill(mx.begin(), mx.end(), 0);
        for (int r = 0; r + hop < P; r++) mx[size_t(r) * P + (r + hop)] = all[size_t(r) * 2];
        const u64 rau = rank - hop >= 0 ? all[size_t(rank - hop) * 2] : 0;
        TDC_TRY(a2a_elems(d, ab[0], ab[2], sizeof(WalkAnswer), mx.data()));
        if (rau) TDC_LAUNCH(apply_answers_kernel, u32(div_up(rau, 256)), 256, 0, st, ab[2], u32(rau), lu, su);
        std::fill(mx.begin(), mx.end(), 0);
        for (int r = hop; r < P; r++) mx[size_t(r) * P + (r - hop)] = all[size_t(r) * 2 + 1];
        const u64 rad = rank + hop < P ? all[size_t(rank + hop) * 2 + 1] : 0;
        TDC_TRY(a2a_elems(d, ab[1], ab[3], sizeof(WalkAnswer), mx.data()));
        if (rad) TDC_LAUNCH(apply_answers_kernel, u32(div_up(rad, 256)), 256, 0, st, ab[3], u32(rad), ld, sd);
        TDC_KCHECK();
        cu = rank > 0 ? fu : 0;        // rank 0 has nobody above: what is still open has no PSV
        cd = rank + 1 < P ? fd : 0;
        cur ^= 1;
    }
    // ---- combine, then to text order through the position owners ----
    u32 *lenside_r = W[4], *src_r = W[5];
    if (mr) {
        TDC_LAUNCH(lpf_combine_kernel, u32(div_up(u64(mr), 256)), 256, 0, st, lu, su, ld, sd, u64(mr), threshold, lenside_r, src_r);
        TDC_KCHECK();
    }
    {
        u32* bufs[8] = {W[0], W[1], W[2], W[3], W[6], W[7], W[8], W[9]};
        TDC_TRY(dist_scatter(d, d.d_sa, lenside_r, mr, lenside_t, bufs, src_r, src_t));
    }
    // ---- greedy chain over my positions; the entry point comes from the previous rank ----
    const u32 n_eff = u32(rank == P - 1 ? d.pos_cnt : d.pos_cnt + 1);  // positions < n_eff - 1 are chain nodes here
    const u32 ntiles = u32(div_up(u64(n_eff), CH_TILE));
    u32* exitp = W[0];
    u32* entry = W[1];
    u32* tile_cnt = W[2];
    u32* region_exit = W[3];
    u32* fmask = W[6];
    u32 regions = 1, tiles_per_region = 1;
    if (n_eff > 1) {
        TDC_LAUNCH(chain_exit_kernel, ntiles, CH_THREADS, 0, st, lenside_t, n_eff, exitp);
        TDC_LAUNCH(fill_u32_kernel, u32(div_up(u64(ntiles), 256)), 256, 0, st, entry, u64(ntiles), CH_NONE);
        while (u64(regions) * regions * 2 < ntiles) regions *= 2;
        tiles_per_region = u32(div_up(u64(ntiles), regions));
        regions = u32(div_up(u64(ntiles), tiles_per_region));
        TDC_LAUNCH(chain_entries_spec_kernel, u32(div_up(u64(regions), 128)), 128, 0, st, exitp, n_eff, tiles_per_region, regions, entry, region_exit);
        TDC_CUDA(cudaMemsetAsync(fmask, 0, sizeof(u32) * u64(ntiles) * (CH_TILE / 32), st));
        TDC_KCHECK();
    }
    u64 chain_pos = 0;  // global position where the chain enters the next rank's range
    for (int r = 0; r < P; r++) {
        u64 mine[1] = {chain_pos};
        if (r == rank && n_eff > 1) {
            // always run the stitcher: it also erases the speculative entries when the chain jumps over this range
            const u64 first = chain_pos > d.pos_lo ? chain_pos - d.pos_lo : 0;
            TDC_LAUNCH(chain_entries_stitch_kernel, 1, 32, 0, st, exitp, n_eff, tiles_per_region, entry, region_exit, u32(first), c.d_scalars + 12);
            TDC_KCHECK();
            TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 12, c.d_scalars + 12, sizeof(u32), cudaMemcpyDeviceToHost, st));
            TDC_CUDA(cudaStreamSynchronize(st));
            mine[0] = d.pos_lo + c.h_scalars[12];
        }
        TDC_TRY(allgather_u64(d, mine, 1, all));
        chain_pos = all[r];
    }
    u32* d_total = c.d_scalars + 0;
    u32* d_minmax = c.d_scalars + 2;
    u64 z = 0;
    c.h_scalars[2] = 0xffffffffu;
    c.h_scalars[3] = 0;
    TDC_CUDA(cudaMemcpyAsync(d_minmax, c.h_scalars + 2, 2 * sizeof(u32), cudaMemcpyHostToDevice, st));
    if (n_eff > 1) {
        TDC_LAUNCH(chain_mark_kernel, u32(div_up(u64(ntiles), 128)), 128, 0, st, lenside_t, n_eff, ntiles, entry, fmask, tile_cnt);
        TDC_LAUNCH(scan_counts_kernel, 1, 1024, 0, st, tile_cnt, ntiles, d_total);
        TDC_KCHECK();
        TDC_CUDA(cudaMemcpyAsync(c.h_scalars, d_total, sizeof(u32), cudaMemcpyDeviceToHost, st));
        TDC_CUDA(cudaStreamSynchronize(st));
        z = c.h_scalars[0];
        if (z > c.factors_cap) {
            if (c.d_factors) TDC_CUDA(cudaFree(c.d_factors));
            c.d_factors = nullptr;
            c.factors_cap = 0;
            const u64 fcap = z + z / 8 + 1024;
            TDC_CUDA(cudaMalloc(&c.d_factors, fcap * sizeof(Factor)));
            c.factors_cap = fcap;
        }
        if (z > 0) {
            auto emit_factors = emit_factors_kernel<true>;
            TDC_LAUNCH(emit_factors, ntiles, CH_TILE / 32, 0, st, T, (const u32*)nullptr, lenside_t, fmask, tile_cnt, threshold, c.d_factors, d_minmax, src_t, u32(d.pos_lo));
            TDC_KCHECK();
        }
    }
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 2, d_minmax, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    u64 mine[3] = {z, c.h_scalars[2], c.h_scalars[3]};
    TDC_TRY(allgather_u64(d, mine, 3, all));
    c.num_factors = z;
    for (int r = 0; r < P; r++) {
        d.total_factors += all[size_t(r) * 3];
        c.flen_min = std::min<u32>(c.flen_min, u32(all[size_t(r) * 3 + 1]));
        c.flen_max = std::max<u32>(c.flen_max, u32(all[size_t(r) * 3 + 2]));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------------------------
static void dist_free_arrays(DistCtx& d) {
    Ctx& c = d.c;
    if (d.p2p && d.comm) {
        // importers unmap first, then (after a barrier) the owners free
        cudaStreamSynchronize(c.stream);
        d.comm->close_window();
        u64 token = 0;
        std::vector<u64> all(d.P);
        d.comm->allgather_host(&token, all.data(), sizeof(u64));
        d.p2p = false;
    }
    void* ps[] = {c.d_text, d.d_sa, d.d_rank, d.d_lcp, c.arena.base};
    for (void* p : ps)
        if (p) cudaFree(p);
    c.d_text = nullptr;
    d.d_sa = d.d_rank = d.d_lcp = nullptr;
    c.arena = Arena();
    c.cap_n = 0;
    d.cap = 0;
    d.have = 0;
}

static int dist_ensure_capacity(DistCtx& d, u64 n) {
    Ctx& c = d.c;
    const u64 per = div_up(n, u64(d.P));
    const u64 cap = per + per / 4 + (u64(1) << 20);  // 25 % head room for uneven buckets
    if (n <= c.cap_n && cap <= d.cap) return 0;
    dist_free_arrays(d);
    d.qcap = std::min<u64>(cap, u64(1) << 24);
    TDC_CUDA(cudaMalloc(&c.d_text, n + 1024 + 16));
    TDC_CUDA(cudaMalloc(&d.d_sa, sizeof(u32) * cap));
    TDC_CUDA(cudaMalloc(&d.d_rank, sizeof(u32) * cap));
    TDC_CUDA(cudaMalloc(&d.d_lcp, sizeof(u32) * cap));
    const size_t arena_bytes = size_t(72) * cap + cap / 4 + size_t(10 * 12) * d.qcap + (size_t(16) << 20);
    TDC_CUDA(cudaMalloc(&c.arena.base, arena_bytes));
    c.arena.cap = arena_bytes;
    c.arena.off = 0;
    TDC_TRY(sort_workspace_init(c.sortws, cap, c.sm_count));
    c.cap_n = n;
    d.cap = cap;
    // peer-memory transport for the exchanges: every rank maps every other rank's arena (same size, carved identically)
    d.p2p = d.comm->open_window(c.arena.base, arena_bytes, d.peer_arena) == 0 && d.P > 1;
    if (d.p2p) {
        for (int p = 0; p < d.P; p++) {
            if (p != d.rank && !d.copy_streams[p]) TDC_CUDA(cudaStreamCreateWithFlags(&d.copy_streams[p], cudaStreamNonBlocking));
            if (!d.copy_events[p]) TDC_CUDA(cudaEventCreateWithFlags(&d.copy_events[p], cudaEventDisableTiming));
        }
        if (!d.copy_events[DIST_MAX_RANKS]) TDC_CUDA(cudaEventCreateWithFlags(&d.copy_events[DIST_MAX_RANKS], cudaEventDisableTiming));
    }
    return 0;
}

}  // namespace tdc

using namespace tdc;

struct tdcgpu_dist {
    DistCtx d;
};

static int dist_create_common(int device, Comm* comm, tdcgpu_dist* h) {
    DistCtx& d = h->d;
    Ctx& c = d.c;
    d.comm = comm;
    d.P = comm->nranks;
    d.rank = comm->rank;
    c.device = device;
    TDC_CUDA(cudaMalloc(&d.d_counts, sizeof(ull) * 2 * DIST_MAX_RANKS));
    TDC_CUDA(cudaMallocHost(&d.h_counts, sizeof(ull) * 64));
    return 0;
}

#define DIST_GUARD(h)                                                          \
    if (!(h)) { set_error("null context"); return TDCGPU_ERR_ARG; }            \
    DistCtx& d = (h)->d;                                                       \
    Ctx& c = d.c;                                                              \
    (void)c;                                                                   \
    if (cudaSetDevice(c.device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", c.device); return TDCGPU_ERR_CUDA; }

static int dist_ctx_init(Ctx& c, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device available (tdcgpu has no CPU fallback)");
        return TDCGPU_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return TDCGPU_ERR_ARG; }
    TDC_CUDA(cudaSetDevice(device));
    c.device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c.sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c.d_scalars, 512 * sizeof(u32)) != cudaSuccess || cudaMallocHost(&c.h_scalars, 512 * sizeof(u32)) != cudaSuccess) {
        set_error("context allocation failed");
        return TDCGPU_ERR_CUDA;
    }
    return 0;
}

extern "C" {

#ifdef TDC_CUSIM
int tdcsim_dist_create(int rank, int nranks, tdcsim_allgather_fn ag, tdcsim_alltoallv_fn a2a, void* user, tdcgpu_dist** out) {
    if (!out || nranks < 1 || nranks > DIST_MAX_RANKS || rank < 0 || rank >= nranks) { set_error("bad arguments"); return TDCGPU_ERR_ARG; }
    tdcgpu_dist* h = new (std::nothrow) tdcgpu_dist();
    if (!h) return TDCGPU_ERR_NOMEM;
    int rc = dist_ctx_init(h->d.c, 0);
    if (rc == 0) rc = dist_create_common(0, make_callback_comm(rank, nranks, ag, a2a, user), h);
    if (rc < 0) { delete h; return rc; }
    *out = h;
    return 0;
}
#else
int tdcgpu_dist_unique_id(uint8_t id[128]) {
    if (!id) { set_error("null id"); return TDCGPU_ERR_ARG; }
    return nccl_unique_id(id) < 0 ? TDCGPU_ERR_CUDA : 0;
}

int tdcgpu_dist_create(int device, int rank, int nranks, const uint8_t id[128], tdcgpu_dist** out) {
    if (!out || nranks < 1 || nranks > DIST_MAX_RANKS || rank < 0 || rank >= nranks || (nranks > 1 && !id)) {
        set_error("tdcgpu_dist_create: bad arguments (1 <= nranks <= %d)", DIST_MAX_RANKS);
        return TDCGPU_ERR_ARG;
    }
    *out = nullptr;
    tdcgpu_dist* h = new (std::nothrow) tdcgpu_dist();
    if (!h) { set_error("out of host memory"); return TDCGPU_ERR_NOMEM; }
    int rc = dist_ctx_init(h->d.c, device);
    if (rc < 0) { delete h; return rc; }
    Comm* comm = make_nccl_comm(rank, nranks, id, h->d.c.stream);
    if (!comm) { delete h; return TDCGPU_ERR_CUDA; }
    rc = dist_create_common(device, comm, h);
    if (rc < 0) { delete comm; delete h; return rc; }
    *out = h;
    return 0;
}
#endif

void tdcgpu_dist_destroy(tdcgpu_dist* h) {
    if (!h) return;
    DistCtx& d = h->d;
    Ctx& c = d.c;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    dist_free_arrays(d);
    for (int p = 0; p < DIST_MAX_RANKS; p++)
        if (d.copy_streams[p]) cudaStreamDestroy(d.copy_streams[p]);
    for (int p = 0; p <= DIST_MAX_RANKS; p++)
        if (d.copy_events[p]) cudaEventDestroy(d.copy_events[p]);
    delete d.comm;
    sort_workspace_free(c.sortws);
    if (c.d_factors) cudaFree(c.d_factors);
    if (c.d_scalars) cudaFree(c.d_scalars);
    if (c.h_scalars) cudaFreeHost(c.h_scalars);
    if (d.d_counts) cudaFree(d.d_counts);
    if (d.h_counts) cudaFreeHost(d.h_counts);
    host_copier_free(c.copier);
    for (auto& e : c.user_events)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(c.stream);
    delete h;
}

int tdcgpu_dist_set_text(tdcgpu_dist* h, const uint8_t* text, uint64_t n, int on_device) {
    DIST_GUARD(h);
    if (!text || n == 0) { set_error("empty text (the path always sees at least the sentinel)"); return TDCGPU_ERR_ARG; }
    if (n >= (uint64_t(1) << 32) - 1) { set_error("n = %llu: indices are 32-bit", (unsigned long long)n); return TDCGPU_ERR_ARG; }
    if (!on_device && text[n - 1] != 0) {
        set_error("Input has no sentinel! (the last text byte must be 0, ds/TextDS.hpp:132-138)");
        return TDCGPU_ERR_SENTINEL;
    }
    TDC_TRY(dist_ensure_capacity(d, n));
    d.n = n;
    c.n = n;
    d.block = u32(div_up(n, u64(d.P)));
    d.pos_lo = std::min<u64>(n, u64(d.rank) * d.block);
    d.pos_cnt = std::min<u64>(n - d.pos_lo, d.block);
    d.have = 0;
    c.max_lcp = 0;
    c.num_factors = 0;
    c.phases.clear();
    // Slice upload: every rank uploads only its own n/P slice over PCIe and receives the other slices from its peers over
    // NVLink, instead of n bytes over PCIe per rank (which bounds the end-to-end figure of the sharded path: 8 ranks x 4 GB
    // through one host, profiles/r2_summary.md §4).  Peer-memory transport: the slice is pushed into every peer's
    // SCRATCH ARENA (the window all ranks have mapped; its content is dead between two texts) at the slice's own text
    // offset, and every rank then copies the received slices from its arena into its text buffer.  Two host barriers:
    // "your arena is free" before the pushes, "all pushes have landed" after them.  Default when the peer window exists;
    // TDCGPU_DIST_SLICE_UPLOAD=0 switches it off, =nccl takes grouped ncclSend/ncclRecv into the text buffer instead
    // (measured 8x slower than the redundant upload at 8 GPUs, r2o_distab.txt).
    const char* slice_env = std::getenv("TDCGPU_DIST_SLICE_UPLOAD");
    const bool slice_off = slice_env && slice_env[0] == '0';
    const bool slice_nccl = slice_env && slice_env[0] == 'n';
    if (!on_device && d.P > 1 && !slice_off && d.p2p && !slice_nccl && c.arena.cap >= n) {
        cudaStream_t st = c.stream;
        TDC_CUDA(cudaMemsetAsync(c.d_text + n, 0, 1024 + 16, st));
        if (d.pos_cnt) TDC_TRY(host_copy(c, c.d_text + d.pos_lo, text + d.pos_lo, d.pos_cnt, true));
        TDC_CUDA(cudaStreamSynchronize(st));
        u64 token = 1;
        std::vector<u64> all(d.P);
        TDC_TRY(d.comm->allgather_host(&token, all.data(), sizeof(u64)));  // every rank is done with its previous text: arenas are free
        u64 bytes = 0;
        if (d.pos_cnt) {
            for (int p = 0; p < d.P; p++) {
                if (p == d.rank) continue;
                TDC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(d.peer_arena[p]) + d.pos_lo, c.d_text + d.pos_lo, d.pos_cnt, cudaMemcpyDeviceToDevice, d.copy_streams[p]));
                bytes += d.pos_cnt;
            }
            for (int p = 0; p < d.P; p++)
                if (p != d.rank) TDC_CUDA(cudaStreamSynchronize(d.copy_streams[p]));
        }
        d.p2p_bytes += bytes;
        TDC_TRY(d.comm->allgather_host(&token, all.data(), sizeof(u64)));  // all pushes have landed
        // the peers' slices: arena -> text (everything but the own slice)
        if (d.pos_lo) TDC_CUDA(cudaMemcpyAsync(c.d_text, c.arena.base, d.pos_lo, cudaMemcpyDeviceToDevice, st));
        const u64 hi = d.pos_lo + d.pos_cnt;
        if (hi < n) TDC_CUDA(cudaMemcpyAsync(c.d_text + hi, c.arena.base + hi, n - hi, cudaMemcpyDeviceToDevice, st));
        TDC_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    if (!on_device && d.P > 1 && slice_nccl) {
        TDC_CUDA(cudaMemsetAsync(c.d_text + n, 0, 1024 + 16, c.stream));
        if (d.pos_cnt) TDC_TRY(host_copy(c, c.d_text + d.pos_lo, text + d.pos_lo, d.pos_cnt, true));
        TDC_CUDA(cudaStreamSynchronize(c.stream));
        u64 soff[DIST_MAX_RANKS], scnt[DIST_MAX_RANKS], roff[DIST_MAX_RANKS], rcnt[DIST_MAX_RANKS];
        for (int p = 0; p < d.P; p++) {
            const u64 lo = std::min<u64>(n, u64(p) * d.block), cnt = std::min<u64>(n - lo, d.block);
            soff[p] = d.pos_lo;
            scnt[p] = p == d.rank ? 0 : d.pos_cnt;  // my slice to every peer
            roff[p] = lo;
            rcnt[p] = p == d.rank ? 0 : cnt;        // every peer's slice to its place in my copy
        }
        TDC_TRY(d.comm->alltoallv(c.d_text, soff, scnt, c.d_text, roff, rcnt, c.stream));
        TDC_CUDA(cudaStreamSynchronize(c.stream));
        return 0;
    }
    TDC_CUDA(cudaMemsetAsync(c.d_text + n, 0, 1024 + 16, c.stream));
    if (on_device) {
        TDC_CUDA(cudaMemcpyAsync(c.d_text, text, n, cudaMemcpyDeviceToDevice, c.stream));
        TDC_CUDA(cudaStreamSynchronize(c.stream));
    } else {
        TDC_TRY(host_copy(c, c.d_text, text, n, true));
    }
    return 0;
}

int tdcgpu_dist_build(tdcgpu_dist* h, uint32_t flags) {
    DIST_GUARD(h);
    if (d.n == 0) { set_error("no text loaded"); return TDCGPU_ERR_STATE; }
    if (flags & ~(DS_SA | DS_ISA | DS_LCP)) { set_error("multi-GPU mode builds SA, ISA and LCP (flags 0x%x)", flags); return TDCGPU_ERR_ARG; }
    c.phases.clear();
    if (!(d.have & DS_SA)) {
        PhaseTimer t(c, "Construct SA");
        TDC_TRY(dist_build_sa(d));
        d.have |= DS_SA | DS_ISA;
    }
    if ((flags & DS_LCP) && !(d.have & DS_LCP)) {
        PhaseTimer t(c, "Construct LCP Array");
        TDC_TRY(dist_build_lcp(d));
        d.have |= DS_LCP;
    }
    return 0;
}

int tdcgpu_dist_shard_info(tdcgpu_dist* h, uint64_t out[4]) {
    DIST_GUARD(h);
    out[0] = d.slot_lo;
    out[1] = d.slot_cnt;
    out[2] = d.pos_lo;
    out[3] = d.pos_cnt;
    return 0;
}

int tdcgpu_dist_get(tdcgpu_dist* h, uint32_t which, void* dst, int to_device) {
    DIST_GUARD(h);
    if (!(d.have & which)) { set_error("structure 0x%x has not been built", which); return TDCGPU_ERR_STATE; }
    const u32* src = which == DS_SA ? d.d_sa : which == DS_ISA ? d.d_rank : which == DS_LCP ? d.d_lcp : nullptr;
    if (!src || !dst) { set_error("bad arguments"); return TDCGPU_ERR_ARG; }
    const u64 cnt = which == DS_ISA ? d.pos_cnt : d.slot_cnt;
    if (cnt) TDC_CUDA(cudaMemcpyAsync(dst, src, sizeof(u32) * cnt, to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_dist_max_lcp(tdcgpu_dist* h, uint32_t* max_lcp) {
    DIST_GUARD(h);
    if (!(d.have & DS_LCP)) { set_error("LCP has not been built"); return TDCGPU_ERR_STATE; }
    if (max_lcp) *max_lcp = c.max_lcp;
    return 0;
}

int tdcgpu_dist_lzss_lcp_factorize(tdcgpu_dist* h, uint32_t threshold, uint64_t* local_count, uint64_t* total_count,
                                   uint32_t* min_len, uint32_t* max_len) {
    DIST_GUARD(h);
    if (threshold < 1) { set_error("lzss_lcp: threshold must be >= 1"); return TDCGPU_ERR_ARG; }
    TDC_TRY(tdcgpu_dist_build(h, DS_SA | DS_ISA | DS_LCP));
    if (c.max_lcp >= (1u << 31)) { set_error("multi-GPU lzss_lcp: common prefixes of 2^31 bytes or more are not supported"); return TDCGPU_ERR_ARG; }
    {
        PhaseTimer t(c, "Factorize");
        TDC_TRY(dist_factorize(d, threshold));
    }
    if (local_count) *local_count = c.num_factors;
    if (total_count) *total_count = d.total_factors;
    if (min_len) *min_len = c.flen_min;
    if (max_len) *max_len = c.flen_max;
    return 0;
}

int tdcgpu_dist_get_factors(tdcgpu_dist* h, tdcgpu_factor* dst, uint64_t cap, int to_device) {
    DIST_GUARD(h);
    if (c.num_factors > cap) { set_error("factor buffer too small"); return TDCGPU_ERR_ARG; }
    if (c.num_factors == 0) return 0;
    if (!dst) { set_error("null destination"); return TDCGPU_ERR_ARG; }
    TDC_CUDA(cudaMemcpyAsync(dst, c.d_factors, sizeof(Factor) * c.num_factors, to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

const uint8_t* tdcgpu_dist_text_device_ptr(tdcgpu_dist* h) { return h ? h->d.c.d_text : nullptr; }
const tdcgpu_factor* tdcgpu_dist_factors_device_ptr(tdcgpu_dist* h) {
    return h ? reinterpret_cast<const tdcgpu_factor*>(h->d.c.d_factors) : nullptr;
}

int tdcgpu_dist_sync(tdcgpu_dist* h) {
    DIST_GUARD(h);
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_dist_event_record(tdcgpu_dist* h, int slot) {
    DIST_GUARD(h);
    if (slot < 0 || slot >= 8) { set_error("event slot out of range"); return TDCGPU_ERR_ARG; }
    if (!c.user_events[slot]) TDC_CUDA(cudaEventCreate(&c.user_events[slot]));
    TDC_CUDA(cudaEventRecord(c.user_events[slot], c.stream));
    return 0;
}

int tdcgpu_dist_event_elapsed_ms(tdcgpu_dist* h, int slot_a, int slot_b, float* ms) {
    DIST_GUARD(h);
    if (slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !c.user_events[slot_a] || !c.user_events[slot_b] || !ms) {
        set_error("bad event slots");
        return TDCGPU_ERR_ARG;
    }
    TDC_CUDA(cudaEventSynchronize(c.user_events[slot_b]));
    TDC_CUDA(cudaEventElapsedTime(ms, c.user_events[slot_a], c.user_events[slot_b]));
    return 0;
}

int tdcgpu_dist_stats(tdcgpu_dist* h, uint64_t out[8]) {
    DIST_GUARD(h);
    out[0] = c.sa_rounds;
    out[1] = c.sa_active_sum;
    out[2] = c.sortws.stat_passes;
    out[3] = c.sortws.stat_elems;
    out[4] = c.alphabet;
    out[5] = c.symbols_per_key;
    out[6] = d.cap;
    out[7] = d.p2p ? 1 : 0;  // 1: exchanges pushed through peer memory (CUDA IPC over NVLink), 0: NCCL send/recv
    return 0;
}

int tdcgpu_dist_phase_count(tdcgpu_dist* h) { return h ? int(h->d.c.phases.size()) : 0; }
const char* tdcgpu_dist_phase_name(tdcgpu_dist* h, int i) {
    if (!h || i < 0 || i >= int(h->d.c.phases.size())) return nullptr;
    return h->d.c.phases[i].name.c_str();
}
float tdcgpu_dist_phase_ms(tdcgpu_dist* h, int i) {
    if (!h || i < 0 || i >= int(h->d.c.phases.size())) return -1.f;
    return h->d.c.phases[i].ms;
}

}  // extern "C"
