// Byte-stream stages behind the BWT in `bwt:mtf:rle:encode(huff)`, all byte/integer work:
//   mtf_encode      /root/reference/include/tudocomp/compressors/MTFCompressor.hpp:17-56
//   rle_encode      /root/reference/include/tudocomp/compressors/RunLengthEncoder.hpp:15-31 (+ util/vbyte.hpp:27-37)
//   LiteralEncoder  /root/reference/include/tudocomp/compressors/LiteralEncoder.hpp:23-32 (one code word per byte)
#pragma once
#include "encode_kernels.cuh"

namespace tdc {

// =====================================================================================================================
// Move-to-front.
//
// The reference keeps ONE table (initially 0..255), looks every byte up linearly, emits its index and moves it to the
// front.  The table at any time is: the symbols seen so far by recency of their last occurrence, followed by the
// unseen symbols in their initial (ascending) order.  Hence the effect of a block of text on ANY incoming table T is
//         T' = R ++ (T \ R),      R = the block's distinct symbols ordered by last occurrence, most recent first,
// an associative operator.  Three kernels:
//   1. mtf_tile_kernel<false>: every thread builds R for its chunk (one backward pass), one warp folds the chunks of the
//      tile (starting from 0..255) into the tile's own (R, |R|);
//   2. mtf_scan_kernel: one warp folds the tiles in order: incoming table of every tile;
//   3. mtf_tile_kernel<true>: chunk lists again, folded from the tile's incoming table, which leaves every thread the
//      exact table at the start of its chunk; then each thread runs the reference's loop on its own chunk with its own
//      table in shared memory (cost per byte ~ the emitted rank, which is small on BWT output).
// Tables live in shared memory symbol-index-major (entry k of thread t at k * MTF_THREADS + t).
// =====================================================================================================================
#ifdef TDC_CUSIM
static const u32 MTF_THREADS = 32;   // small tiles so that the CPU tests cross many chunk and tile boundaries
static const u32 MTF_CHUNK = 16;
#else
static const u32 MTF_THREADS = 256;
static const u32 MTF_CHUNK = 1024;   // bytes per thread (multiple of 16)
#endif
static const u32 MTF_TILE = MTF_THREADS * MTF_CHUNK;
static inline size_t mtf_smem_bytes() { return size_t(256) * MTF_THREADS + 512 + 64 + 2 * MTF_THREADS; }

// One warp: T := R ++ (T \ R).  T: 256 bytes (plain layout) in shared memory; this lane holds R entries
// [8 * lane, 8 * lane + 8) in r[], rc = |R|; bitmap: 8 shared words; tmp: 256 shared bytes.
__device__ __forceinline__ void mtf_compose(uint8_t* T, const uint8_t r[8], u32 rc, u32* bitmap, uint8_t* tmp) {
    const u32 lane = lane_id();
    if (rc == 0) return;
    if (lane < 8) bitmap[lane] = 0;
    __syncwarp();
#pragma unroll
    for (u32 j = 0; j < 8; j++)
        if (lane * 8 + j < rc) atomicOr(&bitmap[r[j] >> 5], 1u << (r[j] & 31u));
    __syncwarp();
    // survivors of T (not in R), in order, go behind R
    uint8_t t[8];
    u32 keep = 0;
#pragma unroll
    for (u32 j = 0; j < 8; j++) {
        t[j] = T[lane * 8 + j];
        if (!((bitmap[t[j] >> 5] >> (t[j] & 31u)) & 1u)) keep |= 1u << j;
    }
    const u32 cnt = __popc(keep);
    u32 off = warp_inclusive_sum(cnt) - cnt;
#pragma unroll
    for (u32 j = 0; j < 8; j++) {
        if (lane * 8 + j < rc) tmp[lane * 8 + j] = r[j];
        if ((keep >> j) & 1u) tmp[rc + off++] = t[j];
    }
    __syncwarp();
#pragma unroll
    for (u32 j = 0; j < 8; j++) T[lane * 8 + j] = tmp[lane * 8 + j];
    __syncwarp();
}

// APPLY == false: tile summaries (tile_R[tile][256], tile_rc[tile]);  APPLY == true: incoming tables -> output bytes.
template <bool APPLY>
static __global__ void __launch_bounds__(MTF_THREADS)
mtf_tile_kernel(const uint8_t* __restrict__ in, u64 n, uint8_t* __restrict__ tile_R, u32* __restrict__ tile_rc,
                const uint8_t* __restrict__ incoming, uint8_t* __restrict__ out) {
    TDC_DYN_SMEM(smem_raw);
    uint8_t* slots = smem_raw;                                            // [256][MTF_THREADS]: chunk lists, then chunk tables
    uint8_t* T = slots + 256 * MTF_THREADS;                               // [256] running table of the fold
    uint8_t* tmp = T + 256;                                               // [256]
    u32* bitmap = reinterpret_cast<u32*>(tmp + 256);                      // [8] (+ 8 spare)
    unsigned short* cnts = reinterpret_cast<unsigned short*>(bitmap + 16);  // [MTF_THREADS] |R| of every chunk
    const u32 t = threadIdx.x, lane = lane_id();
    const u64 tile_base = u64(blockIdx.x) * MTF_TILE;
    const u64 c0 = tile_base + u64(t) * MTF_CHUNK;                        // this thread's chunk [c0, c1)
    const u64 c1 = min(c0 + MTF_CHUNK, n);

    // ---- A. recency list of the chunk: backward pass, first sight of a symbol = its last occurrence ----
    {
        u32 seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        u32 rc = 0;
        for (u64 i = c1; i > c0 && rc < 256; i--) {
            const u32 c = in[i - 1];
            // (dynamic register indexing would go to local memory: select the word with a switch-free unrolled loop)
            u32 hit = 0;
#pragma unroll
            for (u32 w = 0; w < 8; w++) {
                const u32 bit = (w == (c >> 5)) ? (1u << (c & 31u)) : 0u;
                hit |= seen[w] & bit;
                seen[w] |= bit;
            }
            if (!hit) { slots[rc * MTF_THREADS + t] = uint8_t(c); rc++; }
        }
        cnts[t] = (unsigned short)rc;
    }
    __syncthreads();

    // ---- B. one warp folds the chunks in order ----
    if (warp_id() == 0) {
#pragma unroll
        for (u32 j = 0; j < 8; j++) T[lane * 8 + j] = APPLY ? incoming[u64(blockIdx.x) * 256 + lane * 8 + j] : uint8_t(lane * 8 + j);
        u32 any[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // (lane 0 only) not needed: |R_tile| is counted from a bitmap below
        (void)any;
        if (!APPLY && lane < 8) bitmap[8 + lane] = 0;  // union of the chunks' symbol sets
        __syncwarp();
        for (u32 k = 0; k < MTF_THREADS; k++) {
            const u32 rc = cnts[k];
            uint8_t r[8];
#pragma unroll
            for (u32 j = 0; j < 8; j++) r[j] = (lane * 8 + j < rc) ? slots[(lane * 8 + j) * MTF_THREADS + k] : uint8_t(0);
            __syncwarp();
            if (APPLY) {
                // the table at the start of chunk k replaces the chunk's list in its slot
#pragma unroll
                for (u32 j = 0; j < 8; j++) slots[(lane * 8 + j) * MTF_THREADS + k] = T[lane * 8 + j];
            } else {
#pragma unroll
                for (u32 j = 0; j < 8; j++)
                    if (lane * 8 + j < rc) atomicOr(&bitmap[8 + (r[j] >> 5)], 1u << (r[j] & 31u));
            }
            __syncwarp();
            mtf_compose(T, r, rc, bitmap, tmp);
        }
        if (!APPLY) {
#pragma unroll
            for (u32 j = 0; j < 8; j++) tile_R[u64(blockIdx.x) * 256 + lane * 8 + j] = T[lane * 8 + j];
            u32 c = lane < 8 ? u32(__popc(bitmap[8 + lane])) : 0u;
            c = warp_sum(c);
            if (lane == 0) tile_rc[blockIdx.x] = c;
        }
    }
    if (!APPLY) return;
    __syncthreads();

    // ---- C. the reference's loop (MTFCompressor.hpp:17-30) on this thread's chunk with its own table ----
    u64 i = c0;
    while (i < c1) {
        // 16 bytes at a time while they are all inside the chunk (chunks start at multiples of 16)
        const u32 m = u32(min(u64(16), c1 - i));
        uint8_t src[16], dst[16];
        if (m == 16) {
            *reinterpret_cast<uint4*>(src) = *reinterpret_cast<const uint4*>(in + i);
        } else {
            for (u32 j = 0; j < m; j++) src[j] = in[i + j];
        }
#pragma unroll
        for (u32 j = 0; j < 16; j++) {
            if (j < m) {
                const uint8_t c = src[j];
                u32 k = 0;
                uint8_t prev = slots[t];
                if (prev != c) {
                    // find c, shifting the entries in front of it down by one on the way
                    do {
                        k++;
                        const uint8_t cur = slots[k * MTF_THREADS + t];
                        slots[k * MTF_THREADS + t] = prev;
                        prev = cur;
                    } while (prev != c);
                    slots[t] = c;
                }
                dst[j] = uint8_t(k);
            }
        }
        if (m == 16) {
            *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(dst);
        } else {
            for (u32 j = 0; j < m; j++) out[i + j] = dst[j];
        }
        i += m;
    }
}

// one warp: incoming table of every tile
static __global__ void __launch_bounds__(32)
mtf_scan_kernel(const uint8_t* __restrict__ tile_R, const u32* __restrict__ tile_rc, u32 ntiles, uint8_t* __restrict__ incoming) {
    __shared__ __align__(16) uint8_t T[256];
    __shared__ __align__(16) uint8_t tmp[256];
    __shared__ u32 bitmap[8];
    const u32 lane = lane_id();
#pragma unroll
    for (u32 j = 0; j < 8; j++) T[lane * 8 + j] = uint8_t(lane * 8 + j);
    __syncwarp();
    for (u32 k = 0; k < ntiles; k++) {
        uint8_t r[8];
#pragma unroll
        for (u32 j = 0; j < 8; j++) {
            incoming[u64(k) * 256 + lane * 8 + j] = T[lane * 8 + j];
            r[j] = tile_R[u64(k) * 256 + lane * 8 + j];  // the first tile_rc[k] entries are the tile's recency list
        }
        __syncwarp();
        mtf_compose(T, r, tile_rc[k], bitmap, tmp);
    }
}

}  // namespace tdc
