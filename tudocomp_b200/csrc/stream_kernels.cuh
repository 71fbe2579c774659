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
//   1. mtf_tile_kernel<false>: every thread builds R for its chunk (one backward pass); every warp folds its 32 chunks
//      (starting from 0..255), one warp folds the warps' results into the tile's own (R, |R|);
//   2. mtf_scan_kernel: one warp folds the tiles in order: incoming table of every tile;
//   3. mtf_tile_kernel<true>: chunk lists and warp results again; the warp results folded from the tile's incoming table
//      give every warp its incoming table, from which it folds its chunks once more — that leaves every thread the exact
//      table at the start of its chunk (32 + warps + 32 sequential fold steps instead of one per chunk of the tile);
//      then each thread runs the reference's loop on its own chunk with its own table in shared memory (cost per byte
//      ~ the emitted rank, which is small on BWT output).
// Tables live in shared memory symbol-index-major (entry k of thread t at k * MTF_THREADS + t).
// =====================================================================================================================
#ifdef TDC_CUSIM
static const u32 MTF_THREADS = 64;   // small tiles so that the CPU tests cross many chunk, warp and tile boundaries
static const u32 MTF_CHUNK = 16;
#else
static const u32 MTF_THREADS = 256;
static const u32 MTF_CHUNK = 1024;   // bytes per thread (multiple of 16)
#endif
static const u32 MTF_TILE = MTF_THREADS * MTF_CHUNK;
static const u32 MTF_WARPS = MTF_THREADS / 32;
static const u32 MTF_FOLD_BYTES = 256 + 256 + 64;  // table, scratch table, two 8-word bitmaps: one fold context
static inline size_t mtf_smem_bytes() { return size_t(256) * MTF_THREADS + size_t(MTF_WARPS + 1) * MTF_FOLD_BYTES + 2 * MTF_THREADS + 4 * MTF_WARPS + 64; }

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
                const uint8_t* __restrict__ incoming, uint8_t* __restrict__ out, bool vec16) {  // vec16: in and out are 16-byte aligned
    TDC_DYN_SMEM(smem_raw);
    uint8_t* slots = smem_raw;                                            // [256][MTF_THREADS]: chunk lists, then chunk tables
    uint8_t* folds = slots + 256 * MTF_THREADS;                           // [MTF_WARPS + 1] fold contexts (last: the tile's)
    unsigned short* cnts = reinterpret_cast<unsigned short*>(folds + (MTF_WARPS + 1) * MTF_FOLD_BYTES);  // [MTF_THREADS] |R| of every chunk
    u32* warp_rc = reinterpret_cast<u32*>(cnts + MTF_THREADS);            // [MTF_WARPS] distinct symbols of every warp's chunks
    const u32 t = threadIdx.x, lane = lane_id(), w = warp_id();
    uint8_t* T = folds + w * MTF_FOLD_BYTES;                              // this warp's running table
    uint8_t* tmp = T + 256;
    u32* bitmap = reinterpret_cast<u32*>(tmp + 256);                      // [8] scratch of mtf_compose, [8] union of symbol sets
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

    // ---- B1. every warp folds its own 32 chunks from 0..255: (R_w, |R_w|) = the warp's effect on any table ----
    {
#pragma unroll
        for (u32 j = 0; j < 8; j++) T[lane * 8 + j] = uint8_t(lane * 8 + j);
        if (lane < 8) bitmap[8 + lane] = 0;
        __syncwarp();
        for (u32 k = w * 32; k < w * 32 + 32; k++) {
            const u32 rc = cnts[k];
            uint8_t r[8];
#pragma unroll
            for (u32 j = 0; j < 8; j++) {
                r[j] = (lane * 8 + j < rc) ? slots[(lane * 8 + j) * MTF_THREADS + k] : uint8_t(0);
                if (lane * 8 + j < rc) atomicOr(&bitmap[8 + (r[j] >> 5)], 1u << (r[j] & 31u));
            }
            __syncwarp();
            mtf_compose(T, r, rc, bitmap, tmp);
        }
        u32 c = lane < 8 ? u32(__popc(bitmap[8 + lane])) : 0u;
        c = warp_sum(c);
        if (lane == 0) warp_rc[w] = c;
    }
    __syncthreads();
    // ---- B2. one warp folds the warps' results in order (from the tile's incoming table when applying) ----
    if (w == 0) {
        uint8_t* TT = folds + MTF_WARPS * MTF_FOLD_BYTES;
        uint8_t* ttmp = TT + 256;
        u32* tbitmap = reinterpret_cast<u32*>(ttmp + 256);
#pragma unroll
        for (u32 j = 0; j < 8; j++) TT[lane * 8 + j] = APPLY ? incoming[u64(blockIdx.x) * 256 + lane * 8 + j] : uint8_t(lane * 8 + j);
        if (lane < 8) tbitmap[8 + lane] = 0;
        __syncwarp();
        for (u32 ww = 0; ww < MTF_WARPS; ww++) {
            uint8_t* Tw = folds + ww * MTF_FOLD_BYTES;
            const u32 rc = warp_rc[ww];
            uint8_t r[8];
#pragma unroll
            for (u32 j = 0; j < 8; j++) r[j] = Tw[lane * 8 + j];  // its first rc entries are the warp's recency list
            if (!APPLY && lane < 8) tbitmap[8 + lane] |= reinterpret_cast<const u32*>(Tw + 512)[8 + lane];
            __syncwarp();
            if (APPLY) {
                // the table at the start of warp ww's chunks replaces the warp's result
#pragma unroll
                for (u32 j = 0; j < 8; j++) Tw[lane * 8 + j] = TT[lane * 8 + j];
                __syncwarp();
            }
            mtf_compose(TT, r, rc, tbitmap, ttmp);
        }
        if (!APPLY) {
#pragma unroll
            for (u32 j = 0; j < 8; j++) tile_R[u64(blockIdx.x) * 256 + lane * 8 + j] = TT[lane * 8 + j];
            u32 c = lane < 8 ? u32(__popc(tbitmap[8 + lane])) : 0u;
            c = warp_sum(c);
            if (lane == 0) tile_rc[blockIdx.x] = c;
        }
    }
    if (!APPLY) return;
    __syncthreads();
    // ---- B3. every warp folds its chunks again, now from its incoming table: the table at the start of chunk k replaces
    //          the chunk's list in its slot ----
    for (u32 k = w * 32; k < w * 32 + 32; k++) {
        const u32 rc = cnts[k];
        uint8_t r[8];
#pragma unroll
        for (u32 j = 0; j < 8; j++) r[j] = (lane * 8 + j < rc) ? slots[(lane * 8 + j) * MTF_THREADS + k] : uint8_t(0);
        __syncwarp();
#pragma unroll
        for (u32 j = 0; j < 8; j++) slots[(lane * 8 + j) * MTF_THREADS + k] = T[lane * 8 + j];
        __syncwarp();
        mtf_compose(T, r, rc, bitmap, tmp);
    }
    __syncthreads();

    // ---- C. the reference's loop (MTFCompressor.hpp:17-30) on this thread's chunk with its own table ----
    u64 i = c0;
    while (i < c1) {
        // 16 bytes at a time while they are all inside the chunk (chunks start at multiples of 16)
        const u32 m = u32(min(u64(16), c1 - i));
        __align__(16) uint8_t src[16];
        __align__(16) uint8_t dst[16];
        if (m == 16 && vec16) {
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
                    } while (prev != c && k < 255);  // (the table is a permutation; the bound only guards against a hang)
                    slots[t] = c;
                }
                dst[j] = uint8_t(k);
            }
        }
        if (m == 16 && vec16) {
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

// =====================================================================================================================
// Run-length encoding (RunLengthEncoder.hpp:15-31).  A maximal run of L equal bytes c < 0x80 becomes  c  (L == 1)  or
// c c vbyte(L - 2 + offset)  (L >= 2).  In text order: position i emits its byte if it is the head of a run
// (i == 0 or in[i] != in[i-1]) or the second byte of one; the second byte is followed by the vbyte, whose value needs the
// next head after i: found inside the tile, else taken from a suffix-minimum over the tiles' first heads.
// Bytes >= 0x80: the reference compares `is.peek() == c` with c a signed char (:24), which never holds for them, so
// their runs are not merged: every non-head byte is written as  c vbyte(offset)  (see oracle/tdc_oracle.c).
// =====================================================================================================================
static const u32 RLE_THREADS = 256;
static const u32 RLE_PPT = 8;
static const u32 RLE_TILE = RLE_THREADS * RLE_PPT;
static const u32 RLE_NONE = 0xffffffffu;
static const u32 RLE_MAX_TILE_BYTES = RLE_TILE * 11 + 16;  // every byte followed by a 10-byte vbyte

__device__ __forceinline__ u32 vbyte_len(u64 v) {  // util/vbyte.hpp:27-37: 7 data bits per byte
    u32 l = 1;
    while (v >>= 7) l++;
    return l;
}

// bytes [p0 - 2, p0 + 8) of the input as 10 values (0x100 = before the start / past the end)
__device__ __forceinline__ void rle_load10(const uint8_t* __restrict__ in, u64 n, u64 p0, u32 b[10]) {
#pragma unroll
    for (u32 j = 0; j < 10; j++) {
        const u64 q = p0 + j;  // index + 2
        b[j] = (q >= 2 && q - 2 < n) ? u32(in[q - 2]) : 0x100u;
    }
}

// per tile: position of its first run head (RLE_NONE if the tile lies inside one run)
static __global__ void __launch_bounds__(RLE_THREADS)
rle_first_head_kernel(const uint8_t* __restrict__ in, u64 n, u32* __restrict__ tile_first) {
    __shared__ u32 s_min;
    if (threadIdx.x == 0) s_min = RLE_NONE;
    __syncthreads();
    const u64 p0 = u64(blockIdx.x) * RLE_TILE + u64(threadIdx.x) * RLE_PPT;
    u32 first = RLE_NONE;
    if (p0 < n) {
        u32 b[10];
        rle_load10(in, n, p0, b);
#pragma unroll
        for (u32 j = 0; j < RLE_PPT; j++) {
            const u64 i = p0 + j;
            if (i < n && first == RLE_NONE && (i == 0 || b[j + 2] != b[j + 1])) first = u32(i);
        }
    }
    first = warp_min(first);
    if (lane_id() == 0 && first != RLE_NONE) atomicMin(&s_min, first);
    __syncthreads();
    if (threadIdx.x == 0) tile_first[blockIdx.x] = s_min;
}

// single CTA, tiles from the back: next_after[t] = first head in a tile > t (n if there is none).  A suffix minimum, done
// as a prefix maximum of ~position over the reversed tile order (RLE_NONE maps to the identity 0).
static __global__ void __launch_bounds__(1024)
rle_next_head_kernel(const u32* __restrict__ tile_first, u32 ntiles, u32 n, u32* __restrict__ next_after) {
    __shared__ u32 scratch[33];
    __shared__ u32 s_inc[1024];
    u32 carry = 0xffffffffu - n;
    for (u32 b = 0; b < ntiles; b += 1024) {
        const u32 r = b + threadIdx.x;  // reversed index
        const u32 f = r < ntiles ? tile_first[ntiles - 1 - r] : RLE_NONE;
        u32 tot;
        const u32 inc = block_inclusive_max(0xffffffffu - f, scratch, &tot);
        s_inc[threadIdx.x] = inc;
        __syncthreads();
        const u32 ex = threadIdx.x ? s_inc[threadIdx.x - 1] : 0u;  // tiles strictly behind this one inside the batch
        if (r < ntiles) next_after[ntiles - 1 - r] = 0xffffffffu - max(carry, ex);
        carry = max(carry, tot);
        __syncthreads();
    }
}

// MODE 1: output bytes per tile; MODE 2: write
template <int MODE>
static __global__ void __launch_bounds__(RLE_THREADS)
rle_tile_kernel(const uint8_t* __restrict__ in, u64 n, u64 offset, const u32* __restrict__ next_after, u32* __restrict__ tile_bytes,
                const u64* __restrict__ tile_off, uint8_t* __restrict__ out) {
    __shared__ u32 scratch[33];
    __shared__ u32 s_first[RLE_THREADS + 1];
    __shared__ uint8_t s_out[MODE == 2 ? RLE_MAX_TILE_BYTES : 1];
    const u32 t = threadIdx.x;
    const u64 p0 = u64(blockIdx.x) * RLE_TILE + u64(t) * RLE_PPT;
    u32 b[10];
    rle_load10(in, n, p0, b);
    u32 heads = 0, seconds = 0, first = RLE_NONE;
#pragma unroll
    for (u32 j = 0; j < RLE_PPT; j++) {
        const u64 i = p0 + j;
        if (i < n) {
            const bool h = i == 0 || b[j + 2] != b[j + 1];
            const bool hp = i >= 1 && (i == 1 || b[j + 1] != b[j]);  // head(i - 1)
            if (h) { heads |= 1u << j; if (first == RLE_NONE) first = u32(i); }
            if (!h && (hp || b[j + 2] >= 0x80u)) seconds |= 1u << j;  // bytes >= 0x80: every non-head byte restarts a "run"
        }
    }
    // next head behind this thread's positions: suffix minimum over the later threads, then the later tiles
    s_first[t] = first;
    if (t == 0) s_first[RLE_THREADS] = next_after[blockIdx.x];
    __syncthreads();
    for (u32 d = 1; d <= RLE_THREADS; d <<= 1) {  // Hillis-Steele suffix-min (positions grow with the index: min = nearest)
        const u32 v = (t + d <= RLE_THREADS) ? s_first[t + d] : RLE_NONE;
        __syncthreads();
        s_first[t] = min(s_first[t], v);
        __syncthreads();
    }
    const u32 after = t + 1 <= RLE_THREADS ? s_first[t + 1] : RLE_NONE;  // first head in threads > t or in later tiles (or n)
    u32 my = 0;
    u32 vlen[RLE_PPT];
    u64 vval[RLE_PPT];
#pragma unroll
    for (u32 j = 0; j < RLE_PPT; j++) {
        vlen[j] = 0;
        vval[j] = 0;
        if ((heads >> j) & 1u) my += 1;
        if ((seconds >> j) & 1u) {
            const u32 above = heads & ~((2u << j) - 1u);  // heads at positions > j inside this thread
            const u64 nh = above ? p0 + (__ffs(int(above)) - 1) : u64(after);
            const u64 L = b[j + 2] >= 0x80u ? 2 : nh - (p0 + j - 1);  // unmerged runs count no further bytes
            vval[j] = L - 2 + offset;
            vlen[j] = vbyte_len(vval[j]);
            my += 1 + vlen[j];
        }
    }
    u32 tile_total;
    u32 o = block_exclusive_sum<u32>(my, scratch, &tile_total);
    if (MODE == 1) {
        if (t == 0) tile_bytes[blockIdx.x] = tile_total;
        return;
    }
    if (MODE == 2) {
#pragma unroll
        for (u32 j = 0; j < RLE_PPT; j++) {
            if (((heads | seconds) >> j) & 1u) s_out[o++] = uint8_t(b[j + 2]);
            if ((seconds >> j) & 1u) {
                u64 v = vval[j];
                for (u32 k = 0; k < vlen[j]; k++) {
                    uint8_t byte = uint8_t(v & 0x7f);
                    v >>= 7;
                    if (v > 0) byte |= 0x80;
                    s_out[o++] = byte;
                }
            }
        }
        __syncthreads();
        uint8_t* dst = out + tile_off[blockIdx.x];
        for (u32 k = t; k < tile_total; k += RLE_THREADS) dst[k] = s_out[k];
    }
}

// =====================================================================================================================
// LiteralEncoder (LiteralEncoder.hpp:23-32): every byte by its code word, MSB first.  Same tile scheme as the lzss
// encoder (encode_kernels.cuh) without factors.
// =====================================================================================================================
static __global__ void __launch_bounds__(256) stream_histogram_kernel(const uint8_t* __restrict__ in, u64 n, ull* __restrict__ hist) {
    __shared__ u32 sh[8 * 256];
    for (u32 j = threadIdx.x; j < 8 * 256; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    const u64 stride = u64(gridDim.x) * blockDim.x;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&sh[warp_id() * 256 + in[i]], 1u);
    __syncthreads();
    if (threadIdx.x < 256) {
        u32 sum = 0;
        for (u32 w = 0; w < 8; w++) sum += sh[w * 256 + threadIdx.x];
        if (sum) atomicAdd(&hist[threadIdx.x], ull(sum));
    }
}

static const u32 LIT_SMEM_WORDS = ENC_TILE * 64 / 32 + 2;

template <int MODE>  // 1: bits per tile, 2: write
static __global__ void __launch_bounds__(ENC_THREADS)
lit_tile_kernel(const uint8_t* __restrict__ in, u64 n, const u64* __restrict__ codes, const uint8_t* __restrict__ lens,
                u32* __restrict__ tile_bits, const u64* __restrict__ tile_off, u32* __restrict__ out32) {
    __shared__ u32 scratch[33];
    __shared__ u64 s_code[MODE == 2 ? 256 : 1];
    __shared__ uint8_t s_len[256];
    __shared__ u32 s_bits[MODE == 2 ? LIT_SMEM_WORDS : 1];
    const u32 t = threadIdx.x;
    s_len[t] = lens[t];
    if (MODE == 2) s_code[t] = codes[t];
    __syncthreads();
    const u64 p0 = u64(blockIdx.x) * ENC_TILE + u64(t) * ENC_PPT;
    u32 c[ENC_PPT], my = 0;
#pragma unroll
    for (u32 j = 0; j < ENC_PPT; j++) {
        c[j] = p0 + j < n ? u32(in[p0 + j]) : 0x100u;
        if (c[j] < 256) my += s_len[c[j]];
    }
    u32 tile_total;
    const u32 off = block_exclusive_sum<u32>(my, scratch, &tile_total);
    if (MODE == 1) {
        if (t == 0) tile_bits[blockIdx.x] = tile_total;
        return;
    }
    if (MODE == 2) {
        const u64 g0 = tile_off[blockIdx.x];
        const u32 shift0 = u32(g0 & 31u);
        const u32 nwords = (shift0 + tile_total + 31) / 32;
        for (u32 j = t; j < nwords + 1; j += ENC_THREADS) s_bits[j] = 0;
        __syncthreads();
        u32 cur = shift0 + off;
#pragma unroll
        for (u32 j = 0; j < ENC_PPT; j++) {
            if (c[j] < 256) {
                const u32 L = s_len[c[j]];
                const u64 code = s_code[c[j]];
                if (L > 32) {
                    enc_put<true>(s_bits, cur, u32(code >> 32), L - 32);
                    enc_put<true>(s_bits, cur + (L - 32), u32(code), 32);
                } else if (L) {
                    enc_put<true>(s_bits, cur, u32(code), L);
                }
                cur += L;
            }
        }
        __syncthreads();
        u32* o = out32 + (g0 >> 5);
        const bool last_partial = ((shift0 + tile_total) & 31u) != 0;
        for (u32 j = t; j < nwords; j += ENC_THREADS) {
            const u32 v = enc_bswap(s_bits[j]);
            if ((j == 0 && shift0) || (j == nwords - 1 && last_partial)) {
                if (v) atomicOr(o + j, v);
            } else {
                o[j] = v;
            }
        }
    }
}

// the `lead_bits` header bits already in the coder's current byte
static __global__ void lit_header_kernel(u32 lead_bits, u32 lead_byte, u32* __restrict__ out32) {
    if (threadIdx.x || blockIdx.x || !lead_bits) return;
    atomicOr(out32, enc_bswap((lead_byte >> (8 - lead_bits)) << (32 - lead_bits)));
}

}  // namespace tdc
