// Device side of lzss::encode_text (/root/reference/include/tudocomp/compressors/lzss/LZSSCoding.hpp:18-92) for coders
// that write integers as plain binary of bits_for(range) bits (the Encoder default, Coder.hpp:63-80) and literals as one
// fixed code word per byte value (BitCoder: the byte itself; HuffmanCoder: coders/HuffmanCoder.hpp:309-322).
//
// The reference walks the factor list once and emits, MSB first (io/BitOStream.hpp:79-102):
//     n:32  flen_min:bn  flen_max:bn  fdist_max:bn                                   (bn = bits_for(n))
//     per factor:  0                       | 1 count:bf  literal codes ...           (bf = bits_for(fdist_max))
//                  src:bn  (len - flen_min):bl                                       (bl = bits_for(flen_max - flen_min))
//     tail:        1 count:bf  literal codes ...      (always present: the sentinel is never inside a factor)
// Here the same stream is produced in text order.  Two bit masks over the text positions are derived from the factor
// records: S (a factor starts here) and E (a factor ended just before here).  Position i then contributes
//     [ 1 + (S(i) ? 0 : bf) bits ]   if i == 0 or E(i): the cursor p of the reference loop stands at i when the next
//                                    item begins, so the flag (and the literal count up to the next S bit) go here
//     [ bn + bl bits ]               if S(i)                       (the factor's src and len)
//     [ code(T[i]) ]                 else if i is not covered      (a literal)
// and nothing when it lies inside a factor.  A tile of ENC_TILE positions sums its bit lengths, an exclusive scan over
// the tiles gives every tile its bit offset, and each tile assembles its bits in shared memory and writes whole
// 32-bit words (the two boundary words are OR-ed into the zeroed output).
#pragma once
#include "tdc_ctx.h"

namespace tdc {

static const u32 ENC_THREADS = 256;
static const u32 ENC_PPT = 8;                           // positions per thread = one byte of each bit mask
static const u32 ENC_TILE = ENC_THREADS * ENC_PPT;      // 2048 positions = 64 mask words
static const u32 ENC_MASK_WORDS = ENC_TILE / 32;
static const u32 ENC_MAX_BITS_PER_POS = 1 + 32 + 64;    // flag + literal count, then a literal code or src + len
static const u32 ENC_SMEM_WORDS = (ENC_TILE * ENC_MAX_BITS_PER_POS + 31) / 32 + 2;

struct EncParams {
    u32 n, z;
    u32 bn, bf, bl;  // bit widths of text positions, literal counts, factor lengths
    u32 flen_min;
};

__device__ __forceinline__ u32 enc_bswap(u32 v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

// OR the low `nbits` (1..32) bits of `value`, MSB first, into the big-endian bit string `w` at bit position `pos`.
// SHARED: the words are shared memory touched by other threads (atomicOr); otherwise thread-private.
template <bool SHARED>
__device__ __forceinline__ void enc_put(u32* w, u32 pos, u32 value, u32 nbits) {
    if (nbits < 32) value &= (1u << nbits) - 1u;
    const u32 o = pos & 31u;
    const u64 v64 = u64(value) << (64u - o - nbits);  // o + nbits <= 63
    const u32 hi = u32(v64 >> 32), lo = u32(v64);
    u32* p = w + (pos >> 5);
    if (SHARED) {
        if (hi) atomicOr(p, hi);
        if (lo) atomicOr(p + 1, lo);
    } else {
        p[0] |= hi;
        if (lo) p[1] |= lo;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// factor records -> S / E masks, longest literal run (fdist_max of encode_text, LZSSCoding.hpp:29-41)
// ---------------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
enc_mark_kernel(const Factor* __restrict__ f, u32 z, u32 n, u32* __restrict__ S, u32* __restrict__ E, u32* __restrict__ fdist_max) {
    const u64 k = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    u32 gap = 0;
    if (k < z) {
        const Factor x = f[k];
        const u32 end = x.pos + x.len;
        atomicOr(&S[x.pos >> 5], 1u << (x.pos & 31u));
        atomicOr(&E[end >> 5], 1u << (end & 31u));
        u32 prev_end = 0;
        if (k) { const Factor y = f[k - 1]; prev_end = y.pos + y.len; }
        gap = x.pos - prev_end;
        if (k == u64(z) - 1) gap = max(gap, n - end);
    }
    gap = warp_max(gap);
    if (lane_id() == 0 && gap) atomicMax(fdist_max, gap);
}

// per tile: number of S bits and of E bits (one warp per tile)
static __global__ void __launch_bounds__(256)
enc_tile_popc_kernel(const u32* __restrict__ S, const u32* __restrict__ E, u32 ntiles, u32* __restrict__ tile_s, u32* __restrict__ tile_e) {
    const u32 tile = blockIdx.x * (blockDim.x / 32) + warp_id();
    if (tile >= ntiles) return;
    const u64 wbase = u64(tile) * ENC_MASK_WORDS;
    u32 cs = 0, ce = 0;
    for (u32 j = lane_id(); j < ENC_MASK_WORDS; j += 32) {
        cs += __popc(S[wbase + j]);
        ce += __popc(E[wbase + j]);
    }
    cs = warp_sum(cs);
    ce = warp_sum(ce);
    if (lane_id() == 0) { tile_s[tile] = cs; tile_e[tile] = ce; }
}

// single CTA: exclusive scan of u32 counts into offsets of type O starting at `first`; *total = first + sum
template <class O>
static __global__ void __launch_bounds__(1024)
enc_scan_kernel(const u32* cnt, u32 ntiles, O first, O* off, O* total) {  // cnt may alias off (in place)
    __shared__ O scratch[33];
    O carry = first;
    for (u32 b = 0; b < ntiles; b += 1024) {
        const u32 i = b + threadIdx.x;
        const O c = i < ntiles ? O(cnt[i]) : O(0);
        O tot;
        const O ex = block_exclusive_sum<O>(c, scratch, &tot);
        if (i < ntiles) off[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}

// ---------------------------------------------------------------------------------------------------------------
// the tile kernel: MODE 0 = histogram of the literals (lzss::TextLiterals, lzss/LZSSLiterals.hpp:10-56, as counted
// by huff::count_alphabet_literals, coders/HuffmanCoder.hpp:37-49), MODE 1 = bits per tile, MODE 2 = write the bits
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
static __global__ void __launch_bounds__(ENC_THREADS)
enc_tile_kernel(const uint8_t* __restrict__ text, const Factor* __restrict__ f, const u32* __restrict__ S, const u32* __restrict__ E,
                const u32* __restrict__ scan_s, const u32* __restrict__ scan_e, EncParams P,
                const u64* __restrict__ codes, const uint8_t* __restrict__ lens,  // MODE 1, 2: literal code table (device)
                ull* __restrict__ hist,                                           // MODE 0: 256 global counters
                u32* __restrict__ tile_bits,                                      // MODE 1 out
                const u64* __restrict__ tile_off, u32* __restrict__ out32) {      // MODE 2
    __shared__ u32 scratch[33];
    __shared__ u32 s_hist[MODE == 0 ? (ENC_THREADS / 32) * 256 : 1];
    __shared__ u64 s_code[MODE == 2 ? 256 : 1];
    __shared__ uint8_t s_len[MODE == 0 ? 1 : 256];
    __shared__ u32 s_bits[MODE == 2 ? ENC_SMEM_WORDS : 1];

    const u32 tile = blockIdx.x, t = threadIdx.x;
    if constexpr (MODE == 0) {
        for (u32 j = t; j < (ENC_THREADS / 32) * 256; j += ENC_THREADS) s_hist[j] = 0;
    } else {
        s_len[t] = lens[t];
        if constexpr (MODE == 2) s_code[t] = codes[t];
    }
    const u32 sS = reinterpret_cast<const uint8_t*>(S)[u64(tile) * (ENC_TILE / 8) + t];
    const u32 sE = reinterpret_cast<const uint8_t*>(E)[u64(tile) * (ENC_TILE / 8) + t];
    const u64 pos0 = u64(tile) * ENC_TILE + u64(t) * ENC_PPT;
    u64 t8 = 0;
    if (pos0 < P.n) t8 = *reinterpret_cast<const u64*>(text + pos0);  // the text buffer is padded past n
    // S and E bits in front of this thread's first position (both counts stay below 2^16 inside a tile)
    u32 tot;
    const u32 pre = block_exclusive_sum<u32>(u32(__popc(sS)) | (u32(__popc(sE)) << 16), scratch, &tot);  // also orders the smem init
    u32 srank = scan_s[tile] + (pre & 0xffffu);                      // factors starting before pos0
    u32 inside = srank - (scan_e[tile] + (pre >> 16));               // 1: pos0 - 1 lies inside a factor that may go on

    // ---- pass A: bit lengths (MODE 1, 2) or literal counts (MODE 0) ----
    u32 my_bits = 0;
    {
        u32 in = inside;
#pragma unroll
        for (u32 j = 0; j < ENC_PPT; j++) {
            const u64 i = pos0 + j;
            if (i >= P.n) break;
            const u32 e = (sE >> j) & 1u, s = (sS >> j) & 1u;
            if (e) in = 0;
            if (s) in = 1;
            const u32 c = u32(t8 >> (8 * j)) & 0xffu;
            if constexpr (MODE == 0) {
                if (!in) atomicAdd(&s_hist[warp_id() * 256 + c], 1u);
            } else {
                if (e || i == 0) my_bits += 1 + (s ? 0 : P.bf);
                if (s) my_bits += P.bn + P.bl;
                else if (!in) my_bits += s_len[c];
            }
        }
    }
    if constexpr (MODE == 0) {
        __syncthreads();
        u32 sum = 0;
#pragma unroll
        for (u32 w = 0; w < ENC_THREADS / 32; w++) sum += s_hist[w * 256 + t];
        if (sum) atomicAdd(&hist[t], ull(sum));
        return;
    }
    u32 tile_total;
    const u32 my_off = block_exclusive_sum<u32>(my_bits, scratch, &tile_total);
    if constexpr (MODE == 1) {
        if (t == 0) tile_bits[tile] = tile_total;
        return;
    }
    if constexpr (MODE == 2) {
        const u64 g0 = tile_off[tile];            // global bit offset of this tile
        const u32 shift0 = u32(g0 & 31u);         // the shared words are aligned with the output words
        const u32 nwords = (shift0 + tile_total + 31) / 32;
        for (u32 j = t; j < nwords + 1; j += ENC_THREADS) s_bits[j] = 0;  // +1: enc_put may touch the next word with 0 bits
        __syncthreads();
        // ---- pass B: emit ----
        u32 cur = shift0 + my_off;
        u32 in = inside;
#pragma unroll
        for (u32 j = 0; j < ENC_PPT; j++) {
            const u64 i = pos0 + j;
            if (i >= P.n) break;
            const u32 e = (sE >> j) & 1u, s = (sS >> j) & 1u;
            if (e) in = 0;
            if (e || i == 0) {
                // LZSSCoding.hpp:58-68 / :82-85: flag, then the number of literals up to the next factor (or the end)
                enc_put<true>(s_bits, cur, s ? 0u : 1u, 1);
                cur += 1;
                if (!s) {
                    const u32 next_pos = srank < P.z ? f[srank].pos : P.n;
                    enc_put<true>(s_bits, cur, next_pos - u32(i), P.bf);
                    cur += P.bf;
                }
            }
            if (s) {
                // LZSSCoding.hpp:76-78: src in bits_for(n) bits, len - flen_min in bits_for(flen_max - flen_min) bits
                const Factor x = f[srank++];
                in = 1;
                enc_put<true>(s_bits, cur, x.src, P.bn);
                cur += P.bn;
                enc_put<true>(s_bits, cur, x.len - P.flen_min, P.bl);
                cur += P.bl;
            } else if (!in) {
                const u32 c = u32(t8 >> (8 * j)) & 0xffu;
                const u32 L = s_len[c];
                const u64 code = s_code[c];
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
        // ---- flush: whole words; the first and the last word may be shared with the neighbouring tiles ----
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

// the stream head: `lead_bits` bits already in the coder's current byte, then n and the three header values
// (LZSSCoding.hpp:46-50).  One thread.
// len_field_bits: 32, or 64 for the wide-index build of the reference (LengthRange over a 64-bit len_t; n < 2^32 here, so
// the upper word is zero).
static __global__ void enc_header_kernel(u32 lead_bits, u32 lead_byte, EncParams P, u32 flen_max, u32 fdist_max, u32 len_field_bits,
                                         u32* __restrict__ out32) {
    if (threadIdx.x || blockIdx.x) return;
    u32 w[7] = {0, 0, 0, 0, 0, 0, 0};
    u32 cur = 0;
    if (lead_bits) { enc_put<false>(w, cur, lead_byte >> (8 - lead_bits), lead_bits); cur += lead_bits; }
    if (len_field_bits == 64) cur += 32;
    enc_put<false>(w, cur, P.n, 32); cur += 32;
    enc_put<false>(w, cur, P.flen_min, P.bn); cur += P.bn;
    enc_put<false>(w, cur, flen_max, P.bn); cur += P.bn;
    enc_put<false>(w, cur, fdist_max, P.bn); cur += P.bn;
    for (u32 j = 0; j * 32 < cur; j++)
        if (w[j]) atomicOr(out32 + j, enc_bswap(w[j]));
}

}  // namespace tdc
