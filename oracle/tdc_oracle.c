/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's text-index + lzss_lcp hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.  The product
 * (tudocomp_b200 / libtdcgpu.so) never links, imports or executes it; there is no CPU fallback in the product.
 *
 * Parity status: PINNED.  Every function here is checked in tests/test_oracle.py against
 *   (a) the known answers recorded from the reference ("abcdebcdeabc\0", "banana\0"),
 *   (b) the reference itself (oracle/_ref/libtdcref.so = the unmodified headers under /root/reference compiled by
 *       oracle/Makefile) on the reference's own test strings (test/test/util.hpp:98-207) and on seeded synthetic inputs,
 *   (c) golden fixtures under tests/golden/ generated from (b) by tests/golden/make_golden.py.
 *
 * All paths below are relative to /root/reference/.  Indices are uint32_t (len_t, include/tudocomp/def.hpp:103,114).
 * The text always carries exactly one 0 byte, at position n-1 (include/tudocomp/ds/TextDS.hpp:132-138).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------------
 * SA — contract of SADivSufSort (include/tudocomp/ds/SADivSufSort.hpp:28-51): SA[r] = start of the r-th smallest suffix
 * of T[0..n), bytes compared UNSIGNED (include/tudocomp/util/divsufsort/divsufsort_def.hpp:12).  divsufsort's induced
 * sorting is not restated; the result is a function of the text, so any correct suffix sorter is bit-exact.
 * Here: prefix doubling with qsort (Manber–Myers / Larsson–Sadakane flavour), O(n log^2 n).
 * ------------------------------------------------------------------------------------------------------------------ */
static const uint32_t* g_rank;
static uint32_t g_h, g_n;

static int cmp_pair(const void* a, const void* b) {
    uint32_t i = *(const uint32_t*)a, j = *(const uint32_t*)b;
    if (g_rank[i] != g_rank[j]) return g_rank[i] < g_rank[j] ? -1 : 1;
    /* suffixes shorter than h sort first; cannot tie because the sentinel is unique */
    uint32_t ri = (i + g_h < g_n) ? g_rank[i + g_h] + 1 : 0;
    uint32_t rj = (j + g_h < g_n) ? g_rank[j + g_h] + 1 : 0;
    if (ri != rj) return ri < rj ? -1 : 1;
    return 0;
}

int tdcoracle_sa(const uint8_t* t, uint32_t n, uint32_t* sa) {
    if (n == 0) return 0;
    uint32_t* rank = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * n);
    if (!rank || !tmp) { free(rank); free(tmp); return -1; }
    for (uint32_t i = 0; i < n; i++) { sa[i] = i; rank[i] = t[i]; }
    g_rank = rank; g_n = n;
    for (g_h = 0;; g_h = g_h ? g_h * 2 : 1) {
        /* g_h == 0: sort by first byte only (rank[i+0] == rank[i], second key is redundant) */
        qsort(sa, n, sizeof(uint32_t), cmp_pair);
        tmp[sa[0]] = 0;
        uint32_t distinct = 1;
        for (uint32_t r = 1; r < n; r++) {
            if (cmp_pair(&sa[r - 1], &sa[r]) != 0) distinct++;
            tmp[sa[r]] = distinct - 1;
        }
        memcpy(rank, tmp, sizeof(uint32_t) * n);
        if (distinct == n) break;
        if (g_h >= n) break; /* not reachable with a unique sentinel */
    }
    free(rank); free(tmp);
    return 0;
}

/* ISA[SA[i]] = i — include/tudocomp/ds/ISAFromSA.hpp:37-39 */
void tdcoracle_isa(const uint32_t* sa, uint32_t n, uint32_t* isa) {
    for (uint32_t i = 0; i < n; i++) isa[sa[i]] = i;
}

/* Phi[SA[i]] = SA[i-1]; Phi[SA[0]] = SA[n-1] — include/tudocomp/ds/PhiFromSA.hpp:37-41 */
void tdcoracle_phi(const uint32_t* sa, uint32_t n, uint32_t* phi) {
    if (n == 0) return;
    for (uint32_t i = 1; i < n; i++) phi[sa[i]] = sa[i - 1];
    phi[sa[0]] = sa[n - 1];
}

/* PLCP in place over Phi, loop bound i < n-1 so PLCP[n-1] keeps Phi[n-1]; returns max over i <= n-2
 * — include/tudocomp/ds/PLCPFromPhi.hpp:36-44 */
uint32_t tdcoracle_plcp(const uint8_t* t, uint32_t n, const uint32_t* phi, uint32_t* plcp) {
    uint32_t mx = 0;
    if (n == 0) return 0;
    if (plcp != phi) memcpy(plcp, phi, sizeof(uint32_t) * n);
    uint32_t l = 0;
    for (uint32_t i = 0; i + 1 < n; i++) {
        uint32_t p = plcp[i];
        while (t[i + l] == t[p + l]) l++;
        if (l > mx) mx = l;
        plcp[i] = l;
        if (l) l--;
    }
    return mx;
}

/* LCP[0] = 0; LCP[i] = PLCP[SA[i]] — include/tudocomp/ds/LCPFromPLCP.hpp:43-47 */
void tdcoracle_lcp(const uint32_t* sa, const uint32_t* plcp, uint32_t n, uint32_t* lcp) {
    if (n == 0) return;
    lcp[0] = 0;
    for (uint32_t i = 1; i < n; i++) lcp[i] = plcp[sa[i]];
}

/* BWT[i] = SA[i]==0 ? T[n-1] : T[SA[i]-1] — include/tudocomp/ds/bwt.hpp:19-22 */
void tdcoracle_bwt(const uint8_t* t, const uint32_t* sa, uint32_t n, uint8_t* out) {
    for (uint32_t i = 0; i < n; i++) out[i] = sa[i] == 0 ? t[n - 1] : t[sa[i] - 1];
}

/* Greedy factorisation with naive PSV/NSV scans over SA/ISA/LCP
 * — include/tudocomp/compressors/LZSSLCPCompressor.hpp:60-115.
 * out receives (pos,src,len) u32 triples (lzss::Factor, lzss/LZSSFactors.hpp:13-20); returns the factor count,
 * or -1 if cap is too small.  threshold must be >= 1 (0 never terminates in the reference either). */
int64_t tdcoracle_lzss_lcp_factorize(const uint32_t* sa, const uint32_t* isa, const uint32_t* lcp, uint32_t n,
                                     uint32_t threshold, uint32_t* out, uint64_t cap) {
    uint64_t z = 0;
    for (uint64_t i = 0; i + 1 < n;) {
        const uint64_t cur = isa[i];
        /* PSV side: include current LCP, exclude the last (LZSSLCPCompressor.hpp:71-77) */
        uint64_t psv_lcp = lcp[cur];
        int64_t psv_pos = (int64_t)cur - 1;
        if (psv_lcp > 0) {
            while (psv_pos >= 0 && sa[psv_pos] > sa[cur]) {
                if (lcp[psv_pos] < psv_lcp) psv_lcp = lcp[psv_pos];
                psv_pos--;
            }
        }
        /* NSV side: exclude current, include the last (LZSSLCPCompressor.hpp:82-96) */
        uint64_t nsv_lcp = 0;
        uint64_t nsv_pos = cur + 1;
        if (nsv_pos < n) {
            nsv_lcp = UINT64_MAX;
            do {
                if (lcp[nsv_pos] < nsv_lcp) nsv_lcp = lcp[nsv_pos];
                if (sa[nsv_pos] < sa[cur]) break;
            } while (++nsv_pos < n);
            if (nsv_pos >= n) nsv_lcp = 0;
        }
        /* PSV wins ties (LZSSLCPCompressor.hpp:99-105) */
        const uint64_t mx = psv_lcp > nsv_lcp ? psv_lcp : nsv_lcp;
        if (mx >= threshold) {
            const uint64_t src_rank = (mx == psv_lcp) ? (uint64_t)psv_pos : nsv_pos;
            if (z >= cap) return -1;
            out[3 * z + 0] = (uint32_t)i;
            out[3 * z + 1] = sa[src_rank];
            out[3 * z + 2] = (uint32_t)mx;
            z++;
            i += mx;
        } else {
            i++;
        }
    }
    return (int64_t)z;
}

/* Header values lzss::encode_text derives from the factor list — lzss/LZSSCoding.hpp:24-39,
 * FactorBuffer bookkeeping lzss/LZSSFactors.hpp:33-47 (shortest starts at INDEX_MAX, longest at 0). */
void tdcoracle_factor_stats(const uint32_t* triples, uint64_t z, uint32_t n, uint32_t* flen_min, uint32_t* flen_max,
                            uint32_t* fdist_max) {
    uint32_t mn = UINT32_MAX, mx = 0;
    uint64_t p = 0, dist = 0;
    for (uint64_t k = 0; k < z; k++) {
        uint32_t pos = triples[3 * k], len = triples[3 * k + 2];
        if (len < mn) mn = len;
        if (len > mx) mx = len;
        if (pos - p > dist) dist = pos - p;
        p = (uint64_t)pos + len;
    }
    if (n - p > dist) dist = n - p;
    *flen_min = mn; *flen_max = mx; *fdist_max = (uint32_t)dist;
}

/* Decoder semantics of lzss::decode_text with DecodeBackBuffer — lzss/LZSSCoding.hpp:94-140,
 * lzss/LZSSDecodeBackBuffer.hpp:24-30: literals are copied, factors copy byte-by-byte from src (overlap allowed).
 * `literals` holds the text positions not covered by factors, in order.  Returns 0 if out == n bytes were produced. */
int tdcoracle_lzss_decode(const uint32_t* triples, uint64_t z, const uint8_t* text, uint32_t n, uint8_t* out) {
    uint64_t p = 0;
    for (uint64_t k = 0; k < z; k++) {
        uint32_t pos = triples[3 * k], src = triples[3 * k + 1], len = triples[3 * k + 2];
        if (pos < p || (uint64_t)pos + len > n || src >= pos) return -1;
        while (p < pos) { out[p] = text[p]; p++; }
        for (uint32_t j = 0; j < len; j++) out[p + j] = out[src + j];
        p += len;
    }
    while (p < n) { out[p] = text[p]; p++; }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * lzss::encode_text — include/tudocomp/compressors/lzss/LZSSCoding.hpp:18-92 — for coders that keep the Encoder
 * defaults for integers (binary in bits_for(max - min) bits, include/tudocomp/Coder.hpp:63-80) and write one fixed
 * code word per literal (BitCoder: 8 bits; HuffmanCoder: include/tudocomp/coders/HuffmanCoder.hpp:309-322, 562-568).
 * Bits go out MSB first (include/tudocomp/io/BitOStream.hpp:79-102).
 * ------------------------------------------------------------------------------------------------------------------ */
static uint32_t bits_for_u64(uint64_t v) { /* include/tudocomp/util.hpp:194: bits_for(0) == 1 */
    uint32_t b = 1;
    while (v >>= 1) b++;
    return b;
}

typedef struct {
    uint8_t* out;
    uint64_t cap, nbits;
    int overflow;
} bitsink;

static void sink_bit(bitsink* s, int bit) { /* BitOStream::write_bit, io/BitOStream.hpp:79-90 */
    uint64_t byte = s->nbits >> 3;
    if (byte >= s->cap) { s->overflow = 1; s->nbits++; return; }
    if ((s->nbits & 7) == 0) s->out[byte] = 0;
    if (bit) s->out[byte] |= (uint8_t)(1u << (7 - (s->nbits & 7)));
    s->nbits++;
}
static void sink_int(bitsink* s, uint64_t v, uint32_t bits) { /* BitOStream::write_int, io/BitOStream.hpp:98-102 */
    for (int i = (int)bits - 1; i >= 0; i--) sink_bit(s, i < 64 ? (int)((v >> i) & 1u) : 0);
}

/* Literals as lzss::TextLiterals yields them (lzss/LZSSLiterals.hpp:10-56): every position outside a factor. */
void tdcoracle_literal_histogram(const uint8_t* text, uint32_t n, const uint32_t* triples, uint64_t z, uint64_t hist[256]) {
    uint64_t p = 0;
    memset(hist, 0, 256 * sizeof(uint64_t));
    for (uint64_t k = 0; k < z; k++) {
        uint32_t pos = triples[3 * k], len = triples[3 * k + 2];
        while (p < pos) hist[text[p++]]++;
        p += len;
    }
    while (p < n) hist[text[p++]]++;
}

/* Width of the text-length field: LengthRange = TypeRange<len_t> (include/tudocomp/Range.hpp:95-99, :115), i.e. 32 bits
 * in the default build and 64 in a wide-index build (-DLEN_BITS=40 makes len_t = fast_t<uint_t<40>> = uint64_t,
 * include/tudocomp/def.hpp:100-114).  Checked against oracle/_ref/libtdcref40.so (the reference compiled with
 * -DLEN_BITS=40). */
static uint32_t g_len_field_bits = 32;
void tdcoracle_set_len_field_bits(uint32_t bits) { g_len_field_bits = bits == 64 ? 64 : 32; }

/* The stream continues a coder header of `lead_bits` (0..7) bits held in the high bits of `lead_byte`.
 * finalize: append BitOStream::~BitOStream's tail (io/BitOStream.hpp:53-64).  Returns bytes written, < 0 if cap is
 * too small; *nbits_out = stream bits incl. lead_bits (before the tail). */
int64_t tdcoracle_lzss_encode(const uint8_t* text, uint32_t n, const uint32_t* triples, uint64_t z, const uint64_t codes[256],
                              const uint8_t lens[256], uint32_t lead_bits, uint8_t lead_byte, int finalize, uint8_t* out,
                              uint64_t cap, uint64_t* nbits_out) {
    uint32_t flen_min, flen_max, fdist_max;
    tdcoracle_factor_stats(triples, z, n, &flen_min, &flen_max, &fdist_max);
    const uint32_t bn = bits_for_u64(n), bf = bits_for_u64(fdist_max);
    const uint32_t bl = bits_for_u64((uint64_t)flen_max - (uint64_t)flen_min); /* size_t arithmetic as in Range::delta */
    bitsink s = {out, cap, 0, 0};
    for (uint32_t i = 0; i < lead_bits; i++) sink_bit(&s, (lead_byte >> (7 - i)) & 1);
    sink_int(&s, n, g_len_field_bits); /* coder.encode(n, len_r)    :47 */
    sink_int(&s, flen_min, bn);  /* coder.encode(flen_min, text_r)  :48 (INDEX_MAX truncated when there is no factor) */
    sink_int(&s, flen_max, bn);  /* :49 */
    sink_int(&s, fdist_max, bn); /* :50 */
    uint64_t p = 0;
    for (uint64_t k = 0; k < z; k++) {
        uint32_t pos = triples[3 * k], src = triples[3 * k + 1], len = triples[3 * k + 2];
        if (pos == p) {
            sink_bit(&s, 0); /* :58-60 */
        } else {
            sink_bit(&s, 1); /* :62-68 */
            sink_int(&s, pos - p, bf);
        }
        while (p < pos) { uint8_t c = text[p++]; sink_int(&s, codes[c], lens[c]); } /* :71-73 */
        sink_int(&s, src, bn);                /* :77 */
        sink_int(&s, len - flen_min, bl);     /* :78 */
        p += len;
    }
    if (p < n) { /* :82-85 */
        sink_bit(&s, 1);
        sink_int(&s, n - p, bf);
    }
    while (p < n) { uint8_t c = text[p++]; sink_int(&s, codes[c], lens[c]); } /* :87-90 */
    if (nbits_out) *nbits_out = s.nbits;
    uint64_t bytes = (s.nbits + 7) / 8;
    if (finalize) {
        const uint32_t used = (uint32_t)(s.nbits & 7);
        const uint64_t whole = s.nbits >> 3;
        if (used <= 5) {
            if (whole >= cap) return -1;
            if (used == 0) out[whole] = 0;
            out[whole] |= (uint8_t)used;
            bytes = whole + 1;
        } else {
            if (whole + 1 >= cap) return -1;
            out[whole + 1] = (uint8_t)used;
            bytes = whole + 2;
        }
    }
    return s.overflow ? -1 : (int64_t)bytes;
}

/* ------------------------------------------------------------------------------------------------------------------
 * The stream stages behind the BWT in `bwt:mtf:rle:encode(huff)`.
 * ------------------------------------------------------------------------------------------------------------------ */
/* mtf_encode — include/tudocomp/compressors/MTFCompressor.hpp:17-30, 46-56: table starts as 0..255; each byte is looked
 * up linearly, its index is emitted and it moves to the front. */
void tdcoracle_mtf_encode(const uint8_t* in, uint64_t n, uint8_t* out) {
    uint8_t table[256];
    for (int i = 0; i < 256; i++) table[i] = (uint8_t)i;
    for (uint64_t p = 0; p < n; p++) {
        const uint8_t v = in[p];
        int i = 0;
        while (table[i] != v) i++;
        for (int j = i; j > 0; j--) table[j] = table[j - 1];
        table[0] = v;
        out[p] = (uint8_t)i;
    }
}

/* rle_encode — include/tudocomp/compressors/RunLengthEncoder.hpp:15-31 with write_vbyte
 * (include/tudocomp/util/vbyte.hpp:27-37): the first byte of a run is copied; a second equal byte is copied and followed
 * by vbyte(number of further equal bytes + offset).  Returns the output length, or -1 if cap is too small.
 * Quirk reproduced on purpose: the compressor instantiates char_type = char (std::istream), and the run counter compares
 * `is.peek() == c` (:24), i.e. an int in 0..255 with a signed char: for bytes >= 0x80 it is never true, so such a run is
 * never merged — every further byte of it is written as  c vbyte(0 + offset).  (At end of input peek() is EOF = -1, which
 * does equal the char 0xFF: the reference then spins forever on inputs that END in 0xFF 0xFF; here such a tail is encoded
 * like any other run of a byte >= 0x80.) */
int64_t tdcoracle_rle_encode(const uint8_t* in, uint64_t n, uint64_t offset, uint8_t* out, uint64_t cap) {
    uint64_t o = 0, i = 0;
    if (n == 0) return 0;
#define PUT(b) do { if (o >= cap) return -1; out[o++] = (uint8_t)(b); } while (0)
    uint8_t prev = in[i++];
    PUT(prev);
    while (i < n) {
        const uint8_t c = in[i++];
        if (prev == c) {
            uint64_t run = 0;
            while (c < 0x80 && i < n && in[i] == c) { run++; i++; }
            PUT(c);
            uint64_t v = run + offset;
            do {
                uint8_t byte = (uint8_t)(v & 0x7f);
                v >>= 7;
                if (v > 0) byte |= 0x80;
                PUT(byte);
            } while (v > 0);
        } else {
            PUT(c);
        }
        prev = c;
    }
#undef PUT
    return (int64_t)o;
}

/* LiteralEncoder::compress — include/tudocomp/compressors/LiteralEncoder.hpp:23-32: after the coder's own header every
 * byte is written with its code word; then BitOStream's tail.  Same conventions as tdcoracle_lzss_encode. */
int64_t tdcoracle_literal_encode(const uint8_t* in, uint64_t n, const uint64_t codes[256], const uint8_t lens[256],
                                 uint32_t lead_bits, uint8_t lead_byte, int finalize, uint8_t* out, uint64_t cap,
                                 uint64_t* nbits_out) {
    bitsink s = {out, cap, 0, 0};
    for (uint32_t i = 0; i < lead_bits; i++) sink_bit(&s, (lead_byte >> (7 - i)) & 1);
    for (uint64_t p = 0; p < n; p++) sink_int(&s, codes[in[p]], lens[in[p]]);
    if (nbits_out) *nbits_out = s.nbits;
    uint64_t bytes = (s.nbits + 7) / 8;
    if (finalize) {
        const uint32_t used = (uint32_t)(s.nbits & 7);
        const uint64_t whole = s.nbits >> 3;
        if (used <= 5) {
            if (whole >= cap) return -1;
            if (used == 0) out[whole] = 0;
            out[whole] |= (uint8_t)used;
            bytes = whole + 1;
        } else {
            if (whole + 1 >= cap) return -1;
            out[whole + 1] = (uint8_t)used;
            bytes = whole + 2;
        }
    }
    return s.overflow ? -1 : (int64_t)bytes;
}

/* One call for the whole TextDS as TextDS::require orders it (include/tudocomp/ds/TextDS.hpp:247-292).
 * Any output may be NULL; scratch is allocated as needed. */
int tdcoracle_textds(const uint8_t* t, uint32_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp, uint32_t* phi,
                     uint32_t* plcp, uint32_t* max_lcp) {
    int own_sa = 0, own_phi = 0, own_plcp = 0;
    if (n == 0) return 0;
    if (t[n - 1] != 0) return -2; /* "Input has no sentinel!" (TextDS.hpp:132-138) */
    if (!sa) { sa = (uint32_t*)malloc(sizeof(uint32_t) * n); own_sa = 1; }
    if (!phi) { phi = (uint32_t*)malloc(sizeof(uint32_t) * n); own_phi = 1; }
    if (!plcp) { plcp = (uint32_t*)malloc(sizeof(uint32_t) * n); own_plcp = 1; }
    if (!sa || !phi || !plcp) return -1;
    int rc = tdcoracle_sa(t, n, sa);
    if (rc == 0) {
        tdcoracle_phi(sa, n, phi);
        uint32_t mx = tdcoracle_plcp(t, n, phi, plcp);
        if (max_lcp) *max_lcp = mx;
        if (lcp) tdcoracle_lcp(sa, plcp, n, lcp);
        if (isa) tdcoracle_isa(sa, n, isa);
    }
    if (own_sa) free(sa);
    if (own_phi) free(phi);
    if (own_plcp) free(plcp);
    return rc;
}
