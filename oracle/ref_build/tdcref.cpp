// TEST INFRASTRUCTURE ONLY — never linked into or loaded by the product (tudocomp_b200 / libtdcgpu.so).
//
// Thin C wrapper around the UNMODIFIED reference (headers compiled where they lie under /root/reference/include,
// plus the two offline shims in shim/).  It is the "oracle/_ref" of the task contract: the reference's own CPU
// implementation of the hot path, callable from the Python tests / bench over ctypes.
//
// What is called (nothing is re-implemented here):
//   * TextDS<> with its default providers  SADivSufSort / PhiFromSA / PLCPFromPhi / LCPFromPLCP / ISAFromSA
//     (include/tudocomp/ds/TextDS.hpp:247-292 and the provider files next to it)
//   * LZSSLCPCompressor<coder>::compress   (include/tudocomp/compressors/LZSSLCPCompressor.hpp:42-124)
//   * bwt::bwt                             (include/tudocomp/ds/bwt.hpp:19-22)
//   * lzss::decode_text via ::decompress   (include/tudocomp/compressors/lzss/LZSSCoding.hpp:94-140)
//
// The factor list is private to LZSSLCPCompressor::compress, so it is observed through a recording coder
// ("spy") plugged in as the compressor's coder_t: lzss::encode_text (LZSSCoding.hpp:18-92) hands every factor
// to coder.encode(src, Range) / coder.encode(len, MinDistributedRange) in order, and we log those calls.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <tudocomp/CreateAlgorithm.hpp>
#include <tudocomp/coders/ASCIICoder.hpp>
#include <tudocomp/coders/BitCoder.hpp>
#include <tudocomp/coders/HuffmanCoder.hpp>
#include <tudocomp/compressors/BWTCompressor.hpp>
#include <tudocomp/compressors/LZSSLCPCompressor.hpp>
#include <tudocomp/compressors/LiteralEncoder.hpp>
#include <tudocomp/compressors/MTFCompressor.hpp>
#include <tudocomp/compressors/RunLengthEncoder.hpp>
#include <tudocomp/ds/TextDS.hpp>
#include <tudocomp/ds/bwt.hpp>
#include <tudocomp/io.hpp>
#include <tudocomp_stat/StatPhase.hpp>

using namespace tdc;

namespace {

struct SpyLog {
    std::vector<uint32_t> triples;  // pos, src, len
    uint64_t pos = 0;               // running text cursor
    uint64_t pending_src = 0;
    int header_left = 3;            // flen_min, flen_max, fdist_max come first as plain Range values
    bool expect_dist = false;
    uint64_t n = 0, flen_min = 0, flen_max = 0, fdist_max = 0;
    int header_idx = 0;
};
SpyLog* g_spy = nullptr;

// A coder that writes nothing and records the call sequence of lzss::encode_text.
class SpyCoder : public Algorithm {
public:
    inline static Meta meta() {
        Meta m("coder", "spy", "records encode() calls");
        return m;
    }
    SpyCoder() = delete;

    class Encoder : public tdc::Encoder {
    public:
        using tdc::Encoder::Encoder;

        template <typename value_t>
        inline void encode(value_t v, const LengthRange&) {  // n  (LZSSCoding.hpp:47)
            g_spy->n = uint64_t(v);
        }
        template <typename value_t>
        inline void encode(value_t v, const BitRange&) {  // literal-run flag (LZSSCoding.hpp:58-62, 83)
            g_spy->expect_dist = bool(v);
        }
        template <typename value_t>
        inline void encode(value_t, const LiteralRange&) {  // one literal (LZSSCoding.hpp:72, 89)
            g_spy->pos += 1;
        }
        template <typename value_t>
        inline void encode(value_t v, const MinDistributedRange&) {  // factor length (LZSSCoding.hpp:78)
            g_spy->triples.push_back(uint32_t(g_spy->pos));
            g_spy->triples.push_back(uint32_t(g_spy->pending_src));
            g_spy->triples.push_back(uint32_t(v));
            g_spy->pos += uint64_t(v);
        }
        template <typename value_t>
        inline void encode(value_t v, const Range&) {
            if (g_spy->header_left > 0) {  // LZSSCoding.hpp:48-50
                if (g_spy->header_left == 3) g_spy->flen_min = uint64_t(v);
                if (g_spy->header_left == 2) g_spy->flen_max = uint64_t(v);
                if (g_spy->header_left == 1) g_spy->fdist_max = uint64_t(v);
                g_spy->header_left--;
            } else if (g_spy->expect_dist) {  // literal count (LZSSCoding.hpp:67, 84)
                g_spy->expect_dist = false;
            } else {  // factor source (LZSSCoding.hpp:77)
                g_spy->pending_src = uint64_t(v);
            }
        }
    };

    class Decoder : public tdc::Decoder {
    public:
        using tdc::Decoder::Decoder;
    };
};

std::string g_err;
std::string g_stats;

template <class F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

template <class C>
std::vector<uint8_t> run_compress(View text, const std::string& opts, double* secs) {
    std::vector<uint8_t> out;
    {
        Input in(text);
        Output o = Output::from_memory(out);
        auto c = create_algo<C>(opts);
        StatPhase root("root");
        auto t0 = std::chrono::steady_clock::now();
        c.compress(in, o);
        auto t1 = std::chrono::steady_clock::now();
        if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
        g_stats = root.to_json().str();
    }
    return out;
}

}  // namespace

extern "C" {

const char* tdcref_last_error() { return g_err.c_str(); }
const char* tdcref_last_stats_json() { return g_stats.c_str(); }

// text must already be the escaped text with its single trailing 0 (what the path receives).
// Any output pointer may be NULL.  PLCP[n-1] keeps the reference's stale value (Appendix A.3).
int tdcref_textds(const uint8_t* text, uint64_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp, uint32_t* phi,
                  uint32_t* plcp, uint32_t* max_lcp) {
    return guarded([&] {
        View v(text, n);
        ds::dsflags_t flags = 0;
        if (sa) flags |= ds::SA;
        if (isa) flags |= ds::ISA;
        if (lcp) flags |= ds::LCP;
        if (phi) flags |= ds::PHI;
        if (plcp) flags |= ds::PLCP;
        if (max_lcp) flags |= ds::PLCP;
        auto t = create_algo<TextDS<>>("compress=\"plain\"", v, flags);
        if (sa) { auto& a = t.require_sa(); for (uint64_t i = 0; i < n; i++) sa[i] = a[i]; }
        if (isa) { auto& a = t.require_isa(); for (uint64_t i = 0; i < n; i++) isa[i] = a[i]; }
        if (lcp) { auto& a = t.require_lcp(); for (uint64_t i = 0; i < n; i++) lcp[i] = a[i]; }
        if (phi) { auto& a = t.require_phi(); for (uint64_t i = 0; i < n; i++) phi[i] = a[i]; }
        if (plcp || max_lcp) {
            auto& a = t.require_plcp();
            if (plcp) for (uint64_t i = 0; i < n; i++) plcp[i] = a[i];
            if (max_lcp) *max_lcp = a.max_lcp();
        }
    });
}

// SA, ISA, LCP and the BWT from ONE TextDS (tests/golden/make_size_hashes.py: at 2^30 B every extra SA costs ~15 min).
int tdcref_index_bwt(const uint8_t* text, uint64_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp, uint8_t* bwt_out,
                     uint32_t* max_lcp) {
    return guarded([&] {
        View v(text, n);
        auto t = create_algo<TextDS<>>("compress=\"plain\"", v, ds::SA | ds::ISA | ds::LCP);
        auto& s = t.require_sa();
        for (uint64_t i = 0; i < n; i++) { sa[i] = s[i]; bwt_out[i] = bwt::bwt(t, s, i); }
        auto& a = t.require_isa();
        for (uint64_t i = 0; i < n; i++) isa[i] = a[i];
        auto& l = t.require_lcp();
        for (uint64_t i = 0; i < n; i++) lcp[i] = l[i];
        if (max_lcp) *max_lcp = l.max_lcp();
    });
}

// BWT through bwt::bwt over the reference SA.
int tdcref_bwt(const uint8_t* text, uint64_t n, uint8_t* out) {
    return guarded([&] {
        View v(text, n);
        auto t = create_algo<TextDS<>>("", v, ds::SA);
        auto& sa = t.require_sa();
        for (uint64_t i = 0; i < n; i++) out[i] = bwt::bwt(t, sa, i);
    });
}

// Factor list of lzss_lcp(threshold) as (pos,src,len) u32 triples.  Returns the count, or <0 on error.
// Also reports the header values encode_text derives (flen_min/flen_max/fdist_max).
int64_t tdcref_lzss_lcp_factors(const uint8_t* text, uint64_t n, uint32_t threshold, uint32_t* triples, uint64_t cap,
                                uint64_t* header3) {
    SpyLog log;
    g_spy = &log;
    int rc = guarded([&] {
        View v(text, n);
        run_compress<LZSSLCPCompressor<SpyCoder>>(v, "threshold=" + std::to_string(threshold), nullptr);
    });
    g_spy = nullptr;
    if (rc) return rc;
    uint64_t z = log.triples.size() / 3;
    if (header3) { header3[0] = log.flen_min; header3[1] = log.flen_max; header3[2] = log.fdist_max; }
    if (triples) {
        if (z > cap) { g_err = "factor buffer too small"; return -2; }
        std::memcpy(triples, log.triples.data(), log.triples.size() * 4);
    }
    return int64_t(z);
}

// Raw archive (no driver header) of lzss_lcp(coder, threshold).  coder: 0=bit 1=huff 2=ascii.
// Returns archive length (even if > cap, in which case nothing is copied), or <0 on error.  *secs = wall time of compress().
int64_t tdcref_lzss_lcp_compress(const uint8_t* text, uint64_t n, uint32_t threshold, int coder, uint8_t* out,
                                 uint64_t cap, double* secs) {
    std::vector<uint8_t> res;
    int rc = guarded([&] {
        View v(text, n);
        std::string o = "threshold=" + std::to_string(threshold);
        if (coder == 0) res = run_compress<LZSSLCPCompressor<BitCoder>>(v, o, secs);
        else if (coder == 1) res = run_compress<LZSSLCPCompressor<HuffmanCoder>>(v, o, secs);
        else if (coder == 2) res = run_compress<LZSSLCPCompressor<ASCIICoder>>(v, o, secs);
        else throw std::runtime_error("unknown coder id");
    });
    if (rc) return rc;
    if (out && res.size() <= cap) std::memcpy(out, res.data(), res.size());
    return int64_t(res.size());
}

// What HuffmanCoder::Encoder's constructor does with the literal counts (coders/HuffmanCoder.hpp:526-548), driven by a
// histogram instead of a literal iterator: header bits as written by the reference's own BitOStream /
// huff::huffmantable_encode, and the code word of every literal as huff::huffman_encode would write it (:309-322).
// header receives the whole header bytes plus the partial last byte; *header_bits = exact bit count.
// coder: 0=bit (no header, literals in 8 bits, Coder.hpp:63-66 via LiteralRange), 1=huff.
int tdcref_literal_coder(int coder, const uint64_t hist[256], uint8_t* header, uint64_t cap, uint64_t* header_bits,
                         uint64_t codes[256], uint8_t lens[256]) {
    return guarded([&] {
        std::vector<uint8_t> head;
        for (int c = 0; c < 256; c++) { codes[c] = uint64_t(c); lens[c] = 8; }
        {
            Output ho = Output::from_memory(head);
            BitOStream bos(ho);
            if (coder == 1) {
                len_compact_t* C = new len_compact_t[ULITERAL_MAX + 1];
                for (int c = 0; c < 256; c++) C[c] = len_compact_t(hist[c]);
                const len_t alphabet_size = huff::effective_alphabet_size(C);
                if (alphabet_size <= 1) {
                    delete[] C;
                    bos.write_bit(0);
                } else {
                    huff::extended_huffmantable table = huff::gen_huffmantable(C);  // deletes C
                    bos.write_bit(1);
                    huff::huffmantable_encode(bos, table);
                    for (int c = 0; c < 256; c++) { codes[c] = 0; lens[c] = 0; }
                    for (size_t i = 0; i < table.alphabet_size; i++) {
                        codes[table.ordered_map_from_effective[i]] = table.codewords[i];
                        lens[table.ordered_map_from_effective[i]] = table.ordered_codelengths[i];
                    }
                }
            } else if (coder != 0) {
                throw std::runtime_error("unknown coder id");
            }
        }  // ~BitOStream appends its tail (io/BitOStream.hpp:53-64); undo it to get the exact bit count
        const uint8_t last = head.back();
        const unsigned used = last & 7u;
        uint64_t bits;
        if (used <= 5) {
            bits = 8 * uint64_t(head.size() - 1) + used;
            head.back() = uint8_t(last & ~7u);
        } else {
            bits = 8 * uint64_t(head.size() - 2) + used;
            head.pop_back();
        }
        if (bits % 8 == 0 && !head.empty() && head.size() * 8 > bits) head.pop_back();
        if (head.size() > cap) throw std::runtime_error("header buffer too small");
        if (!head.empty()) std::memcpy(header, head.data(), head.size());
        *header_bits = bits;
    });
}

// One stream stage of `bwt:mtf:rle:encode(huff)` through the reference's own Compressor classes (no sentinel / escaping:
// these stages declare no input restrictions).  stage: 0 = MTFCompressor, 1 = RunLengthEncoder(offset),
// 2 = LiteralEncoder<BitCoder>, 3 = LiteralEncoder<HuffmanCoder>.  Returns the output length.
int64_t tdcref_stream_stage(int stage, const uint8_t* in, uint64_t n, uint64_t offset, uint8_t* out, uint64_t cap, double* secs) {
    std::vector<uint8_t> res;
    int rc = guarded([&] {
        View v(in, n);
        if (stage == 0) res = run_compress<MTFCompressor>(v, "", secs);
        else if (stage == 1) res = run_compress<RunLengthEncoder>(v, "offset=" + std::to_string(offset), secs);
        else if (stage == 2) res = run_compress<LiteralEncoder<BitCoder>>(v, "", secs);
        else if (stage == 3) res = run_compress<LiteralEncoder<HuffmanCoder>>(v, "", secs);
        else throw std::runtime_error("unknown stage id");
    });
    if (rc) return rc;
    if (out && res.size() <= cap) std::memcpy(out, res.data(), res.size());
    return int64_t(res.size());
}

// Decompress a raw lzss_lcp archive with the reference decoder (round-trip checker).  Output = text incl. trailing 0.
int64_t tdcref_lzss_lcp_decompress(const uint8_t* arc, uint64_t len, int coder, uint8_t* out, uint64_t cap) {
    std::vector<uint8_t> res;
    int rc = guarded([&] {
        Input in(View(arc, len));
        Output o = Output::from_memory(res);
        if (coder == 0) create_algo<LZSSLCPCompressor<BitCoder>>("").decompress(in, o);
        else if (coder == 1) create_algo<LZSSLCPCompressor<HuffmanCoder>>("").decompress(in, o);
        else if (coder == 2) create_algo<LZSSLCPCompressor<ASCIICoder>>("").decompress(in, o);
        else throw std::runtime_error("unknown coder id");
    });
    if (rc) return rc;
    if (out && res.size() <= cap) std::memcpy(out, res.data(), res.size());
    return int64_t(res.size());
}

// BWTCompressor::compress output (n bytes) with wall time: CPU baseline for the bwt stage.
int64_t tdcref_bwt_compress(const uint8_t* text, uint64_t n, uint8_t* out, uint64_t cap, double* secs) {
    std::vector<uint8_t> res;
    int rc = guarded([&] {
        View v(text, n);
        res = run_compress<BWTCompressor<>>(v, "", secs);
    });
    if (rc) return rc;
    if (out && res.size() <= cap) std::memcpy(out, res.data(), res.size());
    return int64_t(res.size());
}

// Escaping + sentinel exactly as the driver applies it for uses_textds compressors
// (src/tudocomp_driver/tudocomp_driver.cpp:268-270; io/EscapeMap.hpp:39-64).  Returns the escaped length incl. the 0.
int64_t tdcref_escape(const uint8_t* raw, uint64_t len, uint8_t* out, uint64_t cap) {
    int64_t r = -1;
    int rc = guarded([&] {
        Input in(View(raw, len));
        Input restricted(in, io::InputRestrictions({0}, true));
        auto v = restricted.as_view();
        r = int64_t(v.size());
        if (out && v.size() <= cap) std::memcpy(out, v.data(), v.size());
    });
    return rc ? rc : r;
}

}  // extern "C"
