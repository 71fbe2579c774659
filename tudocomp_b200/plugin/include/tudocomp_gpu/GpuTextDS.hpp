// GPU-backed drop-ins for tudocomp's text index and lzss_lcp / bwt compressors.
//
// This header is compiled AGAINST the reference's headers (it is the host side "in the reference's own language") and
// talks to the device only through the C ABI in include/tdcgpu.h.  It adds, without touching any reference file:
//
//   GpuTextDS                                  a whole replacement `text_t` (legal template argument of
//                                              LZSSLCPCompressor<coder_t, text_t> / BWTCompressor<text_t>,
//                                              compressors/LZSSLCPCompressor.hpp:23, BWTCompressor.hpp:13) with the
//                                              TextDS surface the consumers use: require_sa/isa/lcp/phi/plcp, require(),
//                                              size(), operator[], text(), SA/ISA/LCP/PHI/PLCP flags,
//                                              common_restrictions(), meta()            (ds/TextDS.hpp:23-346)
//   LZSSLCPCompressor<coder_t, GpuTextDS>      partial specialisation whose "Factorize" phase is the device factoriser;
//                                              same Meta("compressor","lzss_lcp"), same options (coder, textds, threshold),
//                                              same phases, the reference's own encode_text / decode_text
//   BWTCompressor<GpuTextDS>                   partial specialisation writing the device-side BWT gather
//
// Arrays handed back to callers are tdc::DynamicIntVector at width 32 — byte-identical to uint32_t[n]
// (ds/BitPackingVector.hpp:259-271) — and honour the `compress` option like the reference providers do
// (ds/SADivSufSort.hpp:53-63 etc.): "delayed"/"compressed" narrow to bits_for(n) / bits_for(max_lcp) after filling.
//
// Errors: C-ABI failures become std::runtime_error (caught by the driver's main like any Env::error,
// src/tudocomp_driver/tudocomp_driver.cpp:392-395); a text without sentinel throws the same std::logic_error as
// TextDS (ds/TextDS.hpp:132-138).  Single-threaded like the reference; StatPhase is only touched from the caller.
#pragma once

#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <tudocomp/Algorithm.hpp>
#include <tudocomp/coders/BitCoder.hpp>
#include <tudocomp/coders/HuffmanCoder.hpp>
#include <tudocomp/compressors/BWTCompressor.hpp>
#include <tudocomp/compressors/LZSSLCPCompressor.hpp>
#include <tudocomp/ds/ArrayDS.hpp>
#include <tudocomp/ds/CompressMode.hpp>
#include <tudocomp/ds/IntVector.hpp>
#include <tudocomp/ds/TextDS.hpp>
#include <tudocomp/ds/TextDSFlags.hpp>
#include <tudocomp_stat/StatPhase.hpp>

#include <tdcgpu.h>

namespace tdc {

namespace gpu_detail {
inline void check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string("tdcgpu: ") + what + ": " + tdcgpu_last_error());
}
// Device of the calling THREAD: TDCGPU_DEVICE (default 0) unless the host program dealt this thread a device of its own
// (tdc_block's worker threads: one thread per GPU in one process, so that CUDA is initialised once and not once per worker).
inline int& thread_device_override() {
    static thread_local int d = -1;
    return d;
}
inline int device_from_env() {
    if (thread_device_override() >= 0) return thread_device_override();
    const char* e = std::getenv("TDCGPU_DEVICE");
    return e ? std::atoi(e) : 0;
}
// One device context is kept alive between compress() calls of a thread instead of creating and destroying one per text
// (TDCGPU_CTX_CACHE=0 switches this off): the arrays, the scratch arena (~57 n bytes) and the pinned staging buffers are
// then allocated once for the largest text seen — per-call cudaMalloc/cudaFree of that much memory costs more than the
// kernels for mid-size texts.  tdc itself is single-threaded; the cache is per thread, so a multi-threaded host
// gets one cached context per worker thread.  The cached context is deliberately not destroyed at exit (static destructors
// may run after the CUDA runtime has shut down).
inline bool ctx_cache_enabled() {
    const char* e = std::getenv("TDCGPU_CTX_CACHE");
    return !(e && *e == '0');
}
inline tdcgpu_ctx*& cached_ctx() {
    static thread_local tdcgpu_ctx* c = nullptr;
    return c;
}
inline tdcgpu_ctx* acquire_ctx() {
    if (ctx_cache_enabled() && cached_ctx()) {
        tdcgpu_ctx* c = cached_ctx();
        cached_ctx() = nullptr;
        return c;
    }
    // The first context of a single-device process: unless the user restricted the visible devices, restrict them to the one
    // device this process uses — the CUDA runtime then initialises one GPU instead of every GPU of the box (seconds of
    // start-up on an 8-GPU node; the variable is read at the first CUDA call, which is the tdcgpu_create below).
    static bool first = true;
    int device = device_from_env();
    if (first && thread_device_override() < 0) {
        first = false;
        if (!std::getenv("CUDA_VISIBLE_DEVICES") && !std::getenv("TDCGPU_KEEP_ALL_DEVICES_VISIBLE")) {
            setenv("CUDA_VISIBLE_DEVICES", std::to_string(device).c_str(), 1);
            setenv("TDCGPU_DEVICE", "0", 1);
            device = 0;
        }
    }
    tdcgpu_ctx* raw = nullptr;
    check(tdcgpu_create(device, &raw), "create");
    return raw;
}
inline void release_ctx(tdcgpu_ctx* c) {
    if (!c) return;
    if (ctx_cache_enabled() && !cached_ctx()) cached_ctx() = c;
    else tdcgpu_destroy(c);
}
// log the device-side phase times under the current StatPhase (tudocomp_stat/StatPhase.hpp:217-220)
inline void log_phases(tdcgpu_ctx* ctx) {
    for (int i = 0; i < tdcgpu_phase_count(ctx); i++) {
        StatPhase::log((std::string("gpu_ms:") + tdcgpu_phase_name(ctx, i)).c_str(), tdcgpu_phase_ms(ctx, i));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Device-side lzss::encode_text (compressors/lzss/LZSSCoding.hpp:18-92): which coders can hand their literal code words
// to the device, and how their own stream header is produced.  Everything coder-specific below is the reference's own
// code (huff::gen_huffmantable, huff::huffmantable_encode, BitOStream); only the literal COUNTS come from the device.
// ---------------------------------------------------------------------------------------------------------------------
struct LiteralCodeTable {
    uint64_t codes[256];
    uint8_t lens[256];
};

template <typename coder_t>
struct DeviceLiteralCoder {
    static constexpr bool supported = false;  // e.g. ASCIICoder (variable-length decimal integers): host encode_text
    static void header(io::BitOStream&, const uint64_t*, LiteralCodeTable&) {}
};

/// BitCoder: no header; literals use the Encoder default for LiteralRange = 8 bits (Coder.hpp:63-66, Range.hpp:89-93).
template <>
struct DeviceLiteralCoder<BitCoder> {
    static constexpr bool supported = true;
    static void header(io::BitOStream&, const uint64_t*, LiteralCodeTable& t) {
        for (int c = 0; c < 256; c++) { t.codes[c] = uint64_t(c); t.lens[c] = 8; }
    }
};

/// HuffmanCoder: mirrors HuffmanCoder::Encoder's constructor (coders/HuffmanCoder.hpp:526-548) with the literal counts
/// taken from the device instead of huff::count_alphabet_literals (:37-49); code words as huff::huffman_encode writes
/// them (:309-322), raw 8 bits when the effective alphabet has a single symbol (:562-568).
template <>
struct DeviceLiteralCoder<HuffmanCoder> {
    static constexpr bool supported = true;
    static void header(io::BitOStream& out, const uint64_t* hist, LiteralCodeTable& t) {
        for (int c = 0; c < 256; c++) { t.codes[c] = uint64_t(c); t.lens[c] = 8; }
        len_compact_t* C = new len_compact_t[ULITERAL_MAX + 1];
        for (size_t c = 0; c <= ULITERAL_MAX; c++) C[c] = len_compact_t(hist[c]);
        if (huff::effective_alphabet_size(C) <= 1) {
            delete[] C;
            out.write_bit(0);
            return;
        }
        huff::extended_huffmantable table = huff::gen_huffmantable(C);  // takes ownership of C
        out.write_bit(1);
        huff::huffmantable_encode(out, table);
        for (int c = 0; c < 256; c++) { t.codes[c] = 0; t.lens[c] = 0; }
        for (size_t i = 0; i < table.alphabet_size; i++) {
            t.codes[table.ordered_map_from_effective[i]] = table.codewords[i];
            t.lens[table.ordered_map_from_effective[i]] = table.ordered_codelengths[i];
        }
    }
};

/// Strip what BitOStream::~BitOStream appended (io/BitOStream.hpp:53-64) from a finished scratch stream:
/// returns the exact number of payload bits; `bytes` keeps the whole bytes plus the partial last byte.
inline uint64_t strip_bitstream_tail(std::vector<uint8_t>& bytes) {
    const uint8_t last = bytes.back();
    const unsigned used = last & 7u;
    uint64_t bits;
    if (used <= 5) {
        bits = 8 * uint64_t(bytes.size() - 1) + used;
        bytes.back() = uint8_t(last & ~7u);
        if (used == 0) bytes.pop_back();
    } else {
        bits = 8 * uint64_t(bytes.size() - 2) + used;
        bytes.pop_back();
    }
    return bits;
}

// ---------------------------------------------------------------------------------------------------------------------
// Device-resident chains.  tudocomp_driver/ChainCompressor.hpp:54-62 hands a host vector from stage to stage, so every GPU
// stage of `bwt:mtf:rle:encode(huff)` would cross PCIe twice.  A GPU compressor additionally implements DeviceStage; the
// GPU-aware chain (tudocomp_gpu/GpuChainCompressor.hpp) then keeps the bytes between two such stages in HBM.
// ---------------------------------------------------------------------------------------------------------------------
struct DeviceBytes {  // a byte stream in device memory, owned by this object
    uint8_t* p = nullptr;
    uint64_t n = 0, cap = 0;
    DeviceBytes() = default;
    DeviceBytes(const DeviceBytes&) = delete;
    DeviceBytes& operator=(const DeviceBytes&) = delete;
    ~DeviceBytes() { reset(); }
    void reset() {
        if (p) tdcgpu_device_free(nullptr, p);
        p = nullptr;
        n = cap = 0;
    }
    void alloc(tdcgpu_ctx* ctx, uint64_t bytes) {
        reset();
        p = static_cast<uint8_t*>(tdcgpu_device_alloc(ctx, bytes));
        if (!p) throw std::runtime_error(std::string("tdcgpu: device_alloc: ") + tdcgpu_last_error());
        cap = bytes;
    }
};

struct StageInput {  // exactly one of the two is set
    Input* host = nullptr;
    const DeviceBytes* dev = nullptr;
};

struct StreamCtx {  // one context (stream + scratch) for the duration of a call
    tdcgpu_ctx* ctx = nullptr;
    StreamCtx() : ctx(acquire_ctx()) {}
    ~StreamCtx() { release_ctx(ctx); }
    StreamCtx(const StreamCtx&) = delete;
    StreamCtx& operator=(const StreamCtx&) = delete;
};
inline void download(const DeviceBytes& d, std::vector<uint8_t>& out) {
    out.resize(d.n);
    if (!d.n) return;
    StreamCtx g;
    check(tdcgpu_device_copy(g.ctx, out.data(), d.p, d.n, 1), "device_copy");
}
inline void upload(DeviceBytes& d, const std::vector<uint8_t>& in) {
    StreamCtx g;
    d.alloc(g.ctx, in.size());
    d.n = in.size();
    if (!in.empty()) check(tdcgpu_device_copy(g.ctx, d.p, in.data(), in.size(), 0), "device_copy");
}

class DeviceStage {
public:
    virtual ~DeviceStage() = default;
    /// The stage's compress() with its input in `in` and its output in `dev_out` (device-resident, for the next GPU stage)
    /// if that is non-null, else in `host_out`.  Same bytes as compress(Input&, Output&).
    virtual void compress_stage(const StageInput& in, DeviceBytes* dev_out, Output* host_out) = 0;
};

/// Page-locked staging buffer of the C ABI (falls back to pageable memory, which the ABI stages itself).
struct PinnedBuffer {
    uint8_t* data;
    size_t size;
    bool pinned;
    explicit PinnedBuffer(size_t n) : data(static_cast<uint8_t*>(tdcgpu_pinned_alloc(n))), size(n), pinned(data != nullptr) {
        if (!data) data = new uint8_t[n];
    }
    ~PinnedBuffer() { if (pinned) tdcgpu_pinned_free(data); else delete[] data; }
    PinnedBuffer(const PinnedBuffer&) = delete;
    PinnedBuffer& operator=(const PinnedBuffer&) = delete;
};

// the 16 MiB drain buffer of the archive streams, allocated once per process (tdc is single-threaded; like the cached
// context it is deliberately not released at exit): cudaMallocHost + cudaFreeHost per compress() cost a few ms each
inline PinnedBuffer& drain_buffer() {
    static thread_local PinnedBuffer* b = new PinnedBuffer(size_t(16) << 20);
    return *b;
}

inline bool host_encode_forced() {
    const char* e = std::getenv("TDCGPU_HOST_ENCODE");  // A/B switch: run the reference's encode_text on the host
    return e && *e && *e != '0';
}
}  // namespace gpu_detail

/// One array of the GPU text index, fetched on demand into a DynamicIntVector (what ArrayDS providers expose).
class GpuArray : public DynamicIntVector {
    len_t m_max_lcp = 0;

public:
    using data_type = DynamicIntVector;
    inline GpuArray() {}
    inline GpuArray(DynamicIntVector&& iv, len_t max_lcp) : DynamicIntVector(std::move(iv)), m_max_lcp(max_lcp) {}
    inline len_t max_lcp() const { return m_max_lcp; }
    inline DynamicIntVector relinquish() { return std::move(static_cast<DynamicIntVector&>(*this)); }
    inline DynamicIntVector copy() const { return DynamicIntVector(*this); }
};

class GpuTextDS : public Algorithm {
public:
    using dsflags_t = ds::dsflags_t;
    static const dsflags_t SA = ds::SA;
    static const dsflags_t ISA = ds::ISA;
    static const dsflags_t LCP = ds::LCP;
    static const dsflags_t PHI = ds::PHI;
    static const dsflags_t PLCP = ds::PLCP;
    using value_type = uliteral_t;
    using sa_type = GpuArray;
    using phi_type = GpuArray;
    using plcp_type = GpuArray;
    using lcp_type = GpuArray;
    using isa_type = GpuArray;

    inline static ds::InputRestrictions common_restrictions(dsflags_t) {
        // same as SADivSufSort::restrictions(): escape 0, null-terminate (ds/SADivSufSort.hpp:21-26)
        return ds::InputRestrictions{{0}, true};
    }

    inline static Meta meta() {
        Meta m("textds", "gpu", "Text index built on the GPU (tdcgpu, sm_100a)");
        m.option("compress").dynamic("delayed");
        return m;
    }

private:
    struct CtxDeleter {
        void operator()(tdcgpu_ctx* c) const { gpu_detail::release_ctx(c); }
    };
    View m_text;
    std::unique_ptr<tdcgpu_ctx, CtxDeleter> m_ctx;
    std::unique_ptr<GpuArray> m_sa, m_isa, m_lcp, m_phi, m_plcp;
    CompressMode m_cm;

    inline const GpuArray& fetch(std::unique_ptr<GpuArray>& slot, uint32_t which, const char* title, bool lcp_width) {
        if (!slot) {
            StatPhase::wrap(title, [&] {
                const size_t n = m_text.size();
                gpu_detail::check(tdcgpu_textds_build(m_ctx.get(), which), title);
                gpu_detail::log_phases(m_ctx.get());
                // the device holds 32-bit indices; a wide-index build of the reference (-DLEN_BITS=40: INDEX_FAST_BITS == 64,
                // def.hpp:100-128) gets them widened by the same packing kernel that narrows them for `compress`
                uint32_t mx = 0;
                if (lcp_width) gpu_detail::check(tdcgpu_textds_max_lcp(m_ctx.get(), &mx), title);
                // "delayed" and "compressed" end in the same narrowed state (bits_for(n) / bits_for(max_lcp)); the device
                // packs to that width, so neither 4n bytes cross PCIe nor BitPackingVector::resize re-packs serially
                const uint8_t width = m_cm == CompressMode::plain ? uint8_t(INDEX_FAST_BITS) : uint8_t(lcp_width ? bits_for(mx) : bits_for(n));
                DynamicIntVector iv(n, 0, width);
                if (width == 32) {
                    gpu_detail::check(tdcgpu_textds_get(m_ctx.get(), which, iv.data(), 0), title);
                } else {
                    gpu_detail::check(tdcgpu_textds_get_packed(m_ctx.get(), which, width, iv.data(), (uint64_t(n) * width + 63) / 64, 0), title);
                }
                StatPhase::log("bit_width", size_t(iv.width()));
                StatPhase::log("size", iv.bit_size() / 8);
                slot = std::make_unique<GpuArray>(std::move(iv), len_t(mx));
            });
        }
        return *slot;
    }

public:
    inline GpuTextDS(Env&& env, const View& text) : Algorithm(std::move(env)), m_text(text) {
        if (!m_text.ends_with(uint8_t(0))) {
            throw std::logic_error(
                "Input has no sentinel! Please make sure you declare "
                "the compressor calling this with "
                "`m.needs_sentinel_terminator()` in its `meta()` function.");
        }
        auto& cm_str = this->env().option("compress").as_string();
        m_cm = cm_str == "delayed" ? CompressMode::delayed : (cm_str == "compressed" ? CompressMode::compressed : CompressMode::plain);
        m_ctx.reset(gpu_detail::acquire_ctx());
        gpu_detail::check(tdcgpu_set_text(m_ctx.get(), reinterpret_cast<const uint8_t*>(m_text.data()), m_text.size(), 0), "set_text");
    }

    inline GpuTextDS(Env&& env, const View& text, dsflags_t flags, CompressMode = CompressMode::select)
        : GpuTextDS(std::move(env), text) {
        require(flags);
    }

    /// Builds on the device; nothing is copied to the host until a require_*() asks for it.
    inline void require(dsflags_t flags, CompressMode = CompressMode::select) {
        gpu_detail::check(tdcgpu_textds_build(m_ctx.get(), flags), "require");
        gpu_detail::log_phases(m_ctx.get());
    }

    inline const GpuArray& require_sa(CompressMode = CompressMode::select) { return fetch(m_sa, TDCGPU_SA, "Construct SA", false); }
    inline const GpuArray& require_isa(CompressMode = CompressMode::select) { return fetch(m_isa, TDCGPU_ISA, "Construct ISA", false); }
    inline const GpuArray& require_phi(CompressMode = CompressMode::select) { return fetch(m_phi, TDCGPU_PHI, "Construct Phi Array", false); }
    inline const GpuArray& require_plcp(CompressMode = CompressMode::select) { return fetch(m_plcp, TDCGPU_PLCP, "Construct PLCP Array", true); }
    inline const GpuArray& require_lcp(CompressMode = CompressMode::select) { return fetch(m_lcp, TDCGPU_LCP, "Construct LCP Array", true); }

    inline GpuArray release_sa() { require_sa(); return std::move(*m_sa); }
    inline GpuArray release_isa() { require_isa(); return std::move(*m_isa); }
    inline GpuArray release_phi() { require_phi(); return std::move(*m_phi); }
    inline GpuArray release_plcp() { require_plcp(); return std::move(*m_plcp); }
    inline GpuArray release_lcp() { require_lcp(); return std::move(*m_lcp); }

    inline value_type operator[](size_t i) const { return m_text[i]; }
    inline const value_type* text() const { return m_text.data(); }
    inline size_t size() const { return m_text.size(); }

    /// Device handle for the GPU-side consumers below.
    inline tdcgpu_ctx* device() const { return m_ctx.get(); }
};

// ---------------------------------------------------------------------------------------------------------------------
// lzss_lcp with the device factoriser.  Mirrors LZSSLCPCompressor::compress (compressors/LZSSLCPCompressor.hpp:42-124)
// phase by phase; only the body of "Factorize" differs.
// ---------------------------------------------------------------------------------------------------------------------
#ifdef TDC_GPU_DEFAULT_TEXTDS
#define TDCGPU_DEFAULT_TEXT_T GpuTextDS  /* GPU-only registry: `lzss_lcp(...)` / `bwt` select the GPU text index */
#else
#define TDCGPU_DEFAULT_TEXT_T TextDS<>   /* mixed registry: defaults must equal the reference's (Meta.hpp:303-316) */
#endif

template <typename coder_t>
class LZSSLCPCompressor<coder_t, GpuTextDS> : public Compressor {
    using text_t = GpuTextDS;

public:
    inline static Meta meta() {
        Meta m("compressor", "lzss_lcp", "LZSS Factorization using LCP");
        m.option("coder").templated<coder_t>("coder");
        m.option("textds").templated<text_t, TDCGPU_DEFAULT_TEXT_T>("textds");
        m.option("threshold").dynamic(3);
        m.uses_textds<text_t>(text_t::SA | text_t::ISA | text_t::LCP);
        return m;
    }

    inline LZSSLCPCompressor() = delete;
    inline LZSSLCPCompressor(Env&& env) : Compressor(std::move(env)) {}

    inline virtual void compress(Input& input, Output& output) override {
        auto view = input.as_view();
        DCHECK(view.ends_with(uint8_t(0)));

        text_t text = StatPhase::wrap("Construct Text DS", [&] {
            return text_t(env().env_for_option("textds"), view, text_t::SA | text_t::ISA | text_t::LCP);
        });

        const bool device_encode = gpu_detail::DeviceLiteralCoder<coder_t>::supported && !gpu_detail::host_encode_forced();
        lzss::FactorBuffer factors;
        StatPhase::wrap("Factorize", [&] {
            const len_t threshold = env().option("threshold").as_integer();
            if (threshold < 1) throw std::runtime_error("lzss_lcp: threshold must be >= 1");
            uint64_t z = 0;
            uint32_t mn = 0, mx = 0;
            gpu_detail::check(tdcgpu_lzss_lcp_factorize(text.device(), threshold, &z, &mn, &mx), "factorize");
            gpu_detail::log_phases(text.device());
            if (!device_encode) {  // the host encoder needs the list; the device encoder reads it where it is
                std::unique_ptr<tdcgpu_factor[]> buf(new tdcgpu_factor[z ? z : 1]);
                gpu_detail::check(tdcgpu_lzss_lcp_get_factors(text.device(), buf.get(), z, 0), "get_factors");
                for (uint64_t k = 0; k < z; k++) factors.emplace_back(buf[k].pos, buf[k].src, buf[k].len);
            }
            StatPhase::log("threshold", threshold);
            StatPhase::log("factors", size_t(z));
        });

        StatPhase::wrap("Encode", [&] {
            if (!device_encode) {
                typename coder_t::Encoder coder(env().env_for_option("coder"), output, lzss::TextLiterals<text_t>(text, factors));
                lzss::encode_text(coder, text, factors);
                return;
            }
            // 1. literal counts from the device -> the coder's own header and code words (reference code, host)
            uint64_t hist[256];
            gpu_detail::check(tdcgpu_lzss_literal_histogram(text.device(), hist, nullptr), "literal histogram");
            gpu_detail::log_phases(text.device());
            gpu_detail::LiteralCodeTable table;
            std::vector<uint8_t> head;
            {
                Output scratch = Output::from_memory(head);
                io::BitOStream bits(scratch);
                gpu_detail::DeviceLiteralCoder<coder_t>::header(bits, hist, table);
            }
            const uint64_t head_bits = gpu_detail::strip_bitstream_tail(head);
            const uint32_t lead_bits = uint32_t(head_bits % 8);
            const uint8_t lead_byte = lead_bits ? head[head_bits / 8] : uint8_t(0);
            // 2. the body of lzss::encode_text on the device, continuing the header's partial byte
            uint64_t nbits = 0, nbytes = 0;
            gpu_detail::check(tdcgpu_set_len_bits(text.device(), uint32_t(8 * sizeof(len_t))), "encode");  // LengthRange of THIS build (LEN_BITS)
            gpu_detail::check(tdcgpu_lzss_encode(text.device(), table.codes, table.lens, lead_bits, lead_byte, &nbits), "encode");
            gpu_detail::log_phases(text.device());
            // 3. drain the stream through one pinned buffer straight into the output (no archive-sized vector in between)
            auto os = output.as_stream();
            os.write(reinterpret_cast<const char*>(head.data()), std::streamsize(head_bits / 8));
            gpu_detail::PinnedBuffer& buf = gpu_detail::drain_buffer();
            for (uint64_t off = 0;;) {
                uint64_t total = 0, wr = 0;
                gpu_detail::check(tdcgpu_lzss_encode_get_chunk(text.device(), off, buf.data, buf.size, 1, &total, &wr), "encode");
                if (wr == 0) break;
                os.write(reinterpret_cast<const char*>(buf.data), std::streamsize(wr));
                off += wr;
                nbytes = total;
            }
            StatPhase::log("archive_bytes", size_t(head_bits / 8 + nbytes));
        });
    }

    inline virtual void decompress(Input& input, Output& output) override {
        typename coder_t::Decoder decoder(env().env_for_option("coder"), input);
        auto outs = output.as_stream();
        lzss::decode_text<typename coder_t::Decoder, lzss::DecodeBackBuffer>(decoder, outs);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// bwt with the device gather (compressors/BWTCompressor.hpp:29-47).  decompress is the reference's.
// ---------------------------------------------------------------------------------------------------------------------
template <>
class BWTCompressor<GpuTextDS> : public Compressor, public gpu_detail::DeviceStage {
    using text_t = GpuTextDS;

public:
    inline static Meta meta() {
        Meta m("compressor", "bwt", "BWT Compressor");
        m.option("textds").templated<text_t, TDCGPU_DEFAULT_TEXT_T>("textds");
        m.uses_textds<text_t>(ds::SA);
        return m;
    }

    using Compressor::Compressor;

    inline virtual void compress(Input& input, Output& output) override {
        gpu_detail::StageInput in;
        in.host = &input;
        compress_stage(in, nullptr, &output);
    }

    /// The BWT stays in device memory when the next stage of a chain runs on the GPU as well (dev_out != nullptr).
    inline void compress_stage(const gpu_detail::StageInput& sin, gpu_detail::DeviceBytes* dev_out, Output* host_out) override {
        if (!sin.host) throw std::runtime_error("bwt: the text index is built from a host text");
        auto in = sin.host->as_view();
        DCHECK(in.ends_with(uint8_t(0)));
        text_t t(env().env_for_option("textds"), in);
        if (dev_out) {
            StatPhase::wrap("Construct Text DS", [&] {
                gpu_detail::check(tdcgpu_textds_build(t.device(), TDCGPU_SA | TDCGPU_BWT), "bwt");
                gpu_detail::log_phases(t.device());
                dev_out->alloc(t.device(), t.size());
                gpu_detail::check(tdcgpu_textds_get(t.device(), TDCGPU_BWT, dev_out->p, 1), "bwt");
                dev_out->n = t.size();
            });
            return;
        }
        auto ostream = host_out->as_stream();
        std::string bwt(t.size(), 0);
        StatPhase::wrap("Construct Text DS", [&] {
            gpu_detail::check(tdcgpu_textds_build(t.device(), TDCGPU_SA | TDCGPU_BWT), "bwt");
            gpu_detail::log_phases(t.device());
            gpu_detail::check(tdcgpu_textds_get(t.device(), TDCGPU_BWT, &bwt[0], 0), "bwt");
        });
        ostream.write(bwt.data(), bwt.size());
    }

    inline virtual void decompress(Input& input, Output& output) override {
        auto in = input.as_view();
        auto ostream = output.as_stream();
        auto decoded_string = StatPhase::wrap("Decode BWT", [&] { return bwt::decode_bwt(in); });
        if (tdc_unlikely(decoded_string.empty())) return;
        StatPhase::wrap("Output Text", [&] { ostream << decoded_string << '\0'; });
    }
};

}  // namespace tdc
