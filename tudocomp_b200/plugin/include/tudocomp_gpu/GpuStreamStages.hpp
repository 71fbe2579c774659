// GPU-backed drop-ins for the stream stages behind the BWT in `bwt:mtf:rle:encode(huff)` (BASELINE config 3):
//
//   GpuMTFCompressor        = MTFCompressor          (compressors/MTFCompressor.hpp:71-92),   Meta("compressor","mtf")
//   GpuRunLengthEncoder     = RunLengthEncoder       (compressors/RunLengthEncoder.hpp:51-76), Meta("compressor","rle"), option offset
//   GpuLiteralEncoder<C>    = LiteralEncoder<C>      (compressors/LiteralEncoder.hpp:11-45),   Meta("compressor","encode"), option coder
//
// Same names, options and archive bytes as the reference classes, so they can only REPLACE them in a registry (the GPU-only
// one, plugin/registry_gpu.py) — two classes under one (type, name) would collide (Meta.hpp:303-316).  compress() calls
// the C ABI (include/tdcgpu.h); decompress() is the reference's own code.  The chain compressor hands host buffers from
// stage to stage (tudocomp_driver/ChainCompressor.hpp:56-66), so every stage copies in and out over PCIe.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <tudocomp/Compressor.hpp>
#include <tudocomp/compressors/LiteralEncoder.hpp>
#include <tudocomp/compressors/MTFCompressor.hpp>
#include <tudocomp/compressors/RunLengthEncoder.hpp>
#include <tudocomp_stat/StatPhase.hpp>

#include "GpuTextDS.hpp"

namespace tdc {

namespace gpu_detail {
struct StreamCtx {  // one context (stream + scratch) for the duration of a compress() call
    tdcgpu_ctx* ctx = nullptr;
    StreamCtx() : ctx(acquire_ctx()) {}
    ~StreamCtx() { release_ctx(ctx); }
    StreamCtx(const StreamCtx&) = delete;
    StreamCtx& operator=(const StreamCtx&) = delete;
};
}  // namespace gpu_detail

class GpuMTFCompressor : public Compressor {
public:
    inline static Meta meta() {
        Meta m("compressor", "mtf", "Move To Front Compressor");
        return m;
    }
    inline GpuMTFCompressor(Env&& env) : Compressor(std::move(env)) {}

    inline virtual void compress(Input& input, Output& output) override {
        auto in = input.as_view();
        std::vector<uint8_t> out(in.size());
        if (!in.empty()) {
            gpu_detail::StreamCtx g;
            gpu_detail::check(tdcgpu_mtf_encode(g.ctx, reinterpret_cast<const uint8_t*>(in.data()), in.size(), out.data(), 0), "mtf");
            gpu_detail::log_phases(g.ctx);
        }
        auto os = output.as_stream();
        os.write(reinterpret_cast<const char*>(out.data()), std::streamsize(out.size()));
    }
    inline virtual void decompress(Input& input, Output& output) override {
        auto is = input.as_stream();
        auto os = output.as_stream();
        mtf_decode(is, os);
    }
};

class GpuRunLengthEncoder : public Compressor {
public:
    inline static Meta meta() {
        Meta m("compressor", "rle", "Run Length Encoding Compressor");
        m.option("offset").dynamic(0);
        return m;
    }
    const size_t m_offset;
    inline GpuRunLengthEncoder(Env&& env) : Compressor(std::move(env)), m_offset(this->env().option("offset").as_integer()) {}

    inline virtual void compress(Input& input, Output& output) override {
        auto in = input.as_view();
        auto os = output.as_stream();
        if (in.empty()) return;
        // worst case of the reference's format: every byte followed by a vbyte (bytes >= 0x80 are never merged)
        size_t vl = 1;
        for (uint64_t v = uint64_t(m_offset) + in.size(); v >>= 7;) vl++;
        std::vector<uint8_t> out(in.size() * (1 + vl) + 16);
        uint64_t produced = 0;
        gpu_detail::StreamCtx g;
        gpu_detail::check(tdcgpu_rle_encode(g.ctx, reinterpret_cast<const uint8_t*>(in.data()), in.size(), m_offset, out.data(), out.size(), &produced, 0), "rle");
        gpu_detail::log_phases(g.ctx);
        os.write(reinterpret_cast<const char*>(out.data()), std::streamsize(produced));
    }
    inline virtual void decompress(Input& input, Output& output) override {
        auto is = input.as_stream();
        auto os = output.as_stream();
        rle_decode(is, os, m_offset);
    }
};

template <typename coder_t>
class GpuLiteralEncoder : public Compressor {
public:
    inline static Meta meta() {
        Meta m("compressor", "encode", "Simply encodes the input's individual characters.");
        m.option("coder").templated<coder_t>("coder");
        return m;
    }
    inline GpuLiteralEncoder(Env&& env) : Compressor(std::move(env)) {}

    inline virtual void compress(Input& input, Output& output) override final {
        auto iview = input.as_view();
        if (!gpu_detail::DeviceLiteralCoder<coder_t>::supported || gpu_detail::host_encode_forced()) {
            // LiteralEncoder::compress as it is (compressors/LiteralEncoder.hpp:23-32)
            typename coder_t::Encoder coder(env().env_for_option("coder"), output, ViewLiterals(iview));
            for (uint8_t c : iview) coder.encode(c, literal_r);
            return;
        }
        gpu_detail::StreamCtx g;
        uint64_t hist[256];
        gpu_detail::check(tdcgpu_literal_encode_begin(g.ctx, reinterpret_cast<const uint8_t*>(iview.data()), iview.size(), 0, hist), "encode");
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
        uint64_t nbits = 0, nbytes = 0;
        gpu_detail::check(tdcgpu_literal_encode(g.ctx, table.codes, table.lens, lead_bits, lead_byte, &nbits), "encode");
        gpu_detail::log_phases(g.ctx);
        std::vector<uint8_t> body(nbits / 8 + 2);
        gpu_detail::check(tdcgpu_literal_encode_get(g.ctx, body.data(), body.size(), 1, &nbytes, 0), "encode");
        auto os = output.as_stream();
        os.write(reinterpret_cast<const char*>(head.data()), std::streamsize(head_bits / 8));
        os.write(reinterpret_cast<const char*>(body.data()), std::streamsize(nbytes));
    }

    inline virtual void decompress(Input& input, Output& output) override final {
        auto ostream = output.as_stream();
        typename coder_t::Decoder decoder(env().env_for_option("coder"), input);
        while (!decoder.eof()) ostream << decoder.template decode<uint8_t>(literal_r);
    }
};

}  // namespace tdc
