// GPU-backed drop-ins for the stream stages behind the BWT in `bwt:mtf:rle:encode(huff)` (BASELINE config 3):
//
//   GpuMTFCompressor        = MTFCompressor          (compressors/MTFCompressor.hpp:71-92),   Meta("compressor","mtf")
//   GpuRunLengthEncoder     = RunLengthEncoder       (compressors/RunLengthEncoder.hpp:51-76), Meta("compressor","rle"), option offset
//   GpuLiteralEncoder<C>    = LiteralEncoder<C>      (compressors/LiteralEncoder.hpp:11-45),   Meta("compressor","encode"), option coder
//
// Same names, options and archive bytes as the reference classes, so they can only REPLACE them in a registry (the GPU-only
// one, plugin/registry_gpu.py) — two classes under one (type, name) would collide (Meta.hpp:303-316).  compress() calls
// the C ABI (include/tdcgpu.h); decompress() is the reference's own code.  The reference's chain compressor hands host
// buffers from stage to stage (tudocomp_driver/ChainCompressor.hpp:56-66); every class here also implements
// gpu_detail::DeviceStage, so that the GPU-aware chain (GpuChainCompressor.hpp) can leave the bytes in HBM in between.
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
// where the stage's in / out buffers are, in the C ABI's terms
inline int buf_mode(bool in_dev, bool out_dev) {
    return in_dev ? (out_dev ? TDCGPU_BUF_DEVICE : TDCGPU_BUF_IN_DEVICE) : (out_dev ? TDCGPU_BUF_OUT_DEVICE : TDCGPU_BUF_HOST);
}
struct StageBytes {  // the input of a stage as (pointer, length), host or device
    const uint8_t* p = nullptr;
    uint64_t n = 0;
    bool dev = false;
    std::unique_ptr<io::InputView> host_view;  // owns the host bytes (a mapping or a copy) for the duration of the stage
    explicit StageBytes(const StageInput& in) {
        if (in.dev) { p = in.dev->p; n = in.dev->n; dev = true; }
        else {
            host_view = std::make_unique<io::InputView>(in.host->as_view());
            p = reinterpret_cast<const uint8_t*>(host_view->data());
            n = host_view->size();
        }
    }
};
}  // namespace gpu_detail

class GpuMTFCompressor : public Compressor, public gpu_detail::DeviceStage {
public:
    inline static Meta meta() {
        Meta m("compressor", "mtf", "Move To Front Compressor");
        return m;
    }
    inline GpuMTFCompressor(Env&& env) : Compressor(std::move(env)) {}

    inline virtual void compress(Input& input, Output& output) override {
        gpu_detail::StageInput in;
        in.host = &input;
        compress_stage(in, nullptr, &output);
    }
    inline void compress_stage(const gpu_detail::StageInput& sin, gpu_detail::DeviceBytes* dev_out, Output* host_out) override {
        gpu_detail::StageBytes in(sin);
        gpu_detail::StreamCtx g;
        std::vector<uint8_t> out(dev_out ? 0 : in.n);
        if (dev_out) { dev_out->alloc(g.ctx, in.n); dev_out->n = in.n; }
        if (in.n) {
            gpu_detail::check(tdcgpu_mtf_encode(g.ctx, in.p, in.n, dev_out ? dev_out->p : out.data(), gpu_detail::buf_mode(in.dev, dev_out != nullptr)), "mtf");
            gpu_detail::log_phases(g.ctx);
        }
        if (!dev_out) {
            auto os = host_out->as_stream();
            os.write(reinterpret_cast<const char*>(out.data()), std::streamsize(out.size()));
        }
    }
    inline virtual void decompress(Input& input, Output& output) override {
        auto is = input.as_stream();
        auto os = output.as_stream();
        mtf_decode(is, os);
    }
};

class GpuRunLengthEncoder : public Compressor, public gpu_detail::DeviceStage {
public:
    inline static Meta meta() {
        Meta m("compressor", "rle", "Run Length Encoding Compressor");
        m.option("offset").dynamic(0);
        return m;
    }
    const size_t m_offset;
    inline GpuRunLengthEncoder(Env&& env) : Compressor(std::move(env)), m_offset(this->env().option("offset").as_integer()) {}

    inline virtual void compress(Input& input, Output& output) override {
        gpu_detail::StageInput in;
        in.host = &input;
        compress_stage(in, nullptr, &output);
    }
    inline void compress_stage(const gpu_detail::StageInput& sin, gpu_detail::DeviceBytes* dev_out, Output* host_out) override {
        gpu_detail::StageBytes in(sin);
        if (in.n == 0) {
            if (!dev_out) host_out->as_stream();
            return;
        }
        // worst case of the reference's format: every byte followed by a vbyte (bytes >= 0x80 are never merged)
        size_t vl = 1;
        for (uint64_t v = uint64_t(m_offset) + in.n; v >>= 7;) vl++;
        const uint64_t worst = in.n * (1 + vl) + 16;
        uint64_t produced = 0;
        gpu_detail::StreamCtx g;
        if (dev_out) {
            dev_out->alloc(g.ctx, worst);
            gpu_detail::check(tdcgpu_rle_encode(g.ctx, in.p, in.n, m_offset, dev_out->p, worst, &produced, gpu_detail::buf_mode(in.dev, true)), "rle");
            gpu_detail::log_phases(g.ctx);
            dev_out->n = produced;
            return;
        }
        std::vector<uint8_t> out(worst);
        gpu_detail::check(tdcgpu_rle_encode(g.ctx, in.p, in.n, m_offset, out.data(), out.size(), &produced, gpu_detail::buf_mode(in.dev, false)), "rle");
        gpu_detail::log_phases(g.ctx);
        auto os = host_out->as_stream();
        os.write(reinterpret_cast<const char*>(out.data()), std::streamsize(produced));
    }
    inline virtual void decompress(Input& input, Output& output) override {
        auto is = input.as_stream();
        auto os = output.as_stream();
        rle_decode(is, os, m_offset);
    }
};

template <typename coder_t>
class GpuLiteralEncoder : public Compressor, public gpu_detail::DeviceStage {
public:
    inline static Meta meta() {
        Meta m("compressor", "encode", "Simply encodes the input's individual characters.");
        m.option("coder").templated<coder_t>("coder");
        return m;
    }
    inline GpuLiteralEncoder(Env&& env) : Compressor(std::move(env)) {}

    inline virtual void compress(Input& input, Output& output) override final {
        gpu_detail::StageInput in;
        in.host = &input;
        compress_stage(in, nullptr, &output);
    }
    inline void compress_stage(const gpu_detail::StageInput& sin, gpu_detail::DeviceBytes* dev_out, Output* host_out) override final {
        if (!gpu_detail::DeviceLiteralCoder<coder_t>::supported || gpu_detail::host_encode_forced() || dev_out) {
            // LiteralEncoder::compress as it is (compressors/LiteralEncoder.hpp:23-32), on a host copy of the input if needed
            // (a coded stream that has to stay on the device — `encode` in the middle of a chain — takes this path too)
            std::vector<uint8_t> copy, coded;
            std::unique_ptr<io::InputView> hv;
            View iview("");
            if (sin.dev) {
                gpu_detail::download(*sin.dev, copy);
                iview = View(copy);
            } else {
                hv = std::make_unique<io::InputView>(sin.host->as_view());
                iview = View(*hv);
            }
            {
                Output mem = Output::from_memory(coded);
                Output& o = dev_out ? mem : *host_out;
                typename coder_t::Encoder coder(env().env_for_option("coder"), o, ViewLiterals(iview));
                for (uint8_t c : iview) coder.encode(c, literal_r);
            }
            if (dev_out) {
                gpu_detail::StreamCtx g;
                dev_out->alloc(g.ctx, coded.size());
                gpu_detail::check(tdcgpu_device_copy(g.ctx, dev_out->p, coded.data(), coded.size(), 0), "encode");
                dev_out->n = coded.size();
            }
            return;
        }
        gpu_detail::StageBytes in(sin);
        gpu_detail::StreamCtx g;
        uint64_t hist[256];
        gpu_detail::check(tdcgpu_literal_encode_begin(g.ctx, in.p, in.n, in.dev ? 1 : 0, hist), "encode");
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
        uint64_t nbits = 0;
        gpu_detail::check(tdcgpu_literal_encode(g.ctx, table.codes, table.lens, lead_bits, lead_byte, &nbits), "encode");
        gpu_detail::log_phases(g.ctx);
        auto os = host_out->as_stream();
        os.write(reinterpret_cast<const char*>(head.data()), std::streamsize(head_bits / 8));
        gpu_detail::PinnedBuffer& buf = gpu_detail::drain_buffer();  // drained chunk by chunk, no stream-sized vector in between
        for (uint64_t off = 0;;) {
            uint64_t total = 0, wr = 0;
            gpu_detail::check(tdcgpu_literal_encode_get_chunk(g.ctx, off, buf.data, buf.size, 1, &total, &wr), "encode");
            if (wr == 0) break;
            os.write(reinterpret_cast<const char*>(buf.data), std::streamsize(wr));
            off += wr;
        }
    }

    inline virtual void decompress(Input& input, Output& output) override final {
        auto ostream = output.as_stream();
        typename coder_t::Decoder decoder(env().env_for_option("coder"), input);
        while (!decoder.eof()) ostream << decoder.template decode<uint8_t>(literal_r);
    }
};

}  // namespace tdc
