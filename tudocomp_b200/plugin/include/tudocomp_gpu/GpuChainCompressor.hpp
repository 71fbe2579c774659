// GPU-aware drop-in for the chain compressor, Meta("compressor", "chain") with the options of
// tudocomp_driver/ChainCompressor.hpp:16-21 — `A:B` in an algorithm string is `chain(A, B)`.
//
// The reference chain runs `first` into a host vector and `second` from it (ChainCompressor.hpp:54-62), so with GPU stages
// every link of `bwt:mtf:rle:encode(huff)` crosses PCIe twice (205 + 89 + 40 + 17 ms of wall time for ~11 ms of stage
// kernels at 2^28 B, profiles/r1o_summary.md).  Here: when both sides of a link implement gpu_detail::DeviceStage
// (GpuTextDS.hpp) the bytes between them stay in device memory; the chain is itself a DeviceStage, so nested chains
// (`a:b:c:d` = chain(a, chain(b, chain(c, d)))) keep the whole pipeline in HBM: one upload of the text, one download of
// the coded stream.  Any other combination takes the reference's route through a host buffer.  Same archive bytes either
// way; decompress() is the reference's order of operations (second, then first).
// It replaces ChainCompressor in the GPU-only registry (plugin/registry_gpu.py) — same (type, name), so it cannot coexist.
#pragma once

#include <cstring>
#include <memory>
#include <vector>

#include <tudocomp/Compressor.hpp>
#include <tudocomp/CreateAlgorithm.hpp>
#include <tudocomp/Env.hpp>
#include <tudocomp/Registry.hpp>
#include <tudocomp/io.hpp>
#include <tudocomp_driver/Registry.hpp>

#include "GpuTextDS.hpp"

namespace tdc {

class GpuChainCompressor : public Compressor, public gpu_detail::DeviceStage {
    struct Link {
        std::unique_ptr<Compressor> algo;
        ds::InputRestrictionsAndFlags flags;
        gpu_detail::DeviceStage* stage = nullptr;  // non-null: the algorithm can take / leave its bytes on the device
    };

    inline Link make(const char* option) {
        auto& option_value = env().option(option);
        DCHECK(option_value.is_algorithm());
        auto av = option_value.as_algorithm();
        Link l;
        l.flags = av.textds_flags();
        l.algo = create_algo_with_registry_dynamic(tdc_algorithms::COMPRESSOR_REGISTRY, av);
        l.stage = dynamic_cast<gpu_detail::DeviceStage*>(l.algo.get());
        return l;
    }
    static inline bool host_chain_forced() {
        const char* e = std::getenv("TDCGPU_HOST_CHAIN");  // debug / A-B switch: the reference's host buffer between all stages
        return e && *e == '1';
    }
    // one link on a HOST input, with the driver's input restrictions (escaping / sentinel) if the algorithm asks for them
    static inline void run_host_in(Link& l, Input& in, gpu_detail::DeviceBytes* dev_out, Output* host_out) {
        auto go = [&](Input& i) {
            if (l.stage) {
                gpu_detail::StageInput si;
                si.host = &i;
                l.stage->compress_stage(si, dev_out, host_out);
            } else {
                l.algo->compress(i, *host_out);
            }
        };
        if (!l.flags.has_restrictions()) {
            go(in);
            return;
        }
        // The text-index restrictions {escape 0x00, append the sentinel} on bytes that contain neither 0x00 nor the escape
        // byte 0xFF are "the same bytes plus one 0" (io/EscapeMap.hpp:39-64, io/RestrictedBuffer.hpp:108-140): two memchr
        // scans and one memcpy instead of the restricted Input's byte loops (~4 s per GiB in front of ~0.7 s of GPU work
        // for `bwt:mtf:rle:encode(huff)`, profiles/r2_summary.md).  Anything else takes the reference's route.
        if (l.stage && l.flags.null_terminate() && l.flags.escape_bytes() == std::vector<uint8_t>{0}) {
            auto v = in.as_view();
            const bool clean = v.size() == 0 || (std::memchr(v.data(), 0x00, v.size()) == nullptr && std::memchr(v.data(), 0xFF, v.size()) == nullptr);
            if (clean) {
                gpu_detail::PinnedBuffer text(v.size() + 1);  // page-locked: the upload is one DMA
                if (v.size()) std::memcpy(text.data, v.data(), v.size());
                text.data[v.size()] = 0;
                Input direct(View(text.data, v.size() + 1));
                go(direct);
                return;
            }
        }
        Input restricted(in, l.flags);
        go(restricted);
    }

public:
    inline static Meta meta() {
        Meta m("compressor", "chain");
        m.option("first").dynamic_compressor();
        m.option("second").dynamic_compressor();
        return m;
    }

    inline GpuChainCompressor() = delete;
    inline GpuChainCompressor(Env&& env) : Compressor(std::move(env)) {}

    inline virtual void compress(Input& input, Output& output) override final {
        gpu_detail::StageInput in;
        in.host = &input;
        compress_stage(in, nullptr, &output);
    }

    inline void compress_stage(const gpu_detail::StageInput& in, gpu_detail::DeviceBytes* dev_out, Output* host_out) override final {
        Link first = make("first"), second = make("second");
        // a device-resident input can only feed a stage without input restrictions (they are defined on host bytes)
        const bool first_takes_dev = first.stage && !first.flags.has_restrictions();
        const bool second_takes_dev = second.stage && !second.flags.has_restrictions();
        std::vector<uint8_t> host_in;  // only if a device-resident input has to come back for a host-only first stage
        Input host_input;
        Input* hin = in.host;
        if (in.dev && !first_takes_dev) {
            gpu_detail::download(*in.dev, host_in);
            host_input = Input(host_in);
            hin = &host_input;
        }
        auto run_first = [&](gpu_detail::DeviceBytes* d, Output* o) {
            if (hin) run_host_in(first, *hin, d, o);
            else first.stage->compress_stage(in, d, o);
        };
        if (first.stage && second_takes_dev && !host_chain_forced()) {
            gpu_detail::DeviceBytes mid;  // the bytes between the two links, in HBM
            run_first(&mid, nullptr);
            gpu_detail::StageInput in2;
            in2.dev = &mid;
            second.stage->compress_stage(in2, dev_out, host_out);
            return;
        }
        // the reference's route: a host buffer in between (ChainCompressor.hpp:54-62)
        std::vector<uint8_t> between_buf;
        {
            Output between(between_buf);
            run_first(nullptr, &between);
        }
        Input between(between_buf);
        if (!dev_out) {
            run_host_in(second, between, nullptr, host_out);
        } else if (second.stage) {
            run_host_in(second, between, dev_out, nullptr);
        } else {  // a host-only last stage inside a device-resident outer chain: hand its bytes up
            std::vector<uint8_t> coded;
            {
                Output mem(coded);
                run_host_in(second, between, nullptr, &mem);
            }
            gpu_detail::upload(*dev_out, coded);
        }
    }

    inline virtual void decompress(Input& input, Output& output) override final {
        Link first = make("first"), second = make("second");
        auto run = [](Link& l, Input& i, Output& o) {
            if (l.flags.has_restrictions()) {
                Output restricted(o, l.flags);
                l.algo->decompress(i, restricted);
            } else {
                l.algo->decompress(i, o);
            }
        };
        std::vector<uint8_t> between_buf;
        {
            Output between(between_buf);
            run(second, input, between);
        }
        {
            Input between(between_buf);
            run(first, between, output);
        }
    }
};

}  // namespace tdc
