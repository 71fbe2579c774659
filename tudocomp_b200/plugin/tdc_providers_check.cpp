// Checker for the per-provider GPU classes (tudocomp_gpu/GpuProviders.hpp): the reference's own
// TextDS<GpuSA, GpuPhi, GpuPLCP, GpuLCP, GpuISA> against TextDS<> (all-CPU providers) on the same text, array by array,
// for every CompressMode.  Used by tests/test_plugin.py (over the interpreter library on the CPU, on the device with -m gpu).
//
//   tdc_providers_check FILE
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include <tudocomp/CreateAlgorithm.hpp>
#include <tudocomp/ds/TextDS.hpp>
#include <tudocomp/io.hpp>
#include <tudocomp_gpu/GpuProviders.hpp>
#include <tudocomp_stat/StatPhase.hpp>

using namespace tdc;
using GpuProvidersTextDS = TextDS<GpuSA, GpuPhi, GpuPLCP, GpuLCP, GpuISA>;

template <class A, class B>
static bool same(const char* what, const char* mode, const A& a, const B& b) {
    if (a.size() != b.size()) { std::fprintf(stderr, "%s (%s): sizes differ\n", what, mode); return false; }
    for (size_t i = 0; i < a.size(); i++)
        if (uint64_t(a[i]) != uint64_t(b[i])) { std::fprintf(stderr, "%s (%s): [%zu] = %llu, reference %llu\n", what, mode, i, (unsigned long long)uint64_t(a[i]), (unsigned long long)uint64_t(b[i])); return false; }
    if (a.width() != b.width()) { std::fprintf(stderr, "%s (%s): width %u, reference %u\n", what, mode, unsigned(a.width()), unsigned(b.width())); return false; }
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s FILE\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
        if (!f) throw std::runtime_error("cannot open input");
        std::vector<uint8_t> raw(size_t(f.tellg()));
        f.seekg(0);
        f.read(reinterpret_cast<char*>(raw.data()), std::streamsize(raw.size()));
        Input plain{View(raw.data(), raw.size())};
        Input restricted(plain, io::InputRestrictions({0}, true));
        auto holder = restricted.as_view();
        View text = holder;
        bool ok = true;
        for (const char* mode : {"plain", "delayed", "compressed"}) {
            StatPhase root("root");
            const std::string opt = std::string("compress=\"") + mode + "\"";
            auto cpu = create_algo<TextDS<>>(opt, text);
            auto gpu = create_algo<GpuProvidersTextDS>(opt, text);
            const ds::dsflags_t all = ds::SA | ds::ISA | ds::PHI | ds::PLCP | ds::LCP;
            // one structure at a time, so that nothing is moved out from under a later comparison (inplace_phi)
            ok = same("sa", mode, gpu.require_sa(), cpu.require_sa()) && ok;
            ok = same("isa", mode, gpu.require_isa(), cpu.require_isa()) && ok;
            ok = same("lcp", mode, gpu.require_lcp(), cpu.require_lcp()) && ok;
            ok = (gpu.require_lcp().max_lcp() == cpu.require_lcp().max_lcp()) && ok;
            ok = same("plcp", mode, gpu.require_plcp(), cpu.require_plcp()) && ok;
            ok = (gpu.require_plcp().max_lcp() == cpu.require_plcp().max_lcp()) && ok;
            (void)all;
        }
        // Phi on fresh objects: the CPU PLCP provider consumes Phi in place (PLCPFromPhi.hpp:33)
        {
            StatPhase root("root");
            auto cpu = create_algo<TextDS<>>("", text);
            auto gpu = create_algo<GpuProvidersTextDS>("", text);
            ok = same("phi", "delayed", gpu.require_phi(), cpu.require_phi()) && ok;
        }
        std::printf("{\"providers_equal\": %s, \"n\": %zu}\n", ok ? "true" : "false", size_t(text.size()));
        return ok ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
