// End-to-end harness over the C++ plugin: the SAME classes the `tdc` driver instantiates from the generated registry
// (LZSSLCPCompressor<coder_t, GpuTextDS>, tudocomp_gpu/GpuTextDS.hpp) driven K times in one process on an in-memory
// Input / Output, as a long-running service or `tdc_block` worker would.  bench.py reports its numbers as `plugin_e2e`.
//
//   tdc_plugin_bench FILE bit|huff THRESHOLD STEPS 1|0 [ARCHIVE_OUT]
//
// The file is read once (untimed); the driver's own input preparation — {0}-escaping + sentinel through
// Input(in, InputRestrictions({0}, true)), src/tudocomp_driver/tudocomp_driver.cpp:268-270 — is done once (untimed) so
// that the timed region is exactly Compressor::compress(Input&, Output&): pageable text in, archive bytes out.
// Last argument 0 runs the reference's CPU TextDS<> instead (same binary): the byte-identity checker and CPU baseline.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include <tudocomp/CreateAlgorithm.hpp>
#include <tudocomp/coders/BitCoder.hpp>
#include <tudocomp/coders/HuffmanCoder.hpp>
#include <tudocomp/compressors/LZSSLCPCompressor.hpp>
#include <tudocomp/io.hpp>
#include <tudocomp_gpu/GpuTextDS.hpp>
#include <tudocomp_stat/StatPhase.hpp>

using namespace tdc;

static std::string g_phases;  // StatPhase JSON of the last run (per-phase wall times of compress())
static std::string g_out_path;

// The archive goes to a FILE like in the driver (Output::from_path, src/tudocomp_driver/tudocomp_driver.cpp:232-247): the
// reference's in-memory Output appends byte by byte (1.4 s for a 350 MB archive), which no tdc run ever pays.
template <class C>
static std::vector<uint8_t> run(View text, const std::string& opts, double* ms) {
    {
        Input in(text);
        Output o = Output::from_path(io::Path{g_out_path}, true);
        auto c = create_algo<C>(opts);
        StatPhase root("root");
        auto t0 = std::chrono::steady_clock::now();
        c.compress(in, o);
        auto t1 = std::chrono::steady_clock::now();
        *ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        g_phases = root.to_json().str();
    }
    std::ifstream f(g_out_path, std::ios::binary | std::ios::ate);
    std::vector<uint8_t> out(size_t(f.tellg()));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(out.data()), std::streamsize(out.size()));
    return out;
}

static uint64_t fnv1a(const std::vector<uint8_t>& v) {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : v) { h ^= b; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s FILE bit|huff THRESHOLD STEPS 1|0 [ARCHIVE_OUT]\n", argv[0]); return 2; }
    const std::string path = argv[1], coder = argv[2], opts = std::string("threshold=") + argv[3];
    const int steps = std::atoi(argv[4]);
    const bool gpu = std::atoi(argv[5]) != 0;
    g_out_path = argc > 6 ? std::string(argv[6]) : path + ".plugin_out";
    try {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (!f) throw std::runtime_error("cannot open " + path);
        std::vector<uint8_t> raw(size_t(f.tellg()));
        f.seekg(0);
        f.read(reinterpret_cast<char*>(raw.data()), std::streamsize(raw.size()));
        Input plain{View(raw.data(), raw.size())};
        Input restricted(plain, io::InputRestrictions({0}, true));
        auto holder = restricted.as_view();  // owns the escaped copy (io/InputView.hpp); must outlive `text`
        View text = holder;                  // escaped + sentinel: what uses_textds compressors receive
        std::vector<uint8_t> arc;
        std::vector<double> times;
        for (int s = 0; s < steps + 1; s++) {  // the first run is the warm-up (context creation, first-touch allocations)
            double ms = 0;
            if (gpu && coder == "huff") arc = run<LZSSLCPCompressor<HuffmanCoder, GpuTextDS>>(text, opts, &ms);
            else if (gpu) arc = run<LZSSLCPCompressor<BitCoder, GpuTextDS>>(text, opts, &ms);
            else if (coder == "huff") arc = run<LZSSLCPCompressor<HuffmanCoder>>(text, opts, &ms);
            else arc = run<LZSSLCPCompressor<BitCoder>>(text, opts, &ms);
            times.push_back(ms);
            if (!gpu) break;  // the CPU reference is timed once
        }
        double sum = 0;
        for (size_t i = times.size() > 1 ? 1 : 0; i < times.size(); i++) sum += times[i];
        const double mean = sum / double(times.size() > 1 ? times.size() - 1 : 1);
        std::vector<double> timed(times.begin() + (times.size() > 1 ? 1 : 0), times.end());
        std::string runs;
        for (double t : timed) { char b[32]; std::snprintf(b, sizeof b, "%s%.1f", runs.empty() ? "" : ", ", t); runs += b; }
        std::sort(timed.begin(), timed.end());
        const double median = timed[timed.size() / 2];  // the mean is what counts; the runs show how much of it is page-cache noise of the file Output
        if (argc <= 6) std::remove(g_out_path.c_str());
        std::string top;  // "title": ms of the first-level phases of the last run
        {
            // the JSON is {"title":..,"timeStart":..,"timeEnd":..,"sub":[{...},...]}: pull (title, timeEnd - timeStart) of depth 1
            int depth = 0;
            std::string title;
            double ts = 0, te = 0;
            for (size_t i = 0; i < g_phases.size(); i++) {
                const char ch = g_phases[i];
                if (ch == '{') depth++;
                else if (ch == '}') {
                    if (depth == 2 && !title.empty()) {
                        char b[160];
                        std::snprintf(b, sizeof b, "%s\"%s\": %.1f", top.empty() ? "" : ", ", title.c_str(), te - ts);
                        top += b;
                        title.clear();
                    }
                    depth--;
                } else if (depth == 2 && ch == '"') {
                    const size_t e = g_phases.find('"', i + 1);
                    const std::string key = g_phases.substr(i + 1, e - i - 1);
                    size_t v = g_phases.find(':', e) + 1;
                    while (v < g_phases.size() && g_phases[v] == ' ') v++;
                    if (key == "title") { const size_t e2 = g_phases.find('"', v + 1); title = g_phases.substr(v + 1, e2 - v - 1); i = e2; continue; }
                    if (key == "timeStart") ts = std::atof(g_phases.c_str() + v);
                    if (key == "timeEnd") te = std::atof(g_phases.c_str() + v);
                    i = e;
                }
            }
        }
        std::printf("{\"phases_ms_last_run\": {%s}, ", top.c_str());
        std::printf("\"what\": \"LZSSLCPCompressor<%s, %s>::compress(Input&, Output&), pageable in-memory Input, file Output, %d timed runs after 1 warm-up\", "
                    "\"text_bytes\": %zu, \"archive_bytes\": %zu, \"archive_fnv1a\": \"%016llx\", \"first_run_ms\": %.3f, \"ms_per_step\": %.3f, \"ms_median\": %.3f, \"ms_runs\": [%s]}\n",
                    coder == "huff" ? "HuffmanCoder" : "BitCoder", gpu ? "GpuTextDS" : "TextDS<>", int(times.size() > 1 ? times.size() - 1 : 1),
                    size_t(text.size()), arc.size(), (unsigned long long)fnv1a(arc), times[0], mean, median, runs.c_str());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
