// tdc_block — block mode (BASELINE config 5: "block-parallel lzss_lcp with Huffman coding, 256 MB independent blocks across
// 8 B200") as a small driver next to the reference's `tdc`.  The reference has no block mode; per block the semantics
// are exactly those of `tdc -a <algo> --raw` on the block's bytes: the same registry (tdc_algorithms::COMPRESSOR_REGISTRY,
// include/tudocomp_driver/Registry.hpp:44-50), the same input restrictions (escaping + sentinel,
// src/tudocomp_driver/tudocomp_driver.cpp:268-270) and the same Compressor::compress / decompress calls (:275, :340).
// With the GPU registry (`tdc_block_gpu`) every block runs on a GPU; blocks are dealt round-robin to `-g N` worker
// processes, worker k using device k (TDCGPU_DEVICE), with no communication between them.
//
//   tdc_block -a "lzss_lcp(coder=huff)" -b 268435456 -g 8 [-c] input -o output.tdcb
//        -c  keep one device context per worker alive across its blocks (TDCGPU_CTX_CACHE=1, GpuTextDS.hpp)
//   tdc_block -d output.tdcb -o roundtrip
//
// Container: "TDCBLOCK1\n", u64 block_bytes, u64 nblocks, u32 algo_len, algo string, then per block u64 archive_len and
// the raw archive.  All integers little-endian.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <tudocomp/Compressor.hpp>
#include <tudocomp/io.hpp>
#include <tudocomp_driver/Registry.hpp>
#include <tudocomp_stat/StatPhase.hpp>

using namespace tdc;

namespace {

const char MAGIC[] = "TDCBLOCK1\n";

std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const std::streamsize n = f.tellg();
    std::vector<uint8_t> buf(static_cast<size_t>(n));
    f.seekg(0);
    if (n && !f.read(reinterpret_cast<char*>(buf.data()), n)) throw std::runtime_error("cannot read " + path);
    return buf;
}

void put_u64(std::ostream& o, uint64_t v) { o.write(reinterpret_cast<const char*>(&v), 8); }
void put_u32(std::ostream& o, uint32_t v) { o.write(reinterpret_cast<const char*>(&v), 4); }
uint64_t get_u64(const uint8_t*& p, const uint8_t* end) {
    if (end - p < 8) throw std::runtime_error("truncated container");
    uint64_t v;
    std::memcpy(&v, p, 8);
    p += 8;
    return v;
}

// one block through the registry, exactly as the driver runs a whole file
std::vector<uint8_t> compress_block(const std::string& algo, const uint8_t* data, size_t len) {
    auto& registry = tdc_algorithms::COMPRESSOR_REGISTRY;
    auto av = registry.parse_algorithm_id(algo);
    auto restrictions = av.textds_flags();
    auto compressor = registry.select_algorithm(av);
    std::vector<uint8_t> arc;
    {
        StatPhase root("root");
        Input inp(View(data, len));
        if (restrictions.has_restrictions()) inp = Input(inp, restrictions);
        Output out = Output::from_memory(arc);
        compressor->compress(inp, out);
    }
    return arc;
}

std::vector<uint8_t> decompress_block(const std::string& algo, const uint8_t* arc, size_t len) {
    auto& registry = tdc_algorithms::COMPRESSOR_REGISTRY;
    auto av = registry.parse_algorithm_id(algo);
    auto restrictions = av.textds_flags();
    auto compressor = registry.select_algorithm(av);
    std::vector<uint8_t> text;
    {
        StatPhase root("root");
        Input inp(View(arc, len));
        Output out = Output::from_memory(text);
        if (restrictions.has_restrictions()) out = Output(out, restrictions);
        compressor->decompress(inp, out);
    }
    return text;
}

std::string block_tmp(const std::string& ofile, uint64_t b) { return ofile + ".blk." + std::to_string(b); }

int usage() {
    std::cerr << "usage: tdc_block -a ALGO [-b BLOCK_BYTES] [-g WORKERS] [-c] INPUT -o OUTPUT\n"
                 "       tdc_block -d CONTAINER -o OUTPUT\n";
    return 2;
}

}  // namespace

int main(int argc, char** argv) {
    std::string algo, input, ofile;
    uint64_t block = uint64_t(256) << 20;
    int workers = 1;
    bool decompress = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-a" && i + 1 < argc) algo = argv[++i];
        else if (a == "-b" && i + 1 < argc) block = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "-g" && i + 1 < argc) workers = std::atoi(argv[++i]);
        else if (a == "-o" && i + 1 < argc) ofile = argv[++i];
        else if (a == "-d") decompress = true;
        else if (a == "-c") setenv("TDCGPU_CTX_CACHE", "1", 1);
        else if (!a.empty() && a[0] != '-') input = a;
        else return usage();
    }
    if (input.empty() || ofile.empty() || (!decompress && algo.empty()) || block == 0 || workers < 1) return usage();
    try {
        const auto t0 = std::chrono::steady_clock::now();
        if (decompress) {
            const std::vector<uint8_t> c = read_file(input);
            const uint8_t *p = c.data(), *end = c.data() + c.size();
            if (c.size() < sizeof(MAGIC) - 1 || std::memcmp(p, MAGIC, sizeof(MAGIC) - 1) != 0) throw std::runtime_error("not a tdc_block container");
            p += sizeof(MAGIC) - 1;
            get_u64(p, end);  // block size (informational)
            const uint64_t nblocks = get_u64(p, end);
            if (end - p < 4) throw std::runtime_error("truncated container");
            uint32_t alen;
            std::memcpy(&alen, p, 4);
            p += 4;
            if (uint64_t(end - p) < alen) throw std::runtime_error("truncated container");
            const std::string stored(reinterpret_cast<const char*>(p), alen);
            p += alen;
            std::ofstream out(ofile, std::ios::binary | std::ios::trunc);
            for (uint64_t b = 0; b < nblocks; b++) {
                const uint64_t len = get_u64(p, end);
                if (uint64_t(end - p) < len) throw std::runtime_error("truncated container");
                const std::vector<uint8_t> text = decompress_block(algo.empty() ? stored : algo, p, len);
                out.write(reinterpret_cast<const char*>(text.data()), std::streamsize(text.size()));
                p += len;
            }
            return 0;
        }
        const std::vector<uint8_t> data = read_file(input);
        const uint64_t nblocks = (data.size() + block - 1) / block;
        workers = int(std::min<uint64_t>(uint64_t(workers), nblocks ? nblocks : 1));
        auto run_worker = [&](int k) {
            if (workers > 1) setenv("TDCGPU_DEVICE", std::to_string(k).c_str(), 1);  // device k for worker k (GPU registry)
            for (uint64_t b = uint64_t(k); b < nblocks; b += uint64_t(workers)) {
                const uint64_t from = b * block, len = std::min<uint64_t>(block, data.size() - from);
                const std::vector<uint8_t> arc = compress_block(algo, data.data() + from, len);
                std::ofstream t(block_tmp(ofile, b), std::ios::binary | std::ios::trunc);
                t.write(reinterpret_cast<const char*>(arc.data()), std::streamsize(arc.size()));
                if (!t) throw std::runtime_error("cannot write " + block_tmp(ofile, b));
            }
        };
        if (workers == 1) {
            run_worker(0);
        } else {
            std::vector<pid_t> pids;
            for (int k = 0; k < workers; k++) {
                const pid_t pid = fork();  // before any CUDA call: every worker creates its own context on its own device
                if (pid < 0) throw std::runtime_error("fork failed");
                if (pid == 0) {
                    int rc = 0;
                    try {
                        run_worker(k);
                    } catch (const std::exception& e) {
                        std::cerr << "Error (worker " << k << "): " << e.what() << std::endl;
                        rc = 1;
                    }
                    _exit(rc);
                }
                pids.push_back(pid);
            }
            bool ok = true;
            for (pid_t pid : pids) {
                int st = 0;
                waitpid(pid, &st, 0);
                ok = ok && WIFEXITED(st) && WEXITSTATUS(st) == 0;
            }
            if (!ok) throw std::runtime_error("a worker failed");
        }
        // assemble the container in block order
        std::ofstream out(ofile, std::ios::binary | std::ios::trunc);
        out.write(MAGIC, sizeof(MAGIC) - 1);
        put_u64(out, block);
        put_u64(out, nblocks);
        put_u32(out, uint32_t(algo.size()));
        out.write(algo.data(), std::streamsize(algo.size()));
        uint64_t total = 0;
        for (uint64_t b = 0; b < nblocks; b++) {
            const std::vector<uint8_t> arc = read_file(block_tmp(ofile, b));
            put_u64(out, arc.size());
            out.write(reinterpret_cast<const char*>(arc.data()), std::streamsize(arc.size()));
            total += arc.size();
            std::remove(block_tmp(ofile, b).c_str());
        }
        if (!out) throw std::runtime_error("cannot write " + ofile);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cerr << "tdc_block: " << data.size() << " bytes in " << nblocks << " block(s) of " << block << " on " << workers
                  << " worker(s) -> " << total << " bytes, " << secs << " s (" << (secs > 0 ? data.size() / 1e6 / secs : 0.0) << " MB/s)\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
