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
// Container: "TDCBLOCK2\n", u64 block_bytes, u64 nblocks, u32 algo_len, algo string; then the raw block archives in the
// order the workers finished them (every worker writes its archive straight into the output file with pwrite at an offset
// reserved from a cursor shared between the processes — nothing is copied a second time); then the index, nblocks x
// (u64 offset, u64 length) in BLOCK order; then u64 index_offset and "TDCBEND\n".  All integers little-endian.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <future>
#include <memory>
#include <new>
#include <streambuf>
#include <string>
#include <thread>
#include <vector>

#include <tudocomp/Compressor.hpp>
#include <tudocomp/io.hpp>
#include <tudocomp_driver/Registry.hpp>
#include <tudocomp_stat/StatPhase.hpp>

#ifdef TDC_GPU_DEFAULT_TEXTDS
#include "tdcgpu.h"  // pinned block buffers: the text goes to the device by DMA straight from where the file was read
#include <tudocomp_gpu/GpuTextDS.hpp>  // gpu_detail::thread_device_override
#endif
#ifdef TDC_BLOCK_THREADS
#include <mutex>
#endif

using namespace tdc;

namespace {

const char MAGIC[] = "TDCBLOCK2\n";
const char MAGIC_END[] = "TDCBEND\n";

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

// ---- host side of a worker -----------------------------------------------------------------------------------------
// A block costs the GPU tens of milliseconds, so the worker's host work decides the throughput (first version: the whole
// file in a vector, the reference's byte-wise escaping pass and byte-wise in-memory Output per block, archive re-read by
// the parent: 1.7 s per 256 MiB block, profiles/r2d_block_mode_2GiB_1gpu.txt).  Now: block b+1 is read (pread) into a
// second buffer while block b is compressed, a block without 0x00 / 0xFF bytes is handed over as a sentinel-terminated
// View (what the restricted Input would have produced, without its copy and byte loops), the archive is collected through
// a streambuf that appends whole chunks, and a writer thread stores it while the next block is compressed.
struct BlockBuffer {
    uint8_t* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    explicit BlockBuffer(size_t n) : cap(n) {
#ifdef TDC_GPU_DEFAULT_TEXTDS
        p = static_cast<uint8_t*>(tdcgpu_pinned_alloc(n));
        pinned = p != nullptr;
#endif
        if (!p) p = static_cast<uint8_t*>(std::malloc(n));
        if (!p) throw std::bad_alloc();
    }
    ~BlockBuffer() {
#ifdef TDC_GPU_DEFAULT_TEXTDS
        if (pinned) { tdcgpu_pinned_free(p); return; }
#endif
        std::free(p);
    }
    BlockBuffer(const BlockBuffer&) = delete;
    BlockBuffer& operator=(const BlockBuffer&) = delete;
};

struct LoadedBlock {
    size_t len = 0;
    bool clean = false;  // no byte that the driver's input restrictions would escape
};

// one contiguous part of a block: pread + scan for the two bytes the input restrictions would escape
bool load_part(int fd, uint64_t from, size_t len, uint8_t* dst) {
    size_t got = 0;
    while (got < len) {
        const ssize_t r = pread(fd, dst + got, len - got, off_t(from + got));
        if (r <= 0) throw std::runtime_error("cannot read the input");
        got += size_t(r);
    }
    return std::memchr(dst, 0x00, len) == nullptr && std::memchr(dst, 0xFF, len) == nullptr;
}

// A 256 MiB block takes one thread ~85 ms to read and scan — more than the GPU needs for it — so four threads share it.
LoadedBlock load_block(int fd, uint64_t from, size_t len, BlockBuffer& buf) {
    constexpr size_t PARTS = 4;
    const size_t part = (len + PARTS - 1) / PARTS;
    std::future<bool> f[PARTS];
    size_t nf = 0;
    for (size_t off = part; off < len; off += part)
        f[nf++] = std::async(std::launch::async, load_part, fd, from + off, std::min(part, len - off), buf.p + off);
    LoadedBlock lb;
    lb.len = len;
    lb.clean = load_part(fd, from, std::min(part, len), buf.p);
    for (size_t i = 0; i < nf; i++) lb.clean = f[i].get() && lb.clean;
    buf.p[len] = 0;  // the sentinel of the clean path (the buffer holds block + 1 bytes)
    return lb;
}

class AppendBuf : public std::streambuf {
    std::vector<uint8_t>& v_;

public:
    explicit AppendBuf(std::vector<uint8_t>& v) : v_(v) {}

protected:
    std::streamsize xsputn(const char* s, std::streamsize n) override {
        v_.insert(v_.end(), reinterpret_cast<const uint8_t*>(s), reinterpret_cast<const uint8_t*>(s) + n);
        return n;
    }
    int_type overflow(int_type c) override {
        if (c != traits_type::eof()) v_.push_back(uint8_t(c));
        return c;
    }
};

// one block through the registry, exactly as the driver runs a whole file
// `arc` is cleared and refilled: the caller reuses two such vectors for all its blocks, so after the first blocks the archive
// lands in memory that is already allocated and faulted in (a fresh vector per block cost 230 ms per 160 MB archive in
// reallocation and page faults — four times the GPU work of the block).
void compress_block(const std::string& algo, const uint8_t* data, size_t len, bool sentinel_follows, std::vector<uint8_t>& arc,
                    std::string* stats_json = nullptr) {
    auto& registry = tdc_algorithms::COMPRESSOR_REGISTRY;
#ifdef TDC_BLOCK_THREADS
    static std::mutex registry_mu;  // the registry is only read, but nothing in the reference promises that is thread-safe
    std::unique_lock<std::mutex> registry_lock(registry_mu);
#endif
    auto av = registry.parse_algorithm_id(algo);
    auto restrictions = av.textds_flags();
    auto compressor = registry.select_algorithm(av);
#ifdef TDC_BLOCK_THREADS
    registry_lock.unlock();
#endif
    arc.clear();
    if (arc.capacity() < len / 2 + 4096) arc.reserve(len / 2 + 4096);
    {
        StatPhase root("root");
        AppendBuf sb(arc);
        std::ostream os(&sb);
        Output out = Output::from_stream(os);
        // restrictions {escape 0, append sentinel} on a block without 0x00 / 0xFF bytes = the same bytes + one 0
        // (io/EscapeMap.hpp:39-64, io/RestrictedBuffer.hpp:108-140): hand that View over directly
        const bool direct = sentinel_follows && restrictions.has_restrictions() && restrictions.null_terminate() &&
                            restrictions.escape_bytes() == std::vector<uint8_t>{0};
        if (direct) {
            Input inp(View(data, len + 1));
            compressor->compress(inp, out);
        } else {
            Input inp(View(data, len));
            if (restrictions.has_restrictions()) inp = Input(inp, restrictions);
            compressor->compress(inp, out);
        }
        os.flush();
#ifndef STATS_DISABLED
        if (stats_json) *stats_json = root.to_json().str();
#endif
    }
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

// state shared between the forked workers (anonymous shared mapping): where the next archive goes, and the index
struct Shared {
    std::atomic<uint64_t> cursor;
    uint64_t* entry(uint64_t b) { return reinterpret_cast<uint64_t*>(this + 1) + 2 * b; }  // {offset, length}
};

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
            // trailer: index offset + end mark; the index holds (offset, length) per block in block order
            const size_t tail = 8 + sizeof(MAGIC_END) - 1;
            if (size_t(end - p) < tail || std::memcmp(end - (sizeof(MAGIC_END) - 1), MAGIC_END, sizeof(MAGIC_END) - 1) != 0) throw std::runtime_error("truncated container");
            const uint8_t* q = end - tail;
            const uint64_t index_off = get_u64(q, end);
            if (index_off > c.size() - tail || (c.size() - tail - index_off) / 16 < nblocks) throw std::runtime_error("corrupt container index");
            const uint8_t* idx = c.data() + index_off;
            std::ofstream out(ofile, std::ios::binary | std::ios::trunc);
            for (uint64_t b = 0; b < nblocks; b++) {
                const uint64_t off = get_u64(idx, end), len = get_u64(idx, end);
                if (off > index_off || len > index_off - off) throw std::runtime_error("corrupt container index");
                const std::vector<uint8_t> text = decompress_block(algo.empty() ? stored : algo, c.data() + off, len);
                out.write(reinterpret_cast<const char*>(text.data()), std::streamsize(text.size()));
            }
            out.close();
            if (!out) throw std::runtime_error("cannot write " + ofile);
            return 0;
        }
        const int fd = open(input.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + input);
        struct stat sb;
        if (fstat(fd, &sb) != 0) throw std::runtime_error("cannot stat " + input);
        const uint64_t in_size = uint64_t(sb.st_size);
        const uint64_t nblocks = (in_size + block - 1) / block;
        workers = int(std::min<uint64_t>(uint64_t(workers), nblocks ? nblocks : 1));
        // the output file with its header, and the state shared with the workers: write cursor + index
        const int out_fd = open(ofile.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (out_fd < 0) throw std::runtime_error("cannot create " + ofile);
        std::vector<uint8_t> head(MAGIC, MAGIC + sizeof(MAGIC) - 1);
        auto head_u64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) head.push_back(uint8_t(v >> (8 * i))); };
        head_u64(block);
        head_u64(nblocks);
        for (int i = 0; i < 4; i++) head.push_back(uint8_t(uint32_t(algo.size()) >> (8 * i)));
        head.insert(head.end(), algo.begin(), algo.end());
        if (pwrite(out_fd, head.data(), head.size(), 0) != ssize_t(head.size())) throw std::runtime_error("cannot write " + ofile);
        const size_t shared_bytes = sizeof(Shared) + size_t(nblocks) * 16;
        void* shm = mmap(nullptr, shared_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (shm == MAP_FAILED) throw std::runtime_error("mmap failed");
        Shared* shared = new (shm) Shared();
        shared->cursor.store(head.size());
        for (uint64_t bb = 0; bb < nblocks; bb++) { shared->entry(bb)[0] = 0; shared->entry(bb)[1] = ~uint64_t(0); }
        auto run_worker = [&](int k) {
            const auto tw0 = std::chrono::steady_clock::now();
#if defined(TDC_GPU_DEFAULT_TEXTDS) && defined(TDC_BLOCK_THREADS)
            // worker THREAD k drives device k (first device of the process + k): its contexts, pinned buffers and copies
            {
                const int dev = (std::getenv("TDCGPU_DEVICE") ? std::atoi(std::getenv("TDCGPU_DEVICE")) : 0) + k;
                gpu_detail::thread_device_override() = dev;
                tdcgpu_set_device(dev);
            }
#elif defined(TDC_GPU_DEFAULT_TEXTDS)
            {
                // Worker PROCESS k sees ONLY its own GPU (CUDA_VISIBLE_DEVICES, set before the first CUDA call of this process):
                // the runtime then initialises one device instead of all of the box.  An existing CUDA_VISIBLE_DEVICES
                // list is honoured: its k-th entry.
                int dev = workers > 1 ? k : (std::getenv("TDCGPU_DEVICE") ? std::atoi(std::getenv("TDCGPU_DEVICE")) : 0);
                std::string pick = std::to_string(dev);
                if (const char* cvd = std::getenv("CUDA_VISIBLE_DEVICES")) {
                    std::vector<std::string> ids;
                    std::string cur;
                    for (const char* p = cvd;; p++) {
                        if (*p == ',' || *p == 0) { if (!cur.empty()) ids.push_back(cur); cur.clear(); if (!*p) break; }
                        else cur.push_back(*p);
                    }
                    if (!ids.empty()) pick = ids[size_t(dev) % ids.size()];
                }
                setenv("CUDA_VISIBLE_DEVICES", pick.c_str(), 1);
                setenv("TDCGPU_DEVICE", "0", 1);
                tdcgpu_set_device(0);
            }
#else
            if (workers > 1) setenv("TDCGPU_DEVICE", std::to_string(k).c_str(), 1);
#endif
            const size_t cap = size_t(std::min<uint64_t>(block, in_size)) + 1;
            std::unique_ptr<BlockBuffer> bufs[2] = {std::make_unique<BlockBuffer>(cap), std::make_unique<BlockBuffer>(cap)};
            auto span = [&](uint64_t b) { return std::make_pair(b * block, size_t(std::min<uint64_t>(block, in_size - b * block))); };
            if (std::getenv("TDC_BLOCK_VERBOSE"))
                std::cerr << "[worker " << k << "] block buffers ready after " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count() << " ms\n";
            std::vector<uint8_t> arcs[2];  // archives of the current and the previous block (the latter being written)
            std::future<LoadedBlock> next;
            std::future<void> writing;
            int cur = 0;
            if (uint64_t(k) < nblocks) next = std::async(std::launch::async, load_block, fd, span(k).first, span(k).second, std::ref(*bufs[0]));
            for (uint64_t b = uint64_t(k); b < nblocks; b += uint64_t(workers)) {
                const auto tb0 = std::chrono::steady_clock::now();
                const LoadedBlock lb = next.get();
                const auto tb1 = std::chrono::steady_clock::now();
                const uint64_t nb = b + uint64_t(workers);
                if (nb < nblocks) next = std::async(std::launch::async, load_block, fd, span(nb).first, span(nb).second, std::ref(*bufs[cur ^ 1]));
                std::string stats;
                const bool verbose = std::getenv("TDC_BLOCK_VERBOSE") != nullptr;
                std::vector<uint8_t>* arc = &arcs[cur];  // its previous content (block b - 2 * workers) has been written out
                compress_block(algo, bufs[cur]->p, lb.len, lb.clean, *arc, verbose ? &stats : nullptr);
                const auto tb2 = std::chrono::steady_clock::now();
                if (writing.valid()) writing.get();
                if (std::getenv("TDC_BLOCK_VERBOSE")) {
                    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point z) { return std::chrono::duration<double, std::milli>(z - a).count(); };
                    std::cerr << "[worker " << k << "] block " << b << ": waited for the read " << ms(tb0, tb1) << " ms, compress " << ms(tb1, tb2)
                              << " ms, waited for the previous write " << ms(tb2, std::chrono::steady_clock::now()) << " ms, " << (lb.clean ? "direct view" : "restricted input")
                              << ", " << arc->size() << " bytes out\n";
                    if (std::getenv("TDC_BLOCK_VERBOSE")[0] == '2') std::cerr << stats << "\n";
                }
                writing = std::async(std::launch::async, [arc, b, shared, out_fd] {
                    const uint64_t len = arc->size();
                    const uint64_t off = shared->cursor.fetch_add(len);  // reserve the archive's place in the container
                    for (uint64_t done = 0; done < len;) {
                        const ssize_t w = pwrite(out_fd, arc->data() + done, size_t(std::min<uint64_t>(len - done, uint64_t(1) << 30)), off_t(off + done));
                        if (w <= 0) throw std::runtime_error("cannot write the container");
                        done += uint64_t(w);
                    }
                    shared->entry(b)[0] = off;
                    shared->entry(b)[1] = len;
                });
                cur ^= 1;
            }
            if (writing.valid()) writing.get();
        };
        bool failed = false;
#ifdef TDC_BLOCK_THREADS
        // One worker THREAD per GPU in this process (the registry units of this build are compiled with -DSTATS_DISABLED, the
        // reference's own switch, because StatPhase keeps an unsynchronised global): CUDA is initialised once.  With one
        // PROCESS per GPU the eight concurrent CUDA start-ups of an 8-GPU box took 10.4 s before the first block
        // (profiles/r2_summary.md §5).
        {
            std::vector<std::thread> th;
            std::vector<int> rcs(size_t(workers), 0);
            for (int k = 0; k < workers; k++)
                th.emplace_back([&, k] {
                    try {
                        run_worker(k);
                    } catch (const std::exception& e) {
                        std::cerr << "Error (worker " << k << "): " << e.what() << std::endl;
                        rcs[size_t(k)] = 1;
                    }
                });
            for (auto& t : th) t.join();
            for (int rc : rcs) failed = failed || rc != 0;
        }
#else
        // Workers are forked before any CUDA call: every worker creates its own context on its own device.
        std::vector<pid_t> pids;
        for (int k = 0; k < workers; k++) {
            const pid_t pid = fork();
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
        for (pid_t pid : pids) {
            int st = 0;
            waitpid(pid, &st, 0);
            if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) failed = true;
        }
#endif
        for (uint64_t bb = 0; bb < nblocks && !failed; bb++)
            if (shared->entry(bb)[1] == ~uint64_t(0)) failed = true;  // a block nobody finished
        if (failed) {
            close(out_fd);
            std::remove(ofile.c_str());
            throw std::runtime_error("a worker failed");
        }
        // index + trailer behind the last archive
        const uint64_t index_off = shared->cursor.load();
        std::vector<uint8_t> trailer;
        auto push_u64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) trailer.push_back(uint8_t(v >> (8 * i))); };
        uint64_t total = 0;
        for (uint64_t bb = 0; bb < nblocks; bb++) {
            push_u64(shared->entry(bb)[0]);
            push_u64(shared->entry(bb)[1]);
            total += shared->entry(bb)[1];
        }
        push_u64(index_off);
        trailer.insert(trailer.end(), MAGIC_END, MAGIC_END + sizeof(MAGIC_END) - 1);
        if (pwrite(out_fd, trailer.data(), trailer.size(), off_t(index_off)) != ssize_t(trailer.size()) || close(out_fd) != 0)
            throw std::runtime_error("cannot write " + ofile);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cerr << "tdc_block: " << in_size << " bytes in " << nblocks << " block(s) of " << block << " on " << workers
                  << " worker(s) -> " << total << " bytes, " << secs << " s (" << (secs > 0 ? in_size / 1e6 / secs : 0.0) << " MB/s)\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
