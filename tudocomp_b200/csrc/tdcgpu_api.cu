// C ABI of libtdcgpu (see include/tdcgpu.h for the contract and the reference interfaces each entry replaces).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>

#include "../../include/tdcgpu.h"
#include "tdc_ctx.h"

namespace tdc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- launch accounting + optional per-kernel event timing -------------------------------------------------------
struct ProfEntry {
    std::string name;
    u64 launches = 0;
    double ms = 0, bytes = 0;
};
struct PendingLaunch {
    int entry;
    cudaEvent_t a, b;
};
static std::atomic<u64> g_launches{0};
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;
static std::vector<PendingLaunch> g_pending;
static std::vector<cudaEvent_t> g_event_pool;

static int prof_entry(const char* name) {
    for (size_t i = 0; i < g_prof.size(); i++)
        if (g_prof[i].name == name) return int(i);
    ProfEntry e;
    e.name = name;
    g_prof.push_back(e);
    return int(g_prof.size()) - 1;
}
static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
#ifndef TDC_CUSIM
LaunchScope::LaunchScope(const char* name, cudaStream_t stream, bool is_kernel) : slot(-1), st(stream) {
    if (is_kernel) g_launches++;
    if (!g_prof_on) return;
    PendingLaunch p;
    p.entry = prof_entry(name);
    p.a = take_event();
    p.b = take_event();
    cudaEventRecord(p.a, st);
    g_pending.push_back(p);
    slot = int(g_pending.size()) - 1;
}
LaunchScope::~LaunchScope() {
    if (slot >= 0) cudaEventRecord(g_pending[slot].b, st);
}
void prof_add_bytes(const char* name, double bytes) {
    if (!g_prof_on) return;
    g_prof[prof_entry(name)].bytes += bytes;
}
#endif
static void prof_resolve() {  // call after the stream has been synchronised
    for (auto& p : g_pending) {
        float ms = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof[p.entry].ms += ms;
            g_prof[p.entry].launches++;
        }
        g_event_pool.push_back(p.a);
        g_event_pool.push_back(p.b);
    }
    g_pending.clear();
}

PhaseTimer::PhaseTimer(Ctx& c_, const char* name_) : c(c_), name(name_), a(nullptr), b(nullptr) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, c.stream);
}
PhaseTimer::~PhaseTimer() {
    cudaEventRecord(b, c.stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    c.phases.push_back(PhaseTime{name, ms});
    cudaEventDestroy(a);
    cudaEventDestroy(b);
}

// One context computes at a time per device.  Several contexts on one GPU (a host thread each) are how a stream of
// independent texts is double-buffered: while one context runs its kernels, another one copies its next text in or its
// last result out.  Letting their kernels run concurrently only makes every kernel slower (each of them fills the GPU, and
// the sort and scatter kernels rely on in-order CTA dispatch for L2 locality), so the compute calls take this lock and the
// copy calls do not.  TDCGPU_NO_COMPUTE_LOCK=1 removes it (A/B measurements).
static std::mutex g_compute_mu[64];
struct ComputeLock {
    std::unique_lock<std::mutex> lk;
    explicit ComputeLock(int device) {
        static const bool off = [] { const char* e = std::getenv("TDCGPU_NO_COMPUTE_LOCK"); return e && *e && *e != '0'; }();
        if (!off) lk = std::unique_lock<std::mutex>(g_compute_mu[device & 63]);
    }
};

static const u64 TEXT_PAD = 1024;  // zero bytes readable past the text (8-byte LCE loads, warp-wide look-ahead)

static void free_arrays(Ctx& c) {
    void* ps[] = {c.d_text, c.d_sa, c.d_isa, c.d_lcp, c.d_phi, c.d_plcp, c.d_bwt, c.arena.base};
    for (void* p : ps)
        if (p) cudaFree(p);
    c.d_text = nullptr;
    c.d_sa = c.d_isa = c.d_lcp = c.d_phi = c.d_plcp = nullptr;
    c.d_bwt = nullptr;
    c.arena = Arena();
    c.cap_n = 0;
    c.have = 0;
}

static int ensure_capacity(Ctx& c, u64 n) {
    if (n <= c.cap_n) return 0;
    free_arrays(c);
    TDC_CUDA(cudaMalloc(&c.d_text, n + TEXT_PAD + 16));
    TDC_CUDA(cudaMalloc(&c.d_sa, sizeof(u32) * n));
    TDC_CUDA(cudaMalloc(&c.d_isa, sizeof(u32) * n));
    // scratch: SA construction is the high-water mark: 2 x u64 keys, 2 x u32 values, 2 x u32 slots, u32 group ids,
    // 2 x u32 scatter buffers
    const size_t arena_bytes = size_t(44) * n + n / 8 + (size_t(4) << 20);
    TDC_CUDA(cudaMalloc(&c.arena.base, arena_bytes));
    c.arena.cap = arena_bytes;
    c.arena.off = 0;
    TDC_TRY(sort_workspace_init(c.sortws, n, c.sm_count));
    c.cap_n = n;
    return 0;
}

template <class T>
static int lazy_alloc(T** p, u64 count) {
    if (*p) return 0;
    TDC_CUDA(cudaMalloc(p, sizeof(T) * count));
    return 0;
}

static int do_build(Ctx& c, u32 flags) {
    if (c.n == 0) { set_error("no text loaded"); return TDCGPU_ERR_STATE; }
    u32 need = flags & ~c.have;  // dependencies below are only pulled in for structures that are actually missing
    if (!need) return 0;
    // LCP alone on a text whose suffixes separate early (estimated mean LCP small): compare characters directly in SA
    // order.  Otherwise (or when Phi/PLCP are wanted anyway) follow the reference's Phi -> PLCP -> LCP data flow.
    bool lcp_direct = false;
    if ((need & DS_LCP) && !(c.have & DS_LCP) && !((need | c.have) & (DS_PHI | DS_PLCP))) {
        if (!(c.have & DS_SA)) {
            PhaseTimer t(c, "Construct SA");
            TDC_TRY(lazy_alloc(&c.d_lcp, c.cap_n));  // the initial sort seeds it with the LCPs its keys decide
            TDC_TRY(build_suffix_array(c, true));
            c.have |= DS_SA | DS_ISA;
        }
        lcp_direct = c.sa_prefix_work / double(c.n) <= 48.0;
    }
    if (lcp_direct) {
        PhaseTimer t(c, "Construct LCP Array");
        TDC_TRY(lazy_alloc(&c.d_lcp, c.cap_n));
        TDC_TRY(build_lcp_direct(c));
        c.have |= DS_LCP;
        c.lcp_route = 1;
        need &= ~DS_LCP;
    }
    if (need & DS_LCP) need |= DS_PLCP | DS_SA;
    if (need & DS_PLCP) need |= DS_PHI | DS_SA;
    if (need & (DS_PHI | DS_ISA | DS_BWT)) need |= DS_SA;
    need |= (need & DS_SA) ? DS_ISA : 0;  // the doubling builder always produces both
    need &= ~c.have;
    if (!need) return 0;
    if (need & DS_SA) {
        PhaseTimer t(c, "Construct SA");  // also yields ISA ("Construct ISA" is free on this path)
        TDC_TRY(build_suffix_array(c, false));
        c.have |= DS_SA | DS_ISA;
    }
    if (need & (DS_PHI | DS_BWT)) {
        PhaseTimer t(c, (need & DS_PHI) ? "Construct Phi Array" : "Construct BWT");
        if (need & DS_PHI) TDC_TRY(lazy_alloc(&c.d_phi, c.cap_n));
        if (need & DS_BWT) TDC_TRY(lazy_alloc(&c.d_bwt, c.cap_n));
        TDC_TRY(build_phi_bwt(c, (need & DS_PHI) != 0, (need & DS_BWT) != 0));
        c.have |= need & (DS_PHI | DS_BWT);
    }
    if (need & DS_PLCP) {
        PhaseTimer t(c, (need & DS_LCP) ? "Construct PLCP+LCP Array" : "Construct PLCP Array");
        TDC_TRY(lazy_alloc(&c.d_plcp, c.cap_n));
        if (need & DS_LCP) TDC_TRY(lazy_alloc(&c.d_lcp, c.cap_n));
        TDC_TRY(build_plcp_lcp(c, (need & DS_LCP) != 0));
        c.have |= need & (DS_PLCP | DS_LCP);
        if (need & DS_LCP) c.lcp_route = 2;
    }
    return 0;
}

// One thread per 64-bit output word: gathers the <= 64/width + 2 elements that overlap it.  Reads of neighbouring
// threads overlap and stay in L1/L2, writes are coalesced: 4 B read + width/8 B written per element.
static __global__ void __launch_bounds__(256)
pack_bits_kernel(const u32* __restrict__ a, u64 n, u32 w, u64* __restrict__ out, u64 nwords) {
    const u64 j = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= nwords) return;
    const u64 lo = j * 64, hi = lo + 64;
    const u32 mask = w >= 32 ? 0xffffffffu : ((1u << w) - 1u);
    u64 acc = 0;
    for (u64 e = lo / w; e < n && e * w < hi; e++) {
        const u64 v = u64(a[e] & mask), start = e * w;
        acc |= start >= lo ? (v << (start - lo)) : (v >> (lo - start));
    }
    out[j] = acc;
}

// do two device buffers of n bytes differ?  (both readable up to the next multiple of 16)
static __global__ void __launch_bounds__(256) bytes_differ_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, u64 n, u32* __restrict__ flag) {
    const u64 nvec = n / 16;
    bool diff = false;
    for (u64 v = u64(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += u64(gridDim.x) * blockDim.x) {
        const uint4 x = reinterpret_cast<const uint4*>(a)[v], y = reinterpret_cast<const uint4*>(b)[v];
        diff |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 15u)) diff |= a[nvec * 16 + threadIdx.x] != b[nvec * 16 + threadIdx.x];
    if (diff) *flag = 1;
}

static void* array_ptr(Ctx& c, u32 which, size_t* elem) {
    *elem = 4;
    switch (which) {
        case DS_SA: return c.d_sa;
        case DS_ISA: return c.d_isa;
        case DS_LCP: return c.d_lcp;
        case DS_PHI: return c.d_phi;
        case DS_PLCP: return c.d_plcp;
        case DS_BWT: *elem = 1; return c.d_bwt;
    }
    return nullptr;
}

}  // namespace tdc

using namespace tdc;

#define API_GUARD(ctx)                                         \
    if (!(ctx)) { set_error("null context"); return TDCGPU_ERR_ARG; } \
    Ctx& c = (ctx)->c;                                         \
    if (cudaSetDevice(c.device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", c.device); return TDCGPU_ERR_CUDA; }

extern "C" {

const char* tdcgpu_last_error(void) { return g_err; }

int tdcgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int tdcgpu_create(int device, tdcgpu_ctx** out) {
    if (!out) { set_error("null out pointer"); return TDCGPU_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device available (tdcgpu has no CPU fallback)");
        return TDCGPU_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return TDCGPU_ERR_ARG; }
    TDC_CUDA(cudaSetDevice(device));
    tdcgpu_ctx* h = new (std::nothrow) tdcgpu_ctx();
    if (!h) { set_error("out of host memory"); return TDCGPU_ERR_NOMEM; }
    Ctx& c = h->c;
    c.device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c.sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c.d_scalars, 512 * sizeof(u32)) != cudaSuccess ||
        cudaMallocHost(&c.h_scalars, 512 * sizeof(u32)) != cudaSuccess) {
        set_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return TDCGPU_ERR_CUDA;
    }
    *out = h;
    return 0;
}

void tdcgpu_destroy(tdcgpu_ctx* ctx) {
    if (!ctx) return;
    Ctx& c = ctx->c;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    free_arrays(c);
    sort_workspace_free(c.sortws);
    if (c.d_factors) cudaFree(c.d_factors);
    if (c.stream_arena.base) cudaFree(c.stream_arena.base);
    if (c.lit.d_out) cudaFree(c.lit.d_out);
    if (c.d_scalars) cudaFree(c.d_scalars);
    if (c.h_scalars) cudaFreeHost(c.h_scalars);
    for (auto& e : c.user_events)
        if (e) cudaEventDestroy(e);
    host_copier_free(c.copier);
    cudaStreamDestroy(c.stream);
    delete ctx;
}

int tdcgpu_set_text(tdcgpu_ctx* ctx, const uint8_t* text, uint64_t n, int on_device) {
    API_GUARD(ctx);
    if (!text || n == 0) { set_error("empty text (the path always sees at least the sentinel)"); return TDCGPU_ERR_ARG; }
    // indices are UNSIGNED 32-bit: the reference's default build stops at 2^31 (divsufsort's sign bit); here a text may have
    // up to 2^32 - 2^20 bytes (the margin keeps tile-rounded indices below 2^32).  Beyond 2^31 two restrictions apply:
    // common prefixes must stay below 2^31 (the (len << 1 | side) words of the factoriser and a flag bit of the PLCP
    // pass) — checked by tdcgpu_lzss_lcp_factorize / the LCP build through max_lcp — and the archive needs the
    // reference's wide-index format (tdcgpu_set_len_bits(ctx, 64)) to be decodable by a -DLEN_BITS=40 build.
    if (n > (uint64_t(1) << 32) - (uint64_t(1) << 20)) { set_error("n = %llu: indices are 32-bit, n must be <= 2^32 - 2^20", (unsigned long long)n); return TDCGPU_ERR_ARG; }
    if (!on_device && text[n - 1] != 0) {  // device input: checked by the builder (suffix_array.cu), which reads d_text[n-1]
        set_error("Input has no sentinel! (the last text byte must be 0, ds/TextDS.hpp:132-138)");
        return TDCGPU_ERR_SENTINEL;
    }
    TDC_TRY(ensure_capacity(c, n));
    c.n = n;
    c.have = 0;
    c.max_lcp = 0;
    c.num_factors = 0;
    c.have_factors = false;
    c.enc.prepared = c.enc.encoded = false;
    c.phases.clear();
    TDC_CUDA(cudaMemsetAsync(c.d_text + n, 0, TEXT_PAD + 16, c.stream));
    if (on_device) {
        TDC_CUDA(cudaMemcpyAsync(c.d_text, text, n, cudaMemcpyDeviceToDevice, c.stream));
        TDC_CUDA(cudaStreamSynchronize(c.stream));  // the source may be overwritten once the call returns
    } else {
        TDC_TRY(host_copy(c, c.d_text, text, n, true));  // blocking: the caller may reuse its buffer
    }
    return 0;
}

int tdcgpu_set_text_cached(tdcgpu_ctx* ctx, const uint8_t* text, uint64_t n, int* reused) {
    if (reused) *reused = 0;
    {
        API_GUARD(ctx);
        if (text && n > 0 && c.n == n && c.d_text && c.arena.cap >= n + 64) {
            // the same length as the resident text: upload next to it and compare on the device (exact, not a hash)
            c.arena.reset();
            uint8_t* tmp = c.arena.take<uint8_t>(n + 16);
            u32* flag = c.d_scalars + 3;
            if (tmp) {
                TDC_CUDA(cudaMemsetAsync(flag, 0, sizeof(u32), c.stream));
                TDC_TRY(host_copy(c, tmp, text, n, true));
                TDC_LAUNCH(bytes_differ_kernel, u32(std::min<u64>(u64(c.sm_count) * 8, div_up(n / 16 + 1, 256))), 256, 0, c.stream, tmp, c.d_text, n, flag);
                TDC_KCHECK();
                u32 h = 1;
                TDC_CUDA(cudaMemcpyAsync(&h, flag, sizeof(u32), cudaMemcpyDeviceToHost, c.stream));
                TDC_CUDA(cudaStreamSynchronize(c.stream));
                if (h == 0) {
                    if (reused) *reused = 1;
                    return 0;
                }
            }
        }
    }
    return tdcgpu_set_text(ctx, text, n, 0);
}

int tdcgpu_textds_build(tdcgpu_ctx* ctx, uint32_t flags) {
    API_GUARD(ctx);
    ComputeLock lock(c.device);
    c.phases.clear();
    return do_build(c, flags);
}

int tdcgpu_textds_get(tdcgpu_ctx* ctx, uint32_t which, void* dst, int to_device) {
    API_GUARD(ctx);
    size_t elem;
    void* src = array_ptr(c, which, &elem);
    if (!dst) { set_error("null destination"); return TDCGPU_ERR_ARG; }
    if (!src || !(c.have & which)) { set_error("structure 0x%x has not been built", which); return TDCGPU_ERR_STATE; }
    if (!to_device) return host_copy(c, dst, src, elem * c.n, false);
    TDC_CUDA(cudaMemcpyAsync(dst, src, elem * c.n, cudaMemcpyDeviceToDevice, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_textds_get_packed(tdcgpu_ctx* ctx, uint32_t which, uint32_t width, uint64_t* dst, uint64_t cap_words, int to_device) {
    API_GUARD(ctx);
    size_t elem;
    void* src = array_ptr(c, which, &elem);
    if (!dst) { set_error("null destination"); return TDCGPU_ERR_ARG; }
    if (!src || !(c.have & which)) { set_error("structure 0x%x has not been built", which); return TDCGPU_ERR_STATE; }
    if (elem != 4 || width < 1 || width > 64) { set_error("get_packed: 32-bit arrays only, 1 <= width <= 64"); return TDCGPU_ERR_ARG; }
    const u64 nwords = div_up(c.n * u64(width), 64);
    ComputeLock lock(c.device);
    if (cap_words < nwords) { set_error("get_packed: buffer too small: %llu < %llu words", (unsigned long long)cap_words, (unsigned long long)nwords); return TDCGPU_ERR_ARG; }
    c.arena.reset();  // scratch: whatever the previous phase kept there (e.g. the encoder's masks) is stale from here on
    u64* packed = c.arena.take<u64>(nwords);
    if (!packed) { set_error("get_packed: scratch arena too small"); return TDCGPU_ERR_NOMEM; }
    TDC_LAUNCH(pack_bits_kernel, u32(div_up(nwords, 256)), 256, 0, c.stream, static_cast<const u32*>(src), c.n, width, packed, nwords);
    prof_add_bytes("pack_bits_kernel", double(c.n) * 4 + double(nwords) * 8);
    TDC_KCHECK();
    if (!to_device) return host_copy(c, dst, packed, nwords * sizeof(u64), false);
    TDC_CUDA(cudaMemcpyAsync(dst, packed, nwords * sizeof(u64), cudaMemcpyDeviceToDevice, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

const void* tdcgpu_textds_device_ptr(tdcgpu_ctx* ctx, uint32_t which) {
    if (!ctx) return nullptr;
    size_t elem;
    void* p = array_ptr(ctx->c, which, &elem);
    return (ctx->c.have & which) ? p : nullptr;
}

int tdcgpu_textds_max_lcp(tdcgpu_ctx* ctx, uint32_t* max_lcp) {
    API_GUARD(ctx);
    if (!(c.have & (DS_PLCP | DS_LCP))) { set_error("neither PLCP nor LCP has been built"); return TDCGPU_ERR_STATE; }
    if (max_lcp) *max_lcp = c.max_lcp;
    return 0;
}

int tdcgpu_lzss_lcp_factorize(tdcgpu_ctx* ctx, uint32_t threshold, uint64_t* count, uint32_t* min_len, uint32_t* max_len) {
    API_GUARD(ctx);
    ComputeLock lock(c.device);
    c.phases.clear();
    if (threshold < 1) { set_error("lzss_lcp: threshold must be >= 1"); return TDCGPU_ERR_ARG; }
    TDC_TRY(do_build(c, DS_SA | DS_ISA | DS_LCP));
    {
        PhaseTimer t(c, "Factorize");
        TDC_TRY(factorize_lzss_lcp(c, threshold));
    }
    if (count) *count = c.num_factors;
    if (min_len) *min_len = c.flen_min;
    if (max_len) *max_len = c.flen_max;
    return 0;
}

int tdcgpu_lzss_lcp_get_factors(tdcgpu_ctx* ctx, tdcgpu_factor* dst, uint64_t cap, int to_device) {
    API_GUARD(ctx);
    if (!c.have_factors) { set_error("no factor list (call tdcgpu_lzss_lcp_factorize first)"); return TDCGPU_ERR_STATE; }
    if (c.num_factors > cap) { set_error("factor buffer too small: %llu > %llu", (unsigned long long)c.num_factors, (unsigned long long)cap); return TDCGPU_ERR_ARG; }
    if (c.num_factors == 0) return 0;
    if (!dst) { set_error("null destination"); return TDCGPU_ERR_ARG; }
    if (!to_device) return host_copy(c, dst, c.d_factors, sizeof(Factor) * c.num_factors, false);
    TDC_CUDA(cudaMemcpyAsync(dst, c.d_factors, sizeof(Factor) * c.num_factors, cudaMemcpyDeviceToDevice, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_lzss_literal_histogram(tdcgpu_ctx* ctx, uint64_t hist[256], uint64_t* fdist_max) {
    API_GUARD(ctx);
    ComputeLock lock(c.device);
    c.phases.clear();
    {
        PhaseTimer t(c, "Encode: literal histogram");
        TDC_TRY(encode_prepare(c));
    }
    if (hist) memcpy(hist, c.enc.hist, sizeof(uint64_t) * 256);
    if (fdist_max) *fdist_max = c.enc.fdist_max;
    return 0;
}

int tdcgpu_lzss_encode(tdcgpu_ctx* ctx, const uint64_t codes[256], const uint8_t lens[256], uint32_t lead_bits,
                       uint8_t lead_byte, uint64_t* nbits) {
    API_GUARD(ctx);
    if (!codes || !lens) { set_error("null code table"); return TDCGPU_ERR_ARG; }
    ComputeLock lock(c.device);
    c.phases.clear();
    {
        PhaseTimer t(c, "Encode: bit stream");
        TDC_TRY(encode_lzss(c, codes, lens, lead_bits, lead_byte));
    }
    if (nbits) *nbits = c.enc.nbits;
    return 0;
}

// Copy a device bit stream of `nbits` bits out; finalize appends BitOStream::~BitOStream's tail (io/BitOStream.hpp:53-64):
// the number of bits used in the current byte goes into its low 3 bits if they are free (used <= 5; an untouched byte is
// written as 0), otherwise the byte is flushed and the count follows in a byte of its own.
static int copy_bitstream_out(Ctx& c, const uint8_t* d_stream, u64 nbits, uint8_t* dst, u64 cap, int finalize, u64* nbytes, int to_device) {
    const u64 whole = nbits / 8;
    const u32 used = u32(nbits % 8);
    const u64 total = finalize ? whole + (used <= 5 ? 1 : 2) : whole + (used ? 1 : 0);
    if (nbytes) *nbytes = total;
    if (total > cap) { set_error("encode buffer too small: %llu > %llu", (unsigned long long)total, (unsigned long long)cap); return TDCGPU_ERR_ARG; }
    if (!dst) { set_error("null destination"); return TDCGPU_ERR_ARG; }
    const u64 body = whole + (used ? 1 : 0);
    if (body && to_device) TDC_CUDA(cudaMemcpyAsync(dst, d_stream, body, cudaMemcpyDeviceToDevice, c.stream));
    if (body && !to_device) TDC_TRY(host_copy(c, dst, d_stream, body, false));
    if (finalize) {
        uint8_t tail[2] = {0, 0};
        if (used) TDC_CUDA(cudaMemcpyAsync(&tail[0], d_stream + whole, 1, cudaMemcpyDeviceToHost, c.stream));
        TDC_CUDA(cudaStreamSynchronize(c.stream));
        u64 ntail;
        if (used <= 5) { tail[0] |= uint8_t(used); ntail = 1; } else { tail[1] = uint8_t(used); ntail = 2; }
        if (to_device) TDC_CUDA(cudaMemcpyAsync(dst + whole, tail, ntail, cudaMemcpyHostToDevice, c.stream));
        else memcpy(dst + whole, tail, ntail);
    }
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

// Bytes [offset, offset + cap) of the (optionally finalized) stream to a HOST buffer: lets a caller drain a large archive
// through one small pinned buffer straight into its output stream.  *total = length of the whole stream, *written = bytes
// stored at dst by this call.
static int copy_bitstream_chunk(Ctx& c, const uint8_t* d_stream, u64 nbits, u64 offset, uint8_t* dst, u64 cap, int finalize, u64* total_out,
                                u64* written) {
    const u64 whole = nbits / 8;
    const u32 used = u32(nbits % 8);
    const u64 body = whole + (used ? 1 : 0);
    const u64 total = finalize ? whole + (used <= 5 ? 1 : 2) : body;
    if (total_out) *total_out = total;
    if (written) *written = 0;
    if (offset >= total || cap == 0) return 0;
    if (!dst) { set_error("null destination"); return TDCGPU_ERR_ARG; }
    const u64 end = std::min(total, offset + cap);
    if (offset < body) TDC_TRY(host_copy(c, dst, d_stream + offset, std::min(end, body) - offset, false));
    if (finalize && end > whole) {  // BitOStream::~BitOStream's tail (io/BitOStream.hpp:53-64), see copy_bitstream_out
        uint8_t tail[2] = {0, 0};
        if (used) {
            TDC_CUDA(cudaMemcpyAsync(&tail[0], d_stream + whole, 1, cudaMemcpyDeviceToHost, c.stream));
            TDC_CUDA(cudaStreamSynchronize(c.stream));
        }
        if (used <= 5) tail[0] |= uint8_t(used); else tail[1] = uint8_t(used);
        for (u64 p = std::max(offset, whole); p < end; p++) dst[p - offset] = tail[p - whole];
    }
    if (written) *written = end - offset;
    return 0;
}

int tdcgpu_lzss_encode_get_chunk(tdcgpu_ctx* ctx, uint64_t offset, uint8_t* dst, uint64_t cap, int finalize, uint64_t* total,
                                 uint64_t* written) {
    API_GUARD(ctx);
    if (!c.enc.encoded || c.enc.gen != c.arena.gen) { set_error("no encoded stream (call tdcgpu_lzss_encode first)"); return TDCGPU_ERR_STATE; }
    return copy_bitstream_chunk(c, c.enc.out, c.enc.nbits, offset, dst, cap, finalize, total, written);
}

int tdcgpu_set_len_bits(tdcgpu_ctx* ctx, uint32_t len_field_bits) {
    API_GUARD(ctx);
    if (len_field_bits != 32 && len_field_bits != 64) { set_error("len_field_bits must be 32 or 64"); return TDCGPU_ERR_ARG; }
    c.len_field_bits = len_field_bits;
    return 0;
}

void* tdcgpu_pinned_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); set_error("pinned allocation of %llu bytes failed", (unsigned long long)bytes); return nullptr; }
    return p;
}
void tdcgpu_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}
void* tdcgpu_device_alloc(tdcgpu_ctx* ctx, uint64_t bytes) {
    if (!ctx) { set_error("null context"); return nullptr; }
    void* p = nullptr;
    if (cudaSetDevice(ctx->c.device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        set_error("device allocation of %llu bytes failed", (unsigned long long)bytes);
        return nullptr;
    }
    return p;
}
void tdcgpu_device_free(tdcgpu_ctx* ctx, void* p) {
    if (!p) return;
    if (ctx) cudaSetDevice(ctx->c.device);
    cudaFree(p);
}
int tdcgpu_device_copy(tdcgpu_ctx* ctx, void* dst, const void* src, uint64_t bytes, int kind) {
    API_GUARD(ctx);
    if (bytes == 0) return 0;
    if (!dst || !src || kind < 0 || kind > 2) { set_error("device_copy: bad argument"); return TDCGPU_ERR_ARG; }
    if (kind == 2) {
        TDC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c.stream));
        TDC_CUDA(cudaStreamSynchronize(c.stream));
        return 0;
    }
    return host_copy(c, dst, src, bytes, kind == 0);
}
int tdcgpu_set_device(int device) {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); set_error("cannot select device %d", device); return TDCGPU_ERR_CUDA; }
    return 0;
}

int tdcgpu_lzss_encode_get(tdcgpu_ctx* ctx, uint8_t* dst, uint64_t cap, int finalize, uint64_t* nbytes, int to_device) {
    API_GUARD(ctx);
    if (!c.enc.encoded || c.enc.gen != c.arena.gen) { set_error("no encoded stream (call tdcgpu_lzss_encode first)"); return TDCGPU_ERR_STATE; }
    return copy_bitstream_out(c, c.enc.out, c.enc.nbits, dst, cap, finalize, nbytes, to_device);
}

// where the caller's in / out buffers of a stream stage live (include/tdcgpu.h: TDCGPU_BUF_*)
static inline bool in_on_device(int on_device) { return on_device == TDCGPU_BUF_DEVICE || on_device == TDCGPU_BUF_IN_DEVICE; }
static inline bool out_on_device(int on_device) { return on_device == TDCGPU_BUF_DEVICE || on_device == TDCGPU_BUF_OUT_DEVICE; }

// in/out staging of the stream stages: [in (n) | out (out_cap) | scratch]; a side that already is on the device is used in place
static int stream_stage_in(Ctx& c, const uint8_t* in, u64 n, u64 out_cap, size_t scratch, int on_device, const uint8_t** d_in, uint8_t** d_out) {
    if (on_device < 0 || on_device > 3) { set_error("bad on_device value %d", on_device); return TDCGPU_ERR_ARG; }
    const bool ind = in_on_device(on_device), outd = out_on_device(on_device);
    TDC_TRY(stream_arena_reserve(c, size_t(ind ? 0 : n + 256) + size_t(outd ? 0 : out_cap + 256) + scratch + 4096));
    if (ind) {
        *d_in = in;
    } else {
        uint8_t* di = c.stream_arena.take<uint8_t>(n + 16);
        if (!di) { set_error("stream scratch too small"); return TDCGPU_ERR_NOMEM; }
        if (n) TDC_TRY(host_copy(c, di, in, n, true));
        *d_in = di;
    }
    if (!outd) {
        *d_out = c.stream_arena.take<uint8_t>(out_cap + 16);
        if (!*d_out) { set_error("stream scratch too small"); return TDCGPU_ERR_NOMEM; }
    }
    return 0;
}

int tdcgpu_mtf_encode(tdcgpu_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, int on_device) {
    API_GUARD(ctx);
    if (n == 0) return 0;
    if (!in || !out) { set_error("null buffer"); return TDCGPU_ERR_ARG; }
    c.phases.clear();
    const uint8_t* d_in = nullptr;
    uint8_t* d_out = out;
    TDC_TRY(stream_stage_in(c, in, n, n, mtf_scratch_bytes(n), on_device, &d_in, &d_out));
    {
        PhaseTimer t(c, "MTF");
        TDC_TRY(mtf_encode_device(c, d_in, n, d_out));
    }
    if (!out_on_device(on_device)) TDC_TRY(host_copy(c, out, d_out, n, false));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_rle_encode(tdcgpu_ctx* ctx, const uint8_t* in, uint64_t n, uint64_t offset, uint8_t* out, uint64_t cap, uint64_t* out_n,
                      int on_device) {
    API_GUARD(ctx);
    if (out_n) *out_n = 0;
    if (n == 0) return 0;
    if (!in || !out) { set_error("null buffer"); return TDCGPU_ERR_ARG; }
    c.phases.clear();
    const u64 worst = rle_max_output(n, offset);
    const bool outd = out_on_device(on_device);
    if (outd && cap < worst) { set_error("rle: a device output buffer must hold the worst case (%llu bytes)", (unsigned long long)worst); return TDCGPU_ERR_ARG; }
    const uint8_t* d_in = nullptr;
    uint8_t* d_out = out;
    TDC_TRY(stream_stage_in(c, in, n, worst, rle_scratch_bytes(n), on_device, &d_in, &d_out));
    u64 produced = 0;
    {
        PhaseTimer t(c, "RLE");
        TDC_TRY(rle_encode_device(c, d_in, n, offset, d_out, &produced));
    }
    if (out_n) *out_n = produced;
    if (!outd) {
        if (produced > cap) { set_error("rle: output buffer too small: %llu > %llu", (unsigned long long)produced, (unsigned long long)cap); return TDCGPU_ERR_ARG; }
        TDC_TRY(host_copy(c, out, d_out, produced, false));
    }
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_literal_encode_begin(tdcgpu_ctx* ctx, const uint8_t* in, uint64_t n, int on_device, uint64_t hist[256]) {
    API_GUARD(ctx);
    if (n && !in) { set_error("null buffer"); return TDCGPU_ERR_ARG; }
    if (!hist) { set_error("null histogram"); return TDCGPU_ERR_ARG; }
    c.phases.clear();
    c.lit.staged = c.lit.encoded = false;
    const uint8_t* d_in = nullptr;
    uint8_t* d_unused = nullptr;
    TDC_TRY(stream_stage_in(c, in, n, 0, literal_scratch_bytes(n), on_device ? TDCGPU_BUF_DEVICE : TDCGPU_BUF_HOST, &d_in, &d_unused));
    {
        PhaseTimer t(c, "Literal histogram");
        TDC_TRY(stream_histogram_device(c, d_in, n, hist));
    }
    c.lit.d_in = d_in;
    c.lit.n = n;
    c.lit.gen = c.stream_arena.gen;
    c.lit.arena_mark = c.stream_arena.off;
    c.lit.staged = true;
    return 0;
}

int tdcgpu_literal_encode(tdcgpu_ctx* ctx, const uint64_t codes[256], const uint8_t lens[256], uint32_t lead_bits, uint8_t lead_byte,
                          uint64_t* nbits) {
    API_GUARD(ctx);
    if (!codes || !lens) { set_error("null code table"); return TDCGPU_ERR_ARG; }
    if (!c.lit.staged || c.lit.gen != c.stream_arena.gen) { set_error("no staged input (call tdcgpu_literal_encode_begin first)"); return TDCGPU_ERR_STATE; }
    c.phases.clear();
    c.stream_arena.off = c.lit.arena_mark;  // re-encoding the staged input (e.g. with another code table) reuses the scratch
    {
        PhaseTimer t(c, "Literal encode");
        TDC_TRY(literal_encode_device(c, c.lit.d_in, c.lit.n, codes, lens, lead_bits, lead_byte, &c.lit.nbits));
    }
    c.lit.encoded = true;
    if (nbits) *nbits = c.lit.nbits;
    return 0;
}

int tdcgpu_literal_encode_get_chunk(tdcgpu_ctx* ctx, uint64_t offset, uint8_t* dst, uint64_t cap, int finalize, uint64_t* total,
                                    uint64_t* written) {
    API_GUARD(ctx);
    if (!c.lit.encoded) { set_error("no encoded stream (call tdcgpu_literal_encode first)"); return TDCGPU_ERR_STATE; }
    return copy_bitstream_chunk(c, c.lit.d_out, c.lit.nbits, offset, dst, cap, finalize, total, written);
}

int tdcgpu_literal_encode_get(tdcgpu_ctx* ctx, uint8_t* dst, uint64_t cap, int finalize, uint64_t* nbytes, int to_device) {
    API_GUARD(ctx);
    if (!c.lit.encoded) { set_error("no encoded stream (call tdcgpu_literal_encode first)"); return TDCGPU_ERR_STATE; }
    return copy_bitstream_out(c, c.lit.d_out, c.lit.nbits, dst, cap, finalize, nbytes, to_device);
}

int tdcgpu_textds_build_host(int device, const uint8_t* text, uint64_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp,
                             uint32_t* phi, uint32_t* plcp, uint32_t* max_lcp) {
    tdcgpu_ctx* h = nullptr;
    int rc = tdcgpu_create(device, &h);
    if (rc < 0) return rc;
    uint32_t flags = 0;
    if (sa) flags |= DS_SA;
    if (isa) flags |= DS_ISA;
    if (lcp) flags |= DS_LCP;
    if (phi) flags |= DS_PHI;
    if (plcp || max_lcp) flags |= DS_PLCP;
    rc = tdcgpu_set_text(h, text, n, 0);
    if (rc == 0) rc = tdcgpu_textds_build(h, flags);
    if (rc == 0 && sa) rc = tdcgpu_textds_get(h, DS_SA, sa, 0);
    if (rc == 0 && isa) rc = tdcgpu_textds_get(h, DS_ISA, isa, 0);
    if (rc == 0 && lcp) rc = tdcgpu_textds_get(h, DS_LCP, lcp, 0);
    if (rc == 0 && phi) rc = tdcgpu_textds_get(h, DS_PHI, phi, 0);
    if (rc == 0 && plcp) rc = tdcgpu_textds_get(h, DS_PLCP, plcp, 0);
    if (rc == 0 && max_lcp) rc = tdcgpu_textds_max_lcp(h, max_lcp);
    tdcgpu_destroy(h);
    return rc;
}

int tdcgpu_bwt_host(int device, const uint8_t* text, uint64_t n, uint8_t* out) {
    tdcgpu_ctx* h = nullptr;
    int rc = tdcgpu_create(device, &h);
    if (rc < 0) return rc;
    rc = tdcgpu_set_text(h, text, n, 0);
    if (rc == 0) rc = tdcgpu_textds_build(h, DS_SA | DS_BWT);
    if (rc == 0) rc = tdcgpu_textds_get(h, DS_BWT, out, 0);
    tdcgpu_destroy(h);
    return rc;
}

int tdcgpu_phase_count(tdcgpu_ctx* ctx) { return ctx ? int(ctx->c.phases.size()) : 0; }
const char* tdcgpu_phase_name(tdcgpu_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= int(ctx->c.phases.size())) return nullptr;
    return ctx->c.phases[i].name.c_str();
}
float tdcgpu_phase_ms(tdcgpu_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= int(ctx->c.phases.size())) return -1.f;
    return ctx->c.phases[i].ms;
}

int tdcgpu_sa_stats(tdcgpu_ctx* ctx, uint64_t out[8]) {
    API_GUARD(ctx);
    out[0] = c.sa_rounds;
    out[1] = c.sa_active_sum;
    out[2] = c.sortws.stat_passes;
    out[3] = c.sortws.stat_elems;
    out[4] = c.alphabet;
    out[5] = c.symbols_per_key;
    out[6] = c.lcp_route;
    out[7] = uint64_t(c.sa_prefix_work);
    return 0;
}

int tdcgpu_sa_layout(tdcgpu_ctx* ctx, uint64_t out[4]) {
    API_GUARD(ctx);
    out[0] = c.sa_packed ? 1 : 0;
    out[1] = c.sa_key_bits;
    out[2] = c.sa_packed ? bits_for_host(c.n > 1 ? c.n - 1 : 1) : 0;
    out[3] = c.symbols_per_key;
    return 0;
}

int tdcgpu_sync(tdcgpu_ctx* ctx) {
    API_GUARD(ctx);
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int tdcgpu_event_record(tdcgpu_ctx* ctx, int slot) {
    API_GUARD(ctx);
    if (slot < 0 || slot >= 8) { set_error("event slot out of range"); return TDCGPU_ERR_ARG; }
    if (!c.user_events[slot]) TDC_CUDA(cudaEventCreate(&c.user_events[slot]));
    TDC_CUDA(cudaEventRecord(c.user_events[slot], c.stream));
    return 0;
}

int tdcgpu_event_elapsed_ms(tdcgpu_ctx* ctx, int slot_a, int slot_b, float* ms) {
    API_GUARD(ctx);
    if (slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !c.user_events[slot_a] || !c.user_events[slot_b] || !ms) {
        set_error("bad event slots");
        return TDCGPU_ERR_ARG;
    }
    TDC_CUDA(cudaEventSynchronize(c.user_events[slot_b]));
    TDC_CUDA(cudaEventElapsedTime(ms, c.user_events[slot_a], c.user_events[slot_b]));
    return 0;
}

uint64_t tdcgpu_launch_count(void) { return g_launches.load(); }

void tdcgpu_profile_enable(int on) { g_prof_on = on != 0; }

void tdcgpu_profile_reset(void) {
    prof_resolve();
    g_prof.clear();
}

int tdcgpu_profile_count(void) {
    prof_resolve();
    return int(g_prof.size());
}

int tdcgpu_profile_entry(int i, const char** name, uint64_t* launches, double* ms, double* bytes) {
    if (i < 0 || i >= int(g_prof.size())) { set_error("profile index out of range"); return TDCGPU_ERR_ARG; }
    if (name) *name = g_prof[i].name.c_str();
    if (launches) *launches = g_prof[i].launches;
    if (ms) *ms = g_prof[i].ms;
    if (bytes) *bytes = g_prof[i].bytes;
    return 0;
}

}  // extern "C"
