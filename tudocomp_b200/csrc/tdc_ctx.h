// Host-side context shared by the tdcgpu translation units.  One context = one device + one stream + one text.
#pragma once
#include <string>
#include <vector>

#include "radix_sort.cuh"
#include "tdc_common.cuh"

namespace tdc {

// bit flags as in /root/reference/include/tudocomp/ds/TextDSFlags.hpp:10-15
enum : u32 { DS_SA = 0x01, DS_ISA = 0x02, DS_LCP = 0x04, DS_PHI = 0x08, DS_PLCP = 0x10, DS_BWT = 0x100 };

struct Factor {  // lzss::Factor, /root/reference/include/tudocomp/compressors/lzss/LZSSFactors.hpp:13-20 (packed 3 x u32)
    u32 pos, src, len;
};

struct PhaseTime {
    std::string name;
    float ms;
};

// A bump allocator over one device allocation.  SA construction, LCP construction and factorisation run one after the
// other and each re-carves the same scratch bytes.
struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, off = 0;
    u64 gen = 0;  // bumped by every reset: state carved before a reset is stale afterwards (EncodeState)
    void reset() { off = 0; gen++; }
    template <class T>
    T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (off + bytes > cap) return nullptr;
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};

// Device-side lzss::encode_text (lzss_encode.cu): masks and scans derived from the factor list, then the bit stream.
// Everything lives in the scratch arena and is valid only while arena.gen == gen.
struct EncodeState {
    u64 gen = 0;
    bool prepared = false, encoded = false;
    u32 *S = nullptr, *E = nullptr, *scan_s = nullptr, *scan_e = nullptr, *tile_bits = nullptr;
    u64* tile_off = nullptr;
    u64* d_code = nullptr;       // 256 code words
    uint8_t* d_len = nullptr;    // 256 code lengths
    uint8_t* out = nullptr;      // bit stream, MSB first
    u64 out_cap = 0;             // bytes
    u32 ntiles = 0;
    u32 fdist_max = 0;
    u64 hist[256];
    u64 nbits = 0;
};

// host_copy.cu: pinned double buffers + copy streams of the threads that stage pageable caller memory
static const int HC_THREADS = 4;
static const size_t HC_CHUNK = size_t(8) << 20;
static const size_t HC_MIN_STAGED = size_t(4) << 20;  // smaller copies: one plain cudaMemcpy
struct HostCopier {
    bool ready = false;
    uint8_t* buf[HC_THREADS][2] = {};
    cudaEvent_t ev[HC_THREADS][2] = {};
    cudaStream_t stream[HC_THREADS] = {};
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = 0;

    // text (device), padded with zero bytes up to a multiple of 16 plus 64
    u64 n = 0;
    u64 cap_n = 0;  // capacity the buffers below were sized for
    uint8_t* d_text = nullptr;

    // results (device, persistent until the next load)
    u32 *d_sa = nullptr, *d_isa = nullptr, *d_lcp = nullptr, *d_phi = nullptr, *d_plcp = nullptr;
    uint8_t* d_bwt = nullptr;
    u32 have = 0;  // DS_* bits that are valid
    u32 max_lcp = 0;

    // factor list
    Factor* d_factors = nullptr;
    u64 factors_cap = 0, num_factors = 0;
    u32 flen_min = 0xffffffffu, flen_max = 0;
    u32 len_field_bits = 32;  // width of the archive's text-length field (tdcgpu_set_len_bits)
    bool have_factors = false;  // factorize_lzss_lcp has run on the current text
    EncodeState enc;

    // scratch
    Arena arena;
    SortWorkspace sortws;
    u32* d_scalars = nullptr;  // small device scratch (256 u32)
    u32* h_scalars = nullptr;  // pinned mirror

    // pinned staging for host<->device copies of pageable caller buffers
    HostCopier copier;

    // byte-stream stages (stream_codecs.cu): one grow-only device buffer, independent of the text-index arrays
    Arena stream_arena;
    struct LiteralStage {  // tdcgpu_literal_encode_begin .. _get: the staged input and the encoded stream
        u64 gen = 0, n = 0, nbits = 0;
        size_t arena_mark = 0;  // stream_arena.off after staging: every (re-)encode carves its scratch from here
        const uint8_t* d_in = nullptr;
        uint8_t* d_out = nullptr;
        u64 out_cap = 0;
        bool staged = false, encoded = false;
    } lit;

    cudaEvent_t user_events[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    std::vector<u64> h_samples;  // host copy of the sorted prefix sample (key-length cost model)

    // stats
    std::vector<PhaseTime> phases;
    u32 sa_rounds = 0;
    u64 sa_active_sum = 0;
    u32 alphabet = 0, symbols_per_key = 0;
    bool sa_packed = false;  // the initial sort ran on packed 64-bit records (key | suffix)
    u32 sa_key_bits = 0;     // key bits of the initial sort
    u32 sa_tail_span = 0;    // packed keys without a length field: symbols per key (suffixes that close to the end were padded)
    double sa_prefix_work = 0;  // sum over doubling rounds of (active suffixes x already-known common prefix): LCP-sum estimate
    bool sa_lcp_seeded = false;  // d_lcp holds key-derived LCPs / LCP_UNKNOWN marks from the initial sort
    u64 sa_first_residue = 0;    // suffixes the initial sort left in groups
    u32 lcp_route = 0;          // 1 = direct comparison in SA order, 2 = Phi/PLCP route
};

// host_copy.cu
int host_copy(Ctx& c, void* dst, const void* src, size_t bytes, bool h2d);  // blocking; pageable memory is staged
void host_copier_free(HostCopier& hc);
// suffix_array.cu
int build_suffix_array(Ctx& c, bool want_lcp);  // fills d_sa and d_isa; want_lcp: seed d_lcp from the initial keys
// lcp.cu
int build_phi_bwt(Ctx& c, bool want_phi, bool want_bwt);
int build_plcp_lcp(Ctx& c, bool want_lcp);
int build_lcp_direct(Ctx& c);  // LCP without Phi/PLCP (texts with short common prefixes)
static const u32 LCP_UNKNOWN = 0xffffffffu;  // LCP slot of a pair the initial keys could not separate
// lzss_factorize.cu
int factorize_lzss_lcp(Ctx& c, u32 threshold);
// stream_codecs.cu — device pointers in, device pointers out; scratch from c.stream_arena (after the caller's buffers)
int stream_arena_reserve(Ctx& c, size_t bytes);
size_t mtf_scratch_bytes(u64 n);
int mtf_encode_device(Ctx& c, const uint8_t* d_in, u64 n, uint8_t* d_out);
size_t rle_scratch_bytes(u64 n);
u64 rle_max_output(u64 n, u64 offset);
int rle_encode_device(Ctx& c, const uint8_t* d_in, u64 n, u64 offset, uint8_t* d_out, u64* out_n);
size_t literal_scratch_bytes(u64 n);
int stream_histogram_device(Ctx& c, const uint8_t* d_in, u64 n, u64 hist[256]);
int literal_encode_device(Ctx& c, const uint8_t* d_in, u64 n, const u64* codes, const uint8_t* lens, u32 lead_bits, u32 lead_byte,
                          u64* nbits);
// lzss_encode.cu
int encode_prepare(Ctx& c);  // masks, scans, literal histogram, fdist_max of the current factor list
int encode_lzss(Ctx& c, const u64* codes, const uint8_t* lens, u32 lead_bits, u32 lead_byte);

struct PhaseTimer {  // CUDA-event timing of one phase on the context's stream
    Ctx& c;
    const char* name;
    cudaEvent_t a, b;
    PhaseTimer(Ctx& c_, const char* name_);
    ~PhaseTimer();
};

}  // namespace tdc

// the opaque handle of include/tdcgpu.h is a Ctx
struct tdcgpu_ctx {
    tdc::Ctx c;
};
