// GPU-backed PROVIDERS for the reference's own text index: the classes TextDS<sa_t, phi_t, plcp_t, lcp_t, isa_t>
// (ds/TextDS.hpp:23-29) takes as template parameters, next to SADivSufSort / PhiFromSA / PLCPFromPhi / LCPFromPLCP / ISAFromSA:
//
//   -a "lzss_lcp(coder=huff, textds=textds(sa=gpu))"                     only the suffix array from the GPU
//   -a "lzss_lcp(coder=huff, textds=textds(sa=gpu, lcp=gpu, isa=gpu))"   everything lzss_lcp asks for from the GPU
//
// Where GpuTextDS (GpuTextDS.hpp) replaces the whole text_t — and with it lets the factorisation and the encoder run on
// the device too — these plug into the unchanged TextDS and the unchanged compressors: the arrays come back as the
// DynamicIntVector an ArrayDS is, everything downstream is the reference's CPU code.  Same Meta types ("sa", "phi", "plcp",
// "lcp", "isa"), name "gpu"; same constructor signature, restrictions(), compress(), max_lcp() as the classes they stand
// next to (ds/SADivSufSort.hpp:13-64, PhiFromSA.hpp, PLCPFromPhi.hpp:20-75, LCPFromPLCP.hpp:14-75, ISAFromSA.hpp:13-62).
//
// TextDS constructs its providers one after the other for the same text.  They share the thread's cached device context:
// tdcgpu_set_text_cached uploads the text and keeps what is already built if the resident text has exactly these bytes, so
// `sa=gpu, lcp=gpu, isa=gpu` builds the suffix array once.
#pragma once

#include <tudocomp/Algorithm.hpp>
#include <tudocomp/ds/ArrayDS.hpp>
#include <tudocomp/ds/CompressMode.hpp>
#include <tudocomp/ds/TextDSFlags.hpp>
#include <tudocomp_stat/StatPhase.hpp>

#include "GpuTextDS.hpp"

namespace tdc {

namespace gpu_detail {
// One array of the text index of `t`, at the width the reference's provider ends with: INDEX_FAST_BITS for
// CompressMode::plain, else bits_for(n) (SA, ISA, Phi) / bits_for(max_lcp) (PLCP, LCP) — packed on the device.
template <typename textds_t>
inline DynamicIntVector provider_array(const textds_t& t, uint32_t which, CompressMode cm, bool lcp_width, len_t* max_lcp, const char* what) {
    const size_t n = t.size();
    StreamCtx g;  // the thread's cached context
    int reused = 0;
    check(tdcgpu_set_text_cached(g.ctx, reinterpret_cast<const uint8_t*>(t.text()), n, &reused), what);
    check(tdcgpu_textds_build(g.ctx, which), what);
    log_phases(g.ctx);
    StatPhase::log("gpu_text_reused", size_t(reused));
    uint32_t mx = 0;
    if (lcp_width) check(tdcgpu_textds_max_lcp(g.ctx, &mx), what);
    if (max_lcp) *max_lcp = len_t(mx);
    const uint8_t width = cm == CompressMode::plain ? uint8_t(INDEX_FAST_BITS) : uint8_t(lcp_width ? bits_for(mx) : bits_for(n));
    DynamicIntVector iv(n, 0, width);
    if (width == 32) check(tdcgpu_textds_get(g.ctx, which, iv.data(), 0), what);
    else check(tdcgpu_textds_get_packed(g.ctx, which, width, iv.data(), (uint64_t(n) * width + 63) / 64, 0), what);
    StatPhase::log("bit_width", size_t(iv.width()));
    StatPhase::log("size", iv.bit_size() / 8);
    return iv;
}
}  // namespace gpu_detail

#define TDCGPU_PROVIDER_COMPRESS(TITLE, WIDTH_EXPR)                   \
    inline void compress() {                                          \
        debug_check_array_is_initialized();                           \
        StatPhase::wrap(TITLE, [this] {                               \
            width(WIDTH_EXPR);                                        \
            shrink_to_fit();                                          \
            StatPhase::log("bit_width", size_t(width()));             \
            StatPhase::log("size", bit_size() / 8);                   \
        });                                                           \
    }

/// Suffix array from the GPU; stands next to SADivSufSort (ds/SADivSufSort.hpp).
class GpuSA : public Algorithm, public ArrayDS {
public:
    inline static Meta meta() { return Meta("sa", "gpu"); }
    inline static ds::InputRestrictions restrictions() { return ds::InputRestrictions{{0}, true}; }

    template <typename textds_t>
    inline GpuSA(Env&& env, const textds_t& t, CompressMode cm) : Algorithm(std::move(env)) {
        StatPhase::wrap("Construct SA", [&] { set_array(gpu_detail::provider_array(t, TDCGPU_SA, cm, false, nullptr, "sa")); });
        if (cm == CompressMode::compressed || cm == CompressMode::delayed) compress();
    }
    TDCGPU_PROVIDER_COMPRESS("Compress SA", bits_for(size()))
};

/// Inverse suffix array from the GPU (a by-product of its suffix sorting); stands next to ISAFromSA.
class GpuISA : public Algorithm, public ArrayDS {
public:
    inline static Meta meta() { return Meta("isa", "gpu"); }
    inline static ds::InputRestrictions restrictions() { return ds::InputRestrictions{{0}, true}; }

    template <typename textds_t>
    inline GpuISA(Env&& env, textds_t& t, CompressMode cm) : Algorithm(std::move(env)) {
        StatPhase::wrap("Construct ISA", [&] { set_array(gpu_detail::provider_array(t, TDCGPU_ISA, cm, false, nullptr, "isa")); });
        if (cm == CompressMode::delayed) compress();
    }
    TDCGPU_PROVIDER_COMPRESS("Compress ISA", bits_for(size()))
};

/// Phi array from the GPU; stands next to PhiFromSA.
class GpuPhi : public Algorithm, public ArrayDS {
public:
    inline static Meta meta() { return Meta("phi", "gpu"); }
    inline static ds::InputRestrictions restrictions() { return ds::InputRestrictions{{0}, true}; }

    template <typename textds_t>
    inline GpuPhi(Env&& env, textds_t& t, CompressMode cm) : Algorithm(std::move(env)) {
        StatPhase::wrap("Construct Phi Array", [&] { set_array(gpu_detail::provider_array(t, TDCGPU_PHI, cm, false, nullptr, "phi")); });
        if (cm == CompressMode::delayed) compress();
    }
    TDCGPU_PROVIDER_COMPRESS("Compress Phi Array", bits_for(size()))
};

/// PLCP array from the GPU (incl. the reference's stale PLCP[n-1], ds/PLCPFromPhi.hpp:38); stands next to PLCPFromPhi.
class GpuPLCP : public Algorithm, public ArrayDS {
    len_t m_max = 0;

public:
    inline static Meta meta() { return Meta("plcp", "gpu"); }
    inline static ds::InputRestrictions restrictions() { return ds::InputRestrictions{{0}, true}; }

    template <typename textds_t>
    inline GpuPLCP(Env&& env, textds_t& t, CompressMode cm) : Algorithm(std::move(env)) {
        StatPhase::wrap("Construct PLCP Array", [&] { set_array(gpu_detail::provider_array(t, TDCGPU_PLCP, cm, true, &m_max, "plcp")); });
        if (cm == CompressMode::compressed || cm == CompressMode::delayed) compress();
    }
    inline len_t max_lcp() const { return m_max; }
    TDCGPU_PROVIDER_COMPRESS("Compress PLCP Array", bits_for(m_max))
};

/// LCP array from the GPU; stands next to LCPFromPLCP.
class GpuLCP : public Algorithm, public ArrayDS {
    len_t m_max = 0;

public:
    inline static Meta meta() { return Meta("lcp", "gpu"); }
    inline static ds::InputRestrictions restrictions() { return ds::InputRestrictions{{0}, true}; }

    template <typename textds_t>
    inline GpuLCP(Env&& env, textds_t& t, CompressMode cm) : Algorithm(std::move(env)) {
        StatPhase::wrap("Construct LCP Array", [&] { set_array(gpu_detail::provider_array(t, TDCGPU_LCP, cm, true, &m_max, "lcp")); });
        if (cm == CompressMode::delayed) compress();
    }
    inline len_t max_lcp() const { return m_max; }
    TDCGPU_PROVIDER_COMPRESS("Compress LCP Array", bits_for(m_max))
};

#undef TDCGPU_PROVIDER_COMPRESS

}  // namespace tdc
