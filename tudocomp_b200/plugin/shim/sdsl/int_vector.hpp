// TEST / OFFLINE-BUILD INFRASTRUCTURE ONLY.  Stand-in for <sdsl/int_vector.hpp>: the reference's lcpcomp decoder
// (include/tudocomp/compressors/lcpcomp/decompress/ScanDec.hpp:28-46) uses sdsl::bit_vector with a rank_1 support;
// sdsl-lite is fetched by the reference's CMake (cmakemodules/DownloadSDSL.cmake) and is not available offline.
// Only the members that code touches: bit_vector(size, value), size(), operator[] (read / assign), rank_1_type.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sdsl {

class rank_support_v_shim;

class bit_vector {
    std::vector<uint64_t> m_words;
    size_t m_size = 0;

public:
    typedef rank_support_v_shim rank_1_type;
    struct reference {
        uint64_t* w;
        uint64_t mask;
        operator bool() const { return (*w & mask) != 0; }
        reference& operator=(bool b) {
            if (b) *w |= mask; else *w &= ~mask;
            return *this;
        }
    };
    bit_vector() {}
    bit_vector(size_t size, uint64_t value) : m_words((size + 63) / 64, value ? ~uint64_t(0) : 0), m_size(size) {}
    size_t size() const { return m_size; }
    bool operator[](size_t i) const { return (m_words[i >> 6] >> (i & 63)) & 1u; }
    reference operator[](size_t i) { return reference{&m_words[i >> 6], uint64_t(1) << (i & 63)}; }
    const uint64_t* data() const { return m_words.data(); }
};

// rank(i) = number of set bits in [0, i)
class rank_support_v_shim {
    const bit_vector* m_bv = nullptr;
    std::vector<uint64_t> m_block;  // ones before word w

public:
    rank_support_v_shim() {}
    explicit rank_support_v_shim(const bit_vector* bv) : m_bv(bv) {
        const size_t words = (bv->size() + 63) / 64;
        m_block.resize(words + 1);
        uint64_t run = 0;
        for (size_t w = 0; w < words; w++) {
            m_block[w] = run;
            run += uint64_t(__builtin_popcountll(bv->data()[w]));
        }
        m_block[words] = run;
    }
    size_t rank(size_t i) const {
        const size_t w = i >> 6, r = i & 63;
        uint64_t res = m_block[w];
        if (r) res += uint64_t(__builtin_popcountll(m_bv->data()[w] & ((uint64_t(1) << r) - 1)));
        return size_t(res);
    }
    size_t operator()(size_t i) const { return rank(i); }
};

}  // namespace sdsl
