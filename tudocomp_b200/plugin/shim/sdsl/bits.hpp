// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for the one SDSL header the hot path pulls in
// (sdsl-lite is downloaded by the reference's CMake, cmakemodules/DownloadSDSL.cmake:1-6; no network here).
// Provides the bit-field accessors used by the reference's bit-packed vectors
// (include/tudocomp/ds/IntRepr.hpp:19-25, IntPtr.hpp:132-139, BitPackingVector.hpp:112,505-523):
// integers of `len` bits stored LSB-first starting at bit `offset` of a 64-bit word, spilling into the next word.
#pragma once
#include <cstdint>

namespace sdsl {
struct bits {
    static inline uint64_t mask(uint8_t len) { return len >= 64 ? ~uint64_t(0) : ((uint64_t(1) << len) - 1); }

    static inline uint64_t read_int(const uint64_t* word, uint8_t offset = 0, uint8_t len = 64) {
        uint64_t lo = (*word) >> offset;
        if (offset + len > 64) {
            lo |= word[1] << (64 - offset);
        }
        return lo & mask(len);
    }

    static inline void write_int(uint64_t* word, uint64_t x, uint8_t offset = 0, uint8_t len = 64) {
        x &= mask(len);
        if (offset + len <= 64) {
            *word = (*word & ~(mask(len) << offset)) | (x << offset);
        } else {
            const uint8_t first = 64 - offset;
            word[0] = (word[0] & ~(~uint64_t(0) << offset)) | (x << offset);
            const uint8_t rest = len - first;
            word[1] = (word[1] & ~mask(rest)) | (x >> first);
        }
    }

    static inline void move_right(const uint64_t*& word, uint8_t& offset, uint8_t len) {
        unsigned o = unsigned(offset) + len;
        word += o >> 6;
        offset = uint8_t(o & 63);
    }

    static inline void move_left(const uint64_t*& word, uint8_t& offset, uint8_t len) {
        int o = int(offset) - int(len);
        if (o < 0) {
            // len <= 64, so at most one word back
            o += 64;
            --word;
        }
        offset = uint8_t(o);
    }

    static inline uint64_t read_int_and_move(const uint64_t*& word, uint8_t& offset, uint8_t len = 64) {
        uint64_t v = read_int(word, offset, len);
        move_right(word, offset, len);
        return v;
    }

    static inline void write_int_and_move(uint64_t*& word, uint64_t x, uint8_t& offset, uint8_t len) {
        write_int(word, x, offset, len);
        const uint64_t* w = word;
        move_right(w, offset, len);
        word = const_cast<uint64_t*>(w);
    }
};
}  // namespace sdsl
