/*
 * tdcgpu — C ABI of the B200-native text index (SA / ISA / Phi / PLCP / LCP), BWT and lzss_lcp factoriser.
 *
 * This is the drop-in boundary for tudocomp's TextDS + lzss_lcp hot path.  Plain pointers and sizes only; no C++
 * or torch types; nothing throws or aborts across this boundary.  Every function returns 0 on success and a
 * negative code on failure (tdcgpu_last_error() describes it).  There is NO CPU fallback: without a CUDA device
 * tdcgpu_create fails with TDCGPU_ERR_CUDA.
 *
 * Each entry point names the reference interface it replaces (paths relative to the tudocomp source tree).
 *
 * Text contract (same as the reference, ds/TextDS.hpp:132-138, ds/SADivSufSort.hpp:21-26, driver.cpp:268-270):
 * `text` is the escaped input followed by exactly one 0 byte; n counts that byte; no other 0 occurs.
 * All indices are UNSIGNED 32-bit (len_t = uint32_t, def.hpp:103,114).  The reference's default build stops at n < 2^31
 * (divsufsort's sign bit); here n may reach 2^32 - 2^20 as long as common prefixes stay below 2^31 (checked).
 */
#ifndef TDCGPU_H
#define TDCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tdcgpu_ctx tdcgpu_ctx;

/* ds::SA/ISA/LCP/PHI/PLCP bit values — ds/TextDSFlags.hpp:10-15 */
#define TDCGPU_SA 0x01u
#define TDCGPU_ISA 0x02u
#define TDCGPU_LCP 0x04u
#define TDCGPU_PHI 0x08u
#define TDCGPU_PLCP 0x10u
/* additional product of the same pass over SA: bwt::bwt, ds/bwt.hpp:19-22 */
#define TDCGPU_BWT 0x100u

#define TDCGPU_ERR_CUDA (-1)      /* CUDA runtime error (incl. no device) */
#define TDCGPU_ERR_NOMEM (-2)     /* device scratch too small */
#define TDCGPU_ERR_SENTINEL (-3)  /* text violates the sentinel contract ("Input has no sentinel!", TextDS.hpp:132-138) */
#define TDCGPU_ERR_INTERNAL (-4)
#define TDCGPU_ERR_ARG (-5)       /* bad argument (threshold 0, NULL, n > 2^32 - 2^20, buffer too small) */
#define TDCGPU_ERR_STATE (-6)     /* a required structure has not been built */

/* lzss::Factor — compressors/lzss/LZSSFactors.hpp:13-20 (packed pos, src, len; len_compact_t = uint32_t) */
typedef struct tdcgpu_factor {
    uint32_t pos, src, len;
} tdcgpu_factor;

const char* tdcgpu_last_error(void);
int tdcgpu_device_count(void);

/* One context per device; owns a stream, the resident text, the index arrays and scratch. */
int tdcgpu_create(int device, tdcgpu_ctx** out);
void tdcgpu_destroy(tdcgpu_ctx* ctx);

/* Make `text` (n bytes incl. the trailing 0) the context's text.  Replaces the View handed to TextDS::TextDS
 * (ds/TextDS.hpp:130-147).
 * on_device == 0: `text` is a host pointer.  Pinned memory (cudaHostAlloc / cudaHostRegister) is copied directly; pageable
 * memory — what the driver's View is, io/RestrictedBuffer.hpp:108-181 — is staged chunk by chunk through the context's
 * pinned buffers by a few copy threads, so the transfer runs at PCIe speed.  The call returns when the copy is complete:
 * the caller may reuse its buffer.  A last byte != 0 is rejected here with TDCGPU_ERR_SENTINEL.
 * on_device != 0: `text` is a device pointer on the context's device (device-to-device copy on the context's stream).
 * The data must be complete before the call (the context's stream is not ordered after the caller's streams), and the
 * call returns after the copy has finished, so the source may be overwritten afterwards. */
int tdcgpu_set_text(tdcgpu_ctx* ctx, const uint8_t* text, uint64_t n, int on_device);

/* tdcgpu_set_text for a HOST text, except that the structures already built stay valid when the context's resident text
 * has exactly these bytes (the new text is uploaded next to it and compared on the device).  For callers that meet the same
 * text several times without being able to say so: the per-provider classes of the plugin (`textds(sa=gpu, lcp=gpu, ...)`,
 * tudocomp_gpu/GpuProviders.hpp), which the reference's TextDS constructs one after the other (ds/TextDS.hpp:247-292).
 * *reused = 1 if nothing had to be invalidated. */
int tdcgpu_set_text_cached(tdcgpu_ctx* ctx, const uint8_t* text, uint64_t n, int* reused);

/* Build the requested structures on the device (they stay resident).  Replaces TextDS::require
 * (ds/TextDS.hpp:247-292) and the provider constructors it calls: SADivSufSort.hpp:28-51, PhiFromSA.hpp:24-48,
 * PLCPFromPhi.hpp:27-53, LCPFromPLCP.hpp:27-54, ISAFromSA.hpp:24-46.  Dependencies are built implicitly. */
int tdcgpu_textds_build(tdcgpu_ctx* ctx, uint32_t flags);

/* Copy one built structure to a caller buffer of n uint32_t (n bytes for TDCGPU_BWT).  `which` is a single flag.
 * to_device != 0: dst is a device pointer.  The uint32_t layout is byte-identical to the reference's
 * DynamicIntVector at width 32 (ds/BitPackingVector.hpp:259-271). */
int tdcgpu_textds_get(tdcgpu_ctx* ctx, uint32_t which, void* dst, int to_device);

/* The same structure bit-packed on the device to `width` bits per element, in the layout of the reference's
 * DynamicIntVector / BitPackingVector (ds/BitPackingVector.hpp:62-98, 259-271): element i occupies bits [i*width,
 * (i+1)*width) of a little-endian stream of 64-bit words, values truncated to their low `width` bits
 * (sdsl::bits::write_int semantics).  This is what the providers' compress() leaves behind with compress=delayed|compressed
 * (bits_for(n) for SA/ISA/Phi, bits_for(max_lcp) for PLCP/LCP: ds/SADivSufSort.hpp:53-63, LCPFromPLCP.hpp:56-66); doing
 * it here replaces the serial re-pack of BitPackingVector::resize (ds/BitPackingVector.hpp:478-540) and shrinks the
 * device-to-host copy from 4n to n*width/8 bytes.  dst holds cap_words >= ceil(n*width/64) words; unused high bits of the
 * last word are 0.  1 <= width <= 64 (the values are 32-bit: widths above 32 widen, for wide-index builds of the caller).  Uses (and thereby invalidates) the context's scratch. */
int tdcgpu_textds_get_packed(tdcgpu_ctx* ctx, uint32_t which, uint32_t width, uint64_t* dst, uint64_t cap_words, int to_device);

/* Device pointer of a built structure (valid until the next set_text/destroy); NULL if not built. */
const void* tdcgpu_textds_device_ptr(tdcgpu_ctx* ctx, uint32_t which);

/* max over PLCP[0..n-2] (= max LCP) — PLCPFromPhi::max_lcp(), ds/PLCPFromPhi.hpp:40,55-57; needs PLCP or LCP built */
int tdcgpu_textds_max_lcp(tdcgpu_ctx* ctx, uint32_t* max_lcp);

/* Greedy lzss_lcp factorisation — the "Factorize" phase of LZSSLCPCompressor::compress
 * (compressors/LZSSLCPCompressor.hpp:60-115).  Needs SA, ISA, LCP (built on demand).  Results stay on the device;
 * *count = number of factors, *min_len / *max_len = FactorBuffer::shortest_factor()/longest_factor()
 * (lzss/LZSSFactors.hpp:33-47; 0xFFFFFFFF / 0 when there is no factor). */
int tdcgpu_lzss_lcp_factorize(tdcgpu_ctx* ctx, uint32_t threshold, uint64_t* count, uint32_t* min_len, uint32_t* max_len);

/* Copy the factor list (position order, FactorBuffer::is_sorted() holds) to a caller buffer of `cap` records. */
int tdcgpu_lzss_lcp_get_factors(tdcgpu_ctx* ctx, tdcgpu_factor* dst, uint64_t cap, int to_device);

/* ---- lzss::encode_text on the device (compressors/lzss/LZSSCoding.hpp:18-92) -------------------------------------
 * For coders that write integers as plain binary of bits_for(range) bits (the tdc::Encoder default, Coder.hpp:63-80)
 * and every literal as one fixed code word: BitCoder (coders/BitCoder.hpp) and HuffmanCoder
 * (coders/HuffmanCoder.hpp:519-570).  The factor list of the last tdcgpu_lzss_lcp_factorize call and the text stay on
 * the device; only the literal histogram comes back before, and the finished bit stream after.
 *
 * Histogram of the literals lzss::TextLiterals iterates (lzss/LZSSLiterals.hpp:10-56: every text position outside a
 * factor, incl. the final 0) — what huff::count_alphabet_literals counts (coders/HuffmanCoder.hpp:37-49) — and
 * fdist_max, the longest literal run (LZSSCoding.hpp:29-41). */
int tdcgpu_lzss_literal_histogram(tdcgpu_ctx* ctx, uint64_t hist[256], uint64_t* fdist_max);

/* Encode n, flen_min, flen_max, fdist_max and the factor/literal items exactly as lzss::encode_text does
 * (LZSSCoding.hpp:46-91), bits MSB first (io/BitOStream.hpp:79-102).  codes[c]/lens[c]: the coder's word for literal c
 * (low lens[c] bits of codes[c], lens[c] <= 64; BitCoder: c in 8 bits; HuffmanCoder: codewords[map_to_effective[c]],
 * coders/HuffmanCoder.hpp:309-322).  The coder's own header already occupies `lead_bits` (0..7) bits of the current
 * byte `lead_byte` (its high bits); the device stream starts with that partial byte so that it can be appended to the
 * header's whole bytes.  *nbits = stream length in bits incl. lead_bits. */
int tdcgpu_lzss_encode(tdcgpu_ctx* ctx, const uint64_t codes[256], const uint8_t lens[256], uint32_t lead_bits,
                       uint8_t lead_byte, uint64_t* nbits);

/* Copy the stream to `dst` (cap bytes).  finalize != 0 appends what BitOStream::~BitOStream writes
 * (io/BitOStream.hpp:53-64: the number of bits used in the last byte goes into its low 3 bits, or into an extra byte when
 * fewer than 3 bits are free).  *nbytes = bytes written (ceil(nbits / 8), + at most 1 when finalized). */
int tdcgpu_lzss_encode_get(tdcgpu_ctx* ctx, uint8_t* dst, uint64_t cap, int finalize, uint64_t* nbytes, int to_device);

/* The same stream in pieces: bytes [offset, offset + cap) to a HOST buffer, so that a large archive is drained through one
 * small (ideally pinned, tdcgpu_pinned_alloc) buffer straight into the caller's output stream instead of through an
 * archive-sized intermediate vector.  *total = length of the whole (finalized) stream, *written = bytes stored by this call
 * (0 once offset >= total). */
int tdcgpu_lzss_encode_get_chunk(tdcgpu_ctx* ctx, uint64_t offset, uint8_t* dst, uint64_t cap, int finalize, uint64_t* total,
                                 uint64_t* written);

/* Width of the text-length field at the head of the lzss archive: `coder.encode(n, len_r)`, compressors/lzss/LZSSCoding.hpp:47,
 * with len_r = TypeRange<len_t> (Range.hpp:95-99, :115) — 32 bits in the reference's default build, 64 in its wide-index
 * build (-DLEN_BITS=40: len_t becomes a 64-bit type, def.hpp:100-114).  A plugin compiled against a wide-index reference
 * passes 8 * sizeof(len_t); default 32.  Sticky for the context; affects tdcgpu_lzss_encode only. */
int tdcgpu_set_len_bits(tdcgpu_ctx* ctx, uint32_t len_field_bits);

/* Page-locked host memory for caller-side staging buffers (cudaMallocHost / cudaFreeHost); NULL on failure. */
void* tdcgpu_pinned_alloc(uint64_t bytes);
void tdcgpu_pinned_free(void* p);
/* Device of the calling thread for the context-free calls above (a block-mode worker pins its block buffers on the
 * device it was dealt, tudocomp_b200/plugin/tdc_block.cpp, instead of opening a context on device 0). */
int tdcgpu_set_device(int device);

/* Plain device memory on the context's device (cudaMalloc / cudaFree), for callers that keep a byte stream in HBM
 * between two stages (the GPU-aware chain of the plugin, tudocomp_gpu/GpuStreamStages.hpp, which replaces the host
 * buffer of tudocomp_driver/ChainCompressor.hpp:54-62).  NULL on failure. */
void* tdcgpu_device_alloc(tdcgpu_ctx* ctx, uint64_t bytes);
void tdcgpu_device_free(tdcgpu_ctx* ctx, void* p); /* ctx may be NULL */
/* Blocking copy on the context's stream; kind: 0 host -> device, 1 device -> host (pageable host memory is staged through
 * pinned buffers like every other transfer of the ABI), 2 device -> device. */
int tdcgpu_device_copy(tdcgpu_ctx* ctx, void* dst, const void* src, uint64_t bytes, int kind);

/* ---- byte-stream stages behind the BWT in `bwt:mtf:rle:encode(huff)` (BASELINE config 3) -------------------------
 * Stateless with respect to the text index: they only use the context's stream and a scratch buffer of their own.
 * on_device says where the caller's `in` / `out` buffers are (device pointers are on the context's device): */
#define TDCGPU_BUF_HOST 0       /* both host */
#define TDCGPU_BUF_DEVICE 1     /* both device */
#define TDCGPU_BUF_OUT_DEVICE 2 /* in host, out device (first stage of a device-resident chain) */
#define TDCGPU_BUF_IN_DEVICE 3  /* in device, out host (last stage) */
/*
 *
 * mtf_encode (compressors/MTFCompressor.hpp:46-56): out[i] = index of in[i] in the move-to-front table (initially
 * 0..255) before it is moved to the front; n bytes out. */
int tdcgpu_mtf_encode(tdcgpu_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, int on_device);

/* rle_encode (compressors/RunLengthEncoder.hpp:15-31, util/vbyte.hpp:27-37): a run of L >= 2 equal bytes c becomes
 * c c vbyte(L - 2 + offset), single bytes are copied; runs of bytes >= 0x80 are not merged, exactly like the reference, whose
 * run counter compares an int with a signed char (:24).  *out_n = bytes produced (<= cap, else TDCGPU_ERR_ARG; a device
 * output buffer must hold the worst case n * (1 + vbyte_len(offset + n)) + 16).  n < 2^32 - 16. */
int tdcgpu_rle_encode(tdcgpu_ctx* ctx, const uint8_t* in, uint64_t n, uint64_t offset, uint8_t* out, uint64_t cap, uint64_t* out_n,
                      int on_device);

/* LiteralEncoder<coder>::compress (compressors/LiteralEncoder.hpp:23-32) for BitCoder / HuffmanCoder, in three steps like
 * the lzss encoder above: begin = stage the input on the device and count its bytes (what huff::count_alphabet_literals
 * sees through ViewLiterals, coders/HuffmanCoder.hpp:37-49, Literal.hpp:55-70); the caller derives the coder's header and
 * code words (reference code); encode = every byte by its code word behind the header's `lead_bits` bits; get = the stream
 * (finalize: BitOStream's tail).  A device input must stay valid until tdcgpu_literal_encode returns. */
int tdcgpu_literal_encode_begin(tdcgpu_ctx* ctx, const uint8_t* in, uint64_t n, int on_device, uint64_t hist[256]);
int tdcgpu_literal_encode(tdcgpu_ctx* ctx, const uint64_t codes[256], const uint8_t lens[256], uint32_t lead_bits, uint8_t lead_byte,
                          uint64_t* nbits);
int tdcgpu_literal_encode_get(tdcgpu_ctx* ctx, uint8_t* dst, uint64_t cap, int finalize, uint64_t* nbytes, int to_device);
int tdcgpu_literal_encode_get_chunk(tdcgpu_ctx* ctx, uint64_t offset, uint8_t* dst, uint64_t cap, int finalize, uint64_t* total,
                                    uint64_t* written);

/* One-shot host-buffer convenience used by the C++ provider shims: text in, arrays out (NULL = not wanted).
 * Same semantics as constructing TextDS<>(env, view, flags) and reading the providers. */
int tdcgpu_textds_build_host(int device, const uint8_t* text, uint64_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp,
                             uint32_t* phi, uint32_t* plcp, uint32_t* max_lcp);

/* One-shot BWTCompressor::compress payload (compressors/BWTCompressor.hpp:29-47): n bytes out. */
int tdcgpu_bwt_host(int device, const uint8_t* text, uint64_t n, uint8_t* out);

/* Per-phase device times of the last build/factorize call, for StatPhase (tudocomp_stat/StatPhase.hpp:217-220).
 * Returns the number of phases; name/ms of phase i via the getters (NULL / <0 when out of range). */
int tdcgpu_phase_count(tdcgpu_ctx* ctx);
const char* tdcgpu_phase_name(tdcgpu_ctx* ctx, int i);
float tdcgpu_phase_ms(tdcgpu_ctx* ctx, int i);

/* Work-model counters of the last SA build: [0] doubling rounds incl. the initial sort, [1] sum of
 * active suffixes over rounds, [2] radix passes executed, [3] elements moved by those passes, [4] alphabet size,
 * [5] symbols per initial key, [6] LCP route of the last LCP build (1 = direct comparison in SA order, 2 = Phi/PLCP as
 * in the reference), [7] sum over rounds of active suffixes x known common prefix (the route's LCP-sum estimate). */
int tdcgpu_sa_stats(tdcgpu_ctx* ctx, uint64_t out[8]);

/* Record layout of the initial sort of the last SA build: [0] 1 = packed 64-bit records (key bits | suffix index, 16 B per
 * suffix and radix pass), 0 = (64-bit key, 32-bit suffix) pairs (24 B); [1] key bits sorted; [2] index bits inside a
 * packed record; [3] whole symbols the key decides. */
int tdcgpu_sa_layout(tdcgpu_ctx* ctx, uint64_t out[4]);

/* Block until all work queued on the context's stream is done. */
int tdcgpu_sync(tdcgpu_ctx* ctx);

/* ---- measurement hooks (no reference counterpart; the reference times phases with StatPhase on the host) ---- */
/* CUDA events on the context's stream: record into slot 0..7, elapsed time between two recorded slots. */
int tdcgpu_event_record(tdcgpu_ctx* ctx, int slot);
int tdcgpu_event_elapsed_ms(tdcgpu_ctx* ctx, int slot_a, int slot_b, float* ms);
/* Number of CUDA kernels this library has launched in this process. */
uint64_t tdcgpu_launch_count(void);
/* Optional per-kernel timing: every launch is bracketed by CUDA events on its stream and aggregated by kernel name
 * (launch count, device ms, algorithmic bytes where the host knows them). */
void tdcgpu_profile_enable(int on);
void tdcgpu_profile_reset(void);
int tdcgpu_profile_count(void);
int tdcgpu_profile_entry(int i, const char** name, uint64_t* launches, double* ms, double* bytes);

/* ---- device-side checkers (no reference counterpart in the product path; they assert what the reference's own tests
 * assert, test/ds_tests.cpp:71-112, and re-run the reference's decision rule, compressors/LZSSLCPCompressor.hpp:60-115) ----
 * Used by `bench.py --verify`, the GPU tests and the sharded path's verification; build/factorize never call them.
 * All pointers are DEVICE pointers on the context's device; the text must be readable 16 bytes past n (the contexts'
 * own text buffers are).  Arrays are the FULL arrays; a rank of a sharded run checks its own slot / position range.
 *
 * check_index: slots [slot_lo, slot_lo + slot_cnt).  out[0] = slots with SA[i] out of range or ISA[SA[i]] != i;
 * out[1] = slots violating the Burkhardt-Kaerkkaeinen order criterion (or SA[0] != n-1); out[2] = slots whose LCP differs
 * from the directly compared common prefix (d_lcp may be NULL: not checked); out[3] = 0. */
int tdcgpu_check_index(tdcgpu_ctx* ctx, const uint8_t* d_text, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                       const uint32_t* d_lcp, uint64_t slot_lo, uint64_t slot_cnt, uint64_t out[4]);
/* check_factors: the z factors (position order) are exactly the lzss_lcp(threshold) parse of the positions
 * [pos_lo, pos_lo + pos_cnt) they cover.  out[0] = malformed records (len < threshold, src >= pos, past the sentinel);
 * out[1] = order / overlap violations; out[2] = factor starts where the reference's PSV/NSV scan decides another
 * (src, len); out[3] = positions outside every factor where that scan finds a factor; out[4] = positions whose scan
 * exceeded 2^16 steps at a text position >= 2^20 (not decided; earlier positions are decided by comparing with all earlier
 * suffixes directly). */
int tdcgpu_check_factors(tdcgpu_ctx* ctx, const uint8_t* d_text, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                         const uint32_t* d_lcp, const tdcgpu_factor* d_factors, uint64_t z, uint32_t threshold,
                         uint64_t pos_lo, uint64_t pos_cnt, uint64_t out[5]);
/* device pointers of the context's resident text (padded) and factor list (NULL before factorize) */
const uint8_t* tdcgpu_text_device_ptr(tdcgpu_ctx* ctx);
const tdcgpu_factor* tdcgpu_factors_device_ptr(tdcgpu_ctx* ctx);

/* ---- one text sharded over several GPUs (no reference counterpart: tudocomp is single-threaded) -------------------
 * One rank = one process = one GPU; ranks exchange data with NCCL (all-to-all of rank buckets, see
 * tudocomp_b200/csrc/dist_textds.cu).  Every rank passes the SAME full text; results come back as shards:
 *   SA, LCP   slots   [slot_lo, slot_lo + slot_cnt)   (ranks in increasing order; concatenation = the full array)
 *   ISA       positions [pos_lo, pos_lo + pos_cnt)
 *   factors   the factors whose pos lies in the rank's position range, in position order
 * All functions are collective: every rank of the communicator must call them in the same order.
 * n < 2^32 - 1 (unsigned 32-bit indices); common prefixes must stay below 2^31. */
typedef struct tdcgpu_dist tdcgpu_dist;
/* rank 0 creates the id (ncclGetUniqueId) and hands the 128 bytes to the other ranks by any means */
int tdcgpu_dist_unique_id(uint8_t id[128]);
int tdcgpu_dist_create(int device, int rank, int nranks, const uint8_t id[128], tdcgpu_dist** out);
void tdcgpu_dist_destroy(tdcgpu_dist* h);
int tdcgpu_dist_set_text(tdcgpu_dist* h, const uint8_t* text, uint64_t n, int on_device);
/* flags: TDCGPU_SA | TDCGPU_ISA | TDCGPU_LCP */
int tdcgpu_dist_build(tdcgpu_dist* h, uint32_t flags);
/* out = { slot_lo, slot_cnt, pos_lo, pos_cnt } */
int tdcgpu_dist_shard_info(tdcgpu_dist* h, uint64_t out[4]);
int tdcgpu_dist_get(tdcgpu_dist* h, uint32_t which, void* dst, int to_device);
int tdcgpu_dist_max_lcp(tdcgpu_dist* h, uint32_t* max_lcp);
int tdcgpu_dist_lzss_lcp_factorize(tdcgpu_dist* h, uint32_t threshold, uint64_t* local_count, uint64_t* total_count,
                                   uint32_t* min_len, uint32_t* max_len);
int tdcgpu_dist_get_factors(tdcgpu_dist* h, tdcgpu_factor* dst, uint64_t cap, int to_device);
/* device pointers of the rank's resident (replicated, padded) text and of its factor shard */
const uint8_t* tdcgpu_dist_text_device_ptr(tdcgpu_dist* h);
const tdcgpu_factor* tdcgpu_dist_factors_device_ptr(tdcgpu_dist* h);
int tdcgpu_dist_sync(tdcgpu_dist* h);
int tdcgpu_dist_event_record(tdcgpu_dist* h, int slot);
int tdcgpu_dist_event_elapsed_ms(tdcgpu_dist* h, int slot_a, int slot_b, float* ms);
/* [0] rounds, [1] sum of active suffixes (this rank), [2] radix passes, [3] elements moved by them, [4] alphabet,
 * [5] symbols per initial key, [6] per-rank element capacity, [7] transport of the exchanges: 1 = pushed through peer
 * memory (CUDA IPC mappings of the peers' scratch arenas, NVLink), 0 = NCCL send/recv */
int tdcgpu_dist_stats(tdcgpu_dist* h, uint64_t out[8]);
int tdcgpu_dist_phase_count(tdcgpu_dist* h);
const char* tdcgpu_dist_phase_name(tdcgpu_dist* h, int i);
float tdcgpu_dist_phase_ms(tdcgpu_dist* h, int i);

#ifdef __cplusplus
}
#endif
#endif
