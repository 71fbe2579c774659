// Device-side CHECKERS for results too large for the CPU oracle (4 GB sharded texts, every step of a 1 GiB bench).
// They are not part of the product path: nothing in build / factorize calls them.  `bench.py --verify`, the GPU tests
// and tudocomp_b200.dist use them to check full arrays instead of a few thousand sampled slots.
//
//   index   SA is a permutation with ISA its inverse (ISA[SA[i]] == i for every slot, SA[0] == n-1);
//           suffix order by the Burkhardt-Kaerkkaeinen criterion: for i >= 1
//               (T[SA[i-1]], ISA[SA[i-1]+1]) < (T[SA[i]], ISA[SA[i]+1])            [SURVEY.md Appendix B]
//           LCP[0] == 0 and LCP[i] == lcp(T[SA[i-1]..], T[SA[i]..]) by direct comparison, 8 bytes per step
//           (what ds_tests.cpp:71-112 of the reference asserts with its own loops)
//   factors every position of the greedy parse is re-decided with the REFERENCE's naive scans
//           (compressors/LZSSLCPCompressor.hpp:60-115: walk up to the PSV, walk down to the NSV, minima of LCP on the
//           way, PSV wins ties) over the checked SA / ISA / LCP: a factor start must reproduce (src, len), a position
//           outside every factor must stay below the threshold.  With the chain starting at 0 this is equivalent to
//           running the reference loop, but every position is independent.
// Arrays are addressed by device pointers, so a sharded result can be checked by the rank that owns a slice against
// gathered arrays.  u64 counters, one per kind of violation; all zero == verified.
#include "../../include/tdcgpu.h"
#include "tdc_ctx.h"

namespace tdc {

#ifdef TDC_CUSIM
static const u32 CHECK_SCAN_CAP = 64;         // tiny, so that the CPU tests exercise the early-position route
#else
static const u32 CHECK_SCAN_CAP = 1u << 16;   // steps of one naive PSV / NSV scan before the position takes the other route
#endif
static const u64 CHECK_EARLY_POS = u64(1) << 20;  // ... which is affordable for text positions below this bound

static __global__ void __launch_bounds__(256)
check_index_kernel(const uint8_t* __restrict__ text, u64 n, const u32* __restrict__ sa, const u32* __restrict__ isa,
                   const u32* __restrict__ lcp, u64 slot_lo, u64 slot_cnt, ull* __restrict__ bad) {
    const u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= slot_cnt) return;
    const u64 i = slot_lo + t;
    const u64 b = sa[i];
    if (b >= n || u64(isa[b]) != i) { atomicAdd(&bad[0], 1ull); return; }
    if (i == 0) {
        if (b != n - 1) atomicAdd(&bad[1], 1ull);
        if (lcp && lcp[0] != 0) atomicAdd(&bad[2], 1ull);
        return;
    }
    const u64 a = sa[i - 1];
    if (a >= n) { atomicAdd(&bad[0], 1ull); return; }
    const uint8_t ca = text[a], cb = text[b];
    bool less = ca < cb;
    if (ca == cb) less = (a + 1 < n && b + 1 < n) ? (isa[a + 1] < isa[b + 1]) : false;
    if (!less) atomicAdd(&bad[1], 1ull);
    if (lcp) {
        // common prefix by direct comparison; the unique 0 at n-1 ends every comparison inside the text (+ padding)
        u64 l = 0;
        const u64 want = lcp[i];
        while (true) {
            const u64 x = load_text8(text, a + l) ^ load_text8(text, b + l);
            if (x) { l += u64(__ffsll((long long)x) - 1) >> 3; break; }
            l += 8;
            if (l > want + 8) break;  // already longer than claimed: a mismatch either way
        }
        if (l != want) atomicAdd(&bad[2], 1ull);
    }
}

// the reference's decision at text position i: returns len (0 if below nothing), *src = source position
static __device__ __forceinline__ u32 naive_decision(const u32* __restrict__ sa, const u32* __restrict__ isa, const u32* __restrict__ lcp,
                                                     u64 n, u64 i, u32 stop_below, u32* src, bool* undecided) {
    const u64 cur = isa[i];
    // upwards: include current, exclude last (LZSSLCPCompressor.hpp:68-77)
    u64 psv_lcp = lcp[cur];
    long long psv_pos = (long long)cur - 1;
    u32 steps = 0;
    if (psv_lcp > 0) {
        while (psv_pos >= 0 && u64(sa[psv_pos]) > i) {
            psv_lcp = min(psv_lcp, u64(lcp[psv_pos]));
            psv_pos--;
            if (psv_lcp < stop_below) break;  // the minimum only falls: this side cannot reach the threshold any more
            if (++steps > CHECK_SCAN_CAP) { *undecided = true; break; }
        }
    }
    if (psv_pos < 0) psv_lcp = min(psv_lcp, u64(0));  // walked off the top: the loop's last min was lcp[0] == 0
    // downwards: exclude current, include last (:82-96)
    u64 nsv_lcp = 0, nsv_pos = cur + 1;
    if (nsv_pos < n) {
        nsv_lcp = ~u64(0);
        steps = 0;
        do {
            nsv_lcp = min(nsv_lcp, u64(lcp[nsv_pos]));
            if (u64(sa[nsv_pos]) < i) break;
            if (nsv_lcp < stop_below) { nsv_pos = n; break; }
            if (++steps > CHECK_SCAN_CAP) { *undecided = true; nsv_pos = n; break; }
        } while (++nsv_pos < n);
        if (nsv_pos >= n) nsv_lcp = 0;
    }
    const u64 mx = max(psv_lcp, nsv_lcp);
    if (mx == 0) { *src = 0; return 0; }
    const u64 at = mx == psv_lcp ? u64(psv_pos) : nsv_pos;
    *src = (at < n) ? sa[at] : 0u;
    return u32(mx);
}

// Longest common prefix of two suffixes by direct comparison (8 bytes per step; the unique 0 at n-1 ends it).
static __device__ __forceinline__ u64 direct_lce(const uint8_t* __restrict__ text, u64 a, u64 b) {
    u64 l = 0;
    while (true) {
        const u64 x = load_text8(text, a + l) ^ load_text8(text, b + l);
        if (x) return l + (u64(__ffsll((long long)x) - 1) >> 3);
        l += 8;
    }
}

// The same decision for an EARLY text position i (few suffixes start before it, so the naive scans run for ~n/i steps):
// PSV / NSV of rank ISA[i] among the suffixes 0..i-1 by looking at all of them, the two lengths by direct comparison
// (equal to the LCP range minima the reference takes, given the checked LCP array).  O(i) instead of O(n/i).
static __device__ u32 early_decision(const uint8_t* __restrict__ text, const u32* __restrict__ isa, u64 i, u32* src) {
    const u32 cur = isa[i];
    long long best_up = -1, best_dn = -1;  // text positions of the nearest smaller / larger rank among j < i
    u32 r_up = 0, r_dn = 0xffffffffu;
    for (u64 j = 0; j < i; j++) {
        const u32 r = isa[j];
        if (r < cur) { if (best_up < 0 || r > r_up) { r_up = r; best_up = (long long)j; } }
        else { if (best_dn < 0 || r < r_dn) { r_dn = r; best_dn = (long long)j; } }
    }
    const u64 l_up = best_up >= 0 ? direct_lce(text, u64(best_up), i) : 0;
    const u64 l_dn = best_dn >= 0 ? direct_lce(text, u64(best_dn), i) : 0;
    const u64 mx = max(l_up, l_dn);
    if (mx == 0) { *src = 0; return 0; }
    *src = u32(mx == l_up ? best_up : best_dn);  // PSV wins ties (LZSSLCPCompressor.hpp:101)
    return u32(mx);
}

// one thread per factor: copy check, order / overlap with the predecessor, and the reference's decision at its start
static __global__ void __launch_bounds__(128)
check_factor_starts_kernel(const uint8_t* __restrict__ text, u64 n, const u32* __restrict__ sa, const u32* __restrict__ isa,
                           const u32* __restrict__ lcp, const Factor* __restrict__ f, u64 z, u32 threshold, ull* __restrict__ bad) {
    const u64 k = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= z) return;
    const Factor me = f[k];
    if (me.len < threshold || me.src >= me.pos || u64(me.pos) + me.len > n - 1) { atomicAdd(&bad[0], 1ull); return; }
    if (k > 0) {
        const Factor pv = f[k - 1];
        if (u64(pv.pos) + pv.len > me.pos) atomicAdd(&bad[1], 1ull);
    }
    u32 src = 0;
    bool undecided = false;
    u32 len = naive_decision(sa, isa, lcp, n, me.pos, 0u, &src, &undecided);
    if (undecided) {
        if (me.pos >= CHECK_EARLY_POS) { atomicAdd(&bad[4], 1ull); return; }
        len = early_decision(text, isa, me.pos, &src);
    }
    if (len != me.len || src != me.src) atomicAdd(&bad[2], 1ull);
}

// one thread per text position of [pos_lo, pos_lo + pos_cnt): positions outside every factor (and before n-1) must not
// admit a factor.  The covering factor is found by binary search over the position-sorted list.
static __global__ void __launch_bounds__(256)
check_literals_kernel(const uint8_t* __restrict__ text, u64 n, const u32* __restrict__ sa, const u32* __restrict__ isa, const u32* __restrict__ lcp,
                      const Factor* __restrict__ f, u64 z, u32 threshold, u64 pos_lo, u64 pos_cnt, ull* __restrict__ bad) {
    const u64 t = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= pos_cnt) return;
    const u64 i = pos_lo + t;
    if (i + 1 >= n) return;  // the sentinel is never examined (loop bound i + 1 < n)
    u64 lo = 0, hi = z;      // last factor with pos <= i
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (u64(f[mid].pos) <= i) lo = mid + 1; else hi = mid;
    }
    if (lo > 0) {
        const Factor c = f[lo - 1];
        if (i < u64(c.pos) + c.len) return;  // inside a factor: not a chain position
    }
    u32 src = 0;
    bool undecided = false;
    u32 len = naive_decision(sa, isa, lcp, n, i, threshold, &src, &undecided);
    if (undecided) {
        if (i >= CHECK_EARLY_POS) { atomicAdd(&bad[4], 1ull); return; }
        len = early_decision(text, isa, i, &src);
    }
    if (len >= threshold) atomicAdd(&bad[3], 1ull);
}

}  // namespace tdc

using namespace tdc;

extern "C" {

int tdcgpu_check_index(tdcgpu_ctx* ctx, const uint8_t* d_text, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                       const uint32_t* d_lcp, uint64_t slot_lo, uint64_t slot_cnt, uint64_t out[4]) {
    if (!ctx || !d_text || !d_sa || !d_isa || !out) { set_error("check_index: null argument"); return TDCGPU_ERR_ARG; }
    Ctx& c = ctx->c;
    TDC_CUDA(cudaSetDevice(c.device));
    if (slot_lo + slot_cnt > n) { set_error("check_index: slot range out of bounds"); return TDCGPU_ERR_ARG; }
    ull* d_bad = reinterpret_cast<ull*>(c.d_scalars + 300);
    TDC_CUDA(cudaMemsetAsync(d_bad, 0, 8 * sizeof(ull), c.stream));
    if (slot_cnt) {
        TDC_LAUNCH(check_index_kernel, u32(div_up(slot_cnt, 256)), 256, 0, c.stream, d_text, n, d_sa, d_isa, d_lcp, slot_lo, slot_cnt, d_bad);
        TDC_KCHECK();
    }
    ull h[8];
    TDC_CUDA(cudaMemcpyAsync(h, d_bad, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 4; i++) out[i] = h[i];
    return 0;
}

int tdcgpu_check_factors(tdcgpu_ctx* ctx, const uint8_t* d_text, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                         const uint32_t* d_lcp, const tdcgpu_factor* d_factors, uint64_t z, uint32_t threshold,
                         uint64_t pos_lo, uint64_t pos_cnt, uint64_t out[5]) {
    if (!ctx || !d_text || !d_sa || !d_isa || !d_lcp || !out || (z && !d_factors)) { set_error("check_factors: null argument"); return TDCGPU_ERR_ARG; }
    if (threshold < 1) { set_error("check_factors: threshold must be >= 1"); return TDCGPU_ERR_ARG; }
    Ctx& c = ctx->c;
    TDC_CUDA(cudaSetDevice(c.device));
    if (pos_lo + pos_cnt > n) { set_error("check_factors: position range out of bounds"); return TDCGPU_ERR_ARG; }
    ull* d_bad = reinterpret_cast<ull*>(c.d_scalars + 300);
    TDC_CUDA(cudaMemsetAsync(d_bad, 0, 8 * sizeof(ull), c.stream));
    const Factor* f = reinterpret_cast<const Factor*>(d_factors);
    if (z) TDC_LAUNCH(check_factor_starts_kernel, u32(div_up(z, 128)), 128, 0, c.stream, d_text, n, d_sa, d_isa, d_lcp, f, z, threshold, d_bad);
    if (pos_cnt) TDC_LAUNCH(check_literals_kernel, u32(div_up(pos_cnt, 256)), 256, 0, c.stream, d_text, n, d_sa, d_isa, d_lcp, f, z, threshold, pos_lo, pos_cnt, d_bad);
    TDC_KCHECK();
    ull h[8];
    TDC_CUDA(cudaMemcpyAsync(h, d_bad, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 5; i++) out[i] = h[i];
    return 0;
}

const uint8_t* tdcgpu_text_device_ptr(tdcgpu_ctx* ctx) { return ctx ? ctx->c.d_text : nullptr; }
const tdcgpu_factor* tdcgpu_factors_device_ptr(tdcgpu_ctx* ctx) {
    return (ctx && ctx->c.have_factors) ? reinterpret_cast<const tdcgpu_factor*>(ctx->c.d_factors) : nullptr;
}

}  // extern "C"
