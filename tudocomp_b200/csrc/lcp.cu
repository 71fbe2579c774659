// Phi / PLCP / LCP / BWT from the suffix array (sm_100a), bit-exact with the reference's
//   PhiFromSA    /root/reference/include/tudocomp/ds/PhiFromSA.hpp:37-41     Phi[SA[i]] = SA[i-1], Phi[SA[0]] = SA[n-1]
//   PLCPFromPhi  /root/reference/include/tudocomp/ds/PLCPFromPhi.hpp:36-44   PLCP[i] = lcp(T[i..], T[Phi[i]..]) for i <= n-2;
//                                                                            PLCP[n-1] keeps Phi[n-1]; max over i <= n-2
//   LCPFromPLCP  /root/reference/include/tudocomp/ds/LCPFromPLCP.hpp:43-47   LCP[0] = 0, LCP[i] = PLCP[SA[i]]
//   bwt::bwt     /root/reference/include/tudocomp/ds/bwt.hpp:19-22           BWT[i] = SA[i] ? T[SA[i]-1] : T[n-1]
//
// The CPU Phi-algorithm carries l from i to i+1, which is sequential.  Here: position i is *reducible* when
// T[i-1] == T[Phi[i]-1]; then PLCP[i] = PLCP[i-1] - 1 exactly.  Only irreducible positions compare characters
// (8 bytes per step; long matches are continued by a whole warp), and reducible runs are filled by a max-scan that
// carries the index of the last irreducible position.
#include "lcp_kernels.cuh"

namespace tdc {

// needs d_sa; fills d_lcp and max_lcp without materialising Phi/PLCP
int build_lcp_direct(Ctx& c) {
    const u64 n = c.n;
    cudaStream_t st = c.stream;
    c.arena.reset();
    u32* queue = c.arena.take<u32>(n);
    if (!queue) { set_error("lcp: scratch arena too small"); return -2; }
    u32* d_qlen = c.d_scalars + 0;
    u32* d_max = c.d_scalars + 1;
    TDC_CUDA(cudaMemsetAsync(c.d_scalars, 0, 2 * sizeof(u32), st));
    if (c.sa_lcp_seeded) {
        if (c.sa_tail_span) TDC_LAUNCH(lcp_tail_fix_kernel, u32(div_up(u64(c.sa_tail_span) + 1, 64)), 64, 0, st, c.d_text, c.d_sa, c.d_isa, n, c.sa_tail_span, c.d_lcp);
        TDC_LAUNCH(lcp_fix_kernel, u32(div_up(n, 256 * LCPFIX_EPT)), 256, 0, st, c.d_text, c.d_sa, n, c.d_lcp, c.symbols_per_key, queue, d_qlen, d_max);
        prof_add_bytes("lcp_fix_kernel", double(n) * 4 + double(c.sa_first_residue) * 72);
    } else {
        TDC_LAUNCH(lcp_direct_kernel, u32(div_up(n, 256)), 256, 0, st, c.d_text, c.d_sa, n, c.d_lcp, queue, d_qlen, d_max);
        prof_add_bytes("lcp_direct_kernel", double(n) * 72);
    }
    TDC_LAUNCH(lcp_direct_long_kernel, u32(c.sm_count * 4), 256, 0, st, c.d_text, c.d_sa, c.d_lcp, queue, d_qlen, d_max);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    c.max_lcp = c.h_scalars[1];
    return 0;
}

int build_phi_bwt(Ctx& c, bool want_phi, bool want_bwt) {
    if (!want_phi && !want_bwt) return 0;
    const u64 n = c.n;
    c.arena.reset();
    u32* sc_idx[2] = {nullptr, nullptr};
    u32* sc_val[2] = {nullptr, nullptr};
    if (want_phi) {
        for (int q = 0; q < 2; q++) { sc_idx[q] = c.arena.take<u32>(n); sc_val[q] = c.arena.take<u32>(n); }
        if (!sc_idx[0] || !sc_idx[1] || !sc_val[0] || !sc_val[1]) { set_error("phi: scratch arena too small"); return -2; }
    }
    TDC_LAUNCH(phi_bwt_kernel, u32(div_up(n, 256)), 256, 0, c.stream, c.d_sa, c.d_text, n, sc_idx[0], sc_val[0],
               want_bwt ? c.d_bwt : nullptr);
    TDC_KCHECK();
    if (want_phi) TDC_TRY(partitioned_scatter(c.sortws, c.stream, sc_idx, sc_val, n, c.d_phi, n, true));
    return 0;
}

// needs d_sa and d_phi; fills d_plcp, max_lcp and (optionally) d_lcp
int build_plcp_lcp(Ctx& c, bool want_lcp) {
    const u64 n = c.n;
    cudaStream_t st = c.stream;
    c.arena.reset();
    u32* tmp = c.arena.take<u32>(n);
    u32* queue = c.arena.take<u32>(n);
    const u32 ntiles = u32(div_up(n, PF_TILE));
    u32* agg = c.arena.take<u32>(ntiles);
    if (!tmp || !queue || !agg) { set_error("plcp: scratch arena too small"); return -2; }
    u32* d_qlen = c.d_scalars + 0;
    u32* d_max = c.d_scalars + 1;
    TDC_CUDA(cudaMemsetAsync(c.d_scalars, 0, 3 * sizeof(u32), st));
    TDC_LAUNCH(plcp_irreducible_kernel, u32(div_up(n, 256)), 256, 0, st, c.d_text, c.d_phi, n, tmp, queue, d_qlen);
    TDC_LAUNCH(plcp_long_kernel, u32(c.sm_count * 4), 256, 0, st, c.d_text, c.d_phi, tmp, queue, d_qlen, c.d_scalars + 2);
    TDC_LAUNCH(plcp_fill_reduce_kernel, ntiles, PF_THREADS, 0, st, tmp, n, agg);
    TDC_LAUNCH(plcp_fill_scan_kernel, 1, 1024, 0, st, agg, ntiles);
    TDC_LAUNCH(plcp_fill_apply_kernel, ntiles, PF_THREADS, 0, st, tmp, c.d_phi, n, agg, c.d_plcp, d_max);
    if (want_lcp) TDC_LAUNCH(lcp_gather_kernel, u32(div_up(n, 256)), 256, 0, st, c.d_sa, c.d_plcp, n, c.d_lcp);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars, c.d_scalars, 3 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    if (c.h_scalars[2]) { set_error("plcp: common prefixes of 2^31 bytes or more are not supported"); return -5; }
    c.max_lcp = c.h_scalars[1];
    return 0;
}

}  // namespace tdc
