// lzss::encode_text on the device (kernels and stream layout: encode_kernels.cuh; reference:
// /root/reference/include/tudocomp/compressors/lzss/LZSSCoding.hpp:18-92).  Input: the factor records left on the
// device by factorize_lzss_lcp and the resident text.  The coder's own header (and, for Huffman, the code table built
// by the reference's huff::gen_huffmantable from the literal histogram computed here) stays on the host.
#include "encode_kernels.cuh"

namespace tdc {

int encode_prepare(Ctx& c) {
    if (!c.have_factors) { set_error("lzss encode: no factor list (call tdcgpu_lzss_lcp_factorize first)"); return -6; }
    EncodeState& e = c.enc;
    if (e.prepared && e.gen == c.arena.gen) return 0;
    cudaStream_t st = c.stream;
    const u32 n = u32(c.n), z = u32(c.num_factors);
    c.arena.reset();
    e.gen = c.arena.gen;
    e.prepared = e.encoded = false;
    e.out = nullptr;
    e.out_cap = 0;
    e.ntiles = u32(div_up(u64(n), ENC_TILE));
    const u64 mask_words = u64(e.ntiles) * ENC_MASK_WORDS;
    e.S = c.arena.take<u32>(mask_words);
    e.E = c.arena.take<u32>(mask_words);
    e.scan_s = c.arena.take<u32>(e.ntiles);
    e.scan_e = c.arena.take<u32>(e.ntiles);
    e.tile_bits = c.arena.take<u32>(e.ntiles);
    e.tile_off = c.arena.take<u64>(e.ntiles);
    e.d_code = c.arena.take<u64>(256);
    e.d_len = c.arena.take<uint8_t>(256);
    ull* d_hist = c.arena.take<ull>(256);
    if (!e.S || !e.E || !e.scan_s || !e.scan_e || !e.tile_bits || !e.tile_off || !e.d_code || !e.d_len || !d_hist) {
        set_error("lzss encode: scratch arena too small");
        return -2;
    }
    u32* d_fdist = c.d_scalars + 8;
    u32* d_tot = c.d_scalars + 10;
    TDC_CUDA(cudaMemsetAsync(e.S, 0, sizeof(u32) * mask_words, st));
    TDC_CUDA(cudaMemsetAsync(e.E, 0, sizeof(u32) * mask_words, st));
    TDC_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(ull) * 256, st));
    TDC_CUDA(cudaMemsetAsync(d_fdist, 0, sizeof(u32), st));
    if (z) {
        TDC_LAUNCH(enc_mark_kernel, u32(div_up(u64(z), 256)), 256, 0, st, c.d_factors, z, n, e.S, e.E, d_fdist);
        prof_add_bytes("enc_mark_kernel", double(z) * 12);
    }
    TDC_LAUNCH(enc_tile_popc_kernel, u32(div_up(u64(e.ntiles), 8)), 256, 0, st, e.S, e.E, e.ntiles, e.scan_s, e.scan_e);
    auto enc_scan_u32 = enc_scan_kernel<u32>;
    TDC_LAUNCH(enc_scan_u32, 1, 1024, 0, st, e.scan_s, e.ntiles, 0u, e.scan_s, d_tot);
    TDC_LAUNCH(enc_scan_u32, 1, 1024, 0, st, e.scan_e, e.ntiles, 0u, e.scan_e, d_tot + 1);
    EncParams P{n, z, 1, 1, 1, c.flen_min};
    auto enc_hist = enc_tile_kernel<0>;
    TDC_LAUNCH(enc_hist, e.ntiles, ENC_THREADS, 0, st, c.d_text, c.d_factors, e.S, e.E, e.scan_s, e.scan_e, P,
               (const u64*)nullptr, (const uint8_t*)nullptr, d_hist, (u32*)nullptr, (const u64*)nullptr, (u32*)nullptr);
    prof_add_bytes("enc_hist", double(n) * 1.25);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(e.hist, d_hist, sizeof(u64) * 256, cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 8, d_fdist, sizeof(u32), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    // no factor: the single literal run spans the whole text (LZSSCoding.hpp:40)
    e.fdist_max = z ? c.h_scalars[8] : n;
    e.prepared = true;
    return 0;
}

int encode_lzss(Ctx& c, const u64* codes, const uint8_t* lens, u32 lead_bits, u32 lead_byte) {
    TDC_TRY(encode_prepare(c));
    EncodeState& e = c.enc;
    cudaStream_t st = c.stream;
    const u32 n = u32(c.n), z = u32(c.num_factors);
    if (lead_bits > 7) { set_error("lzss encode: lead_bits must be 0..7"); return -5; }
    for (int i = 0; i < 256; i++)
        if (lens[i] > 64 || (e.hist[i] && lens[i] == 0)) { set_error("lzss encode: bad code length %u for literal %d", unsigned(lens[i]), i); return -5; }
    EncParams P;
    P.n = n;
    P.z = z;
    P.bn = bits_for_host(n);
    P.bf = bits_for_host(e.fdist_max);
    P.bl = z ? bits_for_host(u64(c.flen_max) - u64(c.flen_min)) : 1;  // unused without factors
    P.flen_min = c.flen_min;
    TDC_CUDA(cudaMemcpyAsync(e.d_code, codes, sizeof(u64) * 256, cudaMemcpyHostToDevice, st));
    TDC_CUDA(cudaMemcpyAsync(e.d_len, lens, 256, cudaMemcpyHostToDevice, st));
    auto enc_count = enc_tile_kernel<1>;
    TDC_LAUNCH(enc_count, e.ntiles, ENC_THREADS, 0, st, c.d_text, c.d_factors, e.S, e.E, e.scan_s, e.scan_e, P, e.d_code, e.d_len,
               (ull*)nullptr, e.tile_bits, (const u64*)nullptr, (u32*)nullptr);
    prof_add_bytes("enc_count", double(n) * 1.25);
    const u64 head_bits = u64(lead_bits) + c.len_field_bits + 3 * u64(P.bn);
    u64* d_total = reinterpret_cast<u64*>(c.d_scalars + 12);  // 8-byte aligned
    auto enc_scan_u64 = enc_scan_kernel<u64>;
    TDC_LAUNCH(enc_scan_u64, 1, 1024, 0, st, e.tile_bits, e.ntiles, head_bits, e.tile_off, d_total);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 12, d_total, sizeof(u64), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    u64 nbits;
    memcpy(&nbits, c.h_scalars + 12, sizeof(u64));
    // output buffer: carved behind the prepared state; a second encode (other code table) reuses the same place
    const u64 out_bytes = ((nbits + 31) / 32) * 4 + 8;
    if (!e.out || e.out_cap < out_bytes) {
        e.out = c.arena.take<uint8_t>(out_bytes);
        e.out_cap = e.out ? out_bytes : 0;
        if (!e.out) { set_error("lzss encode: scratch arena too small for %llu output bytes", (unsigned long long)out_bytes); return -2; }
    }
    TDC_CUDA(cudaMemsetAsync(e.out, 0, out_bytes, st));
    u32* out32 = reinterpret_cast<u32*>(e.out);
    TDC_LAUNCH(enc_header_kernel, 1, 32, 0, st, lead_bits, lead_byte, P, c.flen_max, e.fdist_max, c.len_field_bits, out32);
    auto enc_write = enc_tile_kernel<2>;
    TDC_LAUNCH(enc_write, e.ntiles, ENC_THREADS, 0, st, c.d_text, c.d_factors, e.S, e.E, e.scan_s, e.scan_e, P, e.d_code, e.d_len,
               (ull*)nullptr, (u32*)nullptr, e.tile_off, out32);
    prof_add_bytes("enc_write", double(n) * 1.25 + double(z) * 12 + double(nbits) / 8);
    TDC_KCHECK();
    e.nbits = nbits;
    e.encoded = true;
    return 0;
}

}  // namespace tdc
