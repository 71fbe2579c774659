// Host side of the byte-stream stages that follow the BWT in `bwt:mtf:rle:encode(huff)` (kernels: stream_kernels.cuh).
#include "stream_kernels.cuh"

namespace tdc {

int stream_arena_reserve(Ctx& c, size_t bytes) {
    bytes = (bytes + 4095) & ~size_t(4095);
    if (c.stream_arena.cap < bytes) {
        if (c.stream_arena.base) TDC_CUDA(cudaFree(c.stream_arena.base));
        c.stream_arena = Arena();
        TDC_CUDA(cudaMalloc(&c.stream_arena.base, bytes));
        c.stream_arena.cap = bytes;
    }
    c.stream_arena.reset();
    return 0;
}

size_t mtf_scratch_bytes(u64 n) {
    const u64 ntiles = div_up(n, MTF_TILE);
    return size_t(ntiles) * (256 + 256 + 4) + 3 * 256;
}

// mtf_encode of /root/reference/include/tudocomp/compressors/MTFCompressor.hpp:46-56 over n bytes
int mtf_encode_device(Ctx& c, const uint8_t* d_in, u64 n, uint8_t* d_out) {
    if (n == 0) return 0;
    cudaStream_t st = c.stream;
    const u32 ntiles = u32(div_up(n, MTF_TILE));
    uint8_t* tile_R = c.stream_arena.take<uint8_t>(size_t(ntiles) * 256);
    uint8_t* incoming = c.stream_arena.take<uint8_t>(size_t(ntiles) * 256);
    u32* tile_rc = c.stream_arena.take<u32>(ntiles);
    if (!tile_R || !incoming || !tile_rc) { set_error("mtf: stream scratch too small"); return -2; }
    auto mtf_summaries = mtf_tile_kernel<false>;
    auto mtf_apply = mtf_tile_kernel<true>;
    TDC_CUDA(cudaFuncSetAttribute(mtf_summaries, cudaFuncAttributeMaxDynamicSharedMemorySize, int(mtf_smem_bytes())));
    TDC_CUDA(cudaFuncSetAttribute(mtf_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, int(mtf_smem_bytes())));
    const bool vec16 = ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) == 0;
    TDC_LAUNCH(mtf_summaries, ntiles, MTF_THREADS, mtf_smem_bytes(), st, d_in, n, tile_R, tile_rc, (const uint8_t*)nullptr, (uint8_t*)nullptr, vec16);
    prof_add_bytes("mtf_summaries", double(n));
    TDC_LAUNCH(mtf_scan_kernel, 1, 32, 0, st, tile_R, tile_rc, ntiles, incoming);
    TDC_LAUNCH(mtf_apply, ntiles, MTF_THREADS, mtf_smem_bytes(), st, d_in, n, (uint8_t*)nullptr, (u32*)nullptr, incoming, d_out, vec16);
    prof_add_bytes("mtf_apply", double(n) * 3);
    TDC_KCHECK();
    return 0;
}

size_t rle_scratch_bytes(u64 n) {
    const u64 ntiles = div_up(n, RLE_TILE);
    return size_t(ntiles) * (4 + 4 + 4 + 8) + 5 * 256 + 64;
}

// rle_encode of /root/reference/include/tudocomp/compressors/RunLengthEncoder.hpp:15-31.  d_out must hold
// rle_max_output(n, offset) bytes; *out_n = bytes produced.
int rle_encode_device(Ctx& c, const uint8_t* d_in, u64 n, u64 offset, uint8_t* d_out, u64* out_n) {
    *out_n = 0;
    if (n == 0) return 0;
    if (n >= 0xfffffff0ull) { set_error("rle: n must be < 2^32 - 16"); return -5; }
    cudaStream_t st = c.stream;
    const u32 ntiles = u32(div_up(n, RLE_TILE));
    u32* tile_first = c.stream_arena.take<u32>(ntiles);
    u32* next_after = c.stream_arena.take<u32>(ntiles);
    u32* tile_bytes = c.stream_arena.take<u32>(ntiles);
    u64* tile_off = c.stream_arena.take<u64>(ntiles);
    u64* d_total = c.stream_arena.take<u64>(1);
    if (!tile_first || !next_after || !tile_bytes || !tile_off || !d_total) { set_error("rle: stream scratch too small"); return -2; }
    TDC_LAUNCH(rle_first_head_kernel, ntiles, RLE_THREADS, 0, st, d_in, n, tile_first);
    TDC_LAUNCH(rle_next_head_kernel, 1, 1024, 0, st, tile_first, ntiles, u32(n), next_after);
    auto rle_count = rle_tile_kernel<1>;
    auto rle_write = rle_tile_kernel<2>;
    TDC_LAUNCH(rle_count, ntiles, RLE_THREADS, 0, st, d_in, n, offset, next_after, tile_bytes, (const u64*)nullptr, (uint8_t*)nullptr);
    auto rle_scan = enc_scan_kernel<u64>;
    TDC_LAUNCH(rle_scan, 1, 1024, 0, st, tile_bytes, ntiles, u64(0), tile_off, d_total);
    TDC_LAUNCH(rle_write, ntiles, RLE_THREADS, 0, st, d_in, n, offset, next_after, (u32*)nullptr, tile_off, d_out);
    prof_add_bytes("rle_write", double(n) * 2);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 16, d_total, sizeof(u64), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    memcpy(out_n, c.h_scalars + 16, sizeof(u64));
    return 0;
}

u64 rle_max_output(u64 n, u64 offset) {
    // worst case: one long run of a byte >= 0x80, every byte followed by vbyte(offset)
    u32 vl = 1;
    for (u64 v = offset + n; v >>= 7;) vl++;
    return n * (1 + vl) + 16;
}

// ---- LiteralEncoder (LiteralEncoder.hpp:23-32) ----
size_t literal_scratch_bytes(u64 n) {
    const u64 ntiles = div_up(n, ENC_TILE);
    return size_t(ntiles) * (4 + 8) + 256 * 8 + 256 + 256 * 8 + 6 * 256;
}

int stream_histogram_device(Ctx& c, const uint8_t* d_in, u64 n, u64 hist[256]) {
    cudaStream_t st = c.stream;
    ull* d_hist = c.stream_arena.take<ull>(256);
    if (!d_hist) { set_error("histogram: stream scratch too small"); return -2; }
    TDC_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(ull) * 256, st));
    if (n) {
        const u32 grid = u32(min(u64(c.sm_count) * 8, div_up(n, 256 * 16)));
        TDC_LAUNCH(stream_histogram_kernel, grid, 256, 0, st, d_in, n, d_hist);
        prof_add_bytes("stream_histogram_kernel", double(n));
        TDC_KCHECK();
    }
    TDC_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof(u64) * 256, cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// The stream goes to c.lit.d_out (grow-only allocation of its own).  *nbits = stream length incl. lead_bits.
int literal_encode_device(Ctx& c, const uint8_t* d_in, u64 n, const u64* codes, const uint8_t* lens, u32 lead_bits, u32 lead_byte,
                          u64* nbits) {
    cudaStream_t st = c.stream;
    if (lead_bits > 7) { set_error("literal encode: lead_bits must be 0..7"); return -5; }
    for (int i = 0; i < 256; i++)
        if (lens[i] > 64) { set_error("literal encode: bad code length %u for literal %d", unsigned(lens[i]), i); return -5; }
    const u32 ntiles = u32(div_up(n, ENC_TILE));
    u64* d_code = c.stream_arena.take<u64>(256);
    uint8_t* d_len = c.stream_arena.take<uint8_t>(256);
    u32* tile_bits = c.stream_arena.take<u32>(ntiles + 1);
    u64* tile_off = c.stream_arena.take<u64>(ntiles + 1);
    u64* d_total = c.stream_arena.take<u64>(1);
    if (!d_code || !d_len || !tile_bits || !tile_off || !d_total) { set_error("literal encode: stream scratch too small"); return -2; }
    TDC_CUDA(cudaMemcpyAsync(d_code, codes, sizeof(u64) * 256, cudaMemcpyHostToDevice, st));
    TDC_CUDA(cudaMemcpyAsync(d_len, lens, 256, cudaMemcpyHostToDevice, st));
    auto lit_count = lit_tile_kernel<1>;
    auto lit_write = lit_tile_kernel<2>;
    auto lit_scan = enc_scan_kernel<u64>;
    if (ntiles) TDC_LAUNCH(lit_count, ntiles, ENC_THREADS, 0, st, d_in, n, d_code, d_len, tile_bits, (const u64*)nullptr, (u32*)nullptr);
    TDC_LAUNCH(lit_scan, 1, 1024, 0, st, tile_bits, ntiles, u64(lead_bits), tile_off, d_total);
    TDC_KCHECK();
    TDC_CUDA(cudaMemcpyAsync(c.h_scalars + 16, d_total, sizeof(u64), cudaMemcpyDeviceToHost, st));
    TDC_CUDA(cudaStreamSynchronize(st));
    memcpy(nbits, c.h_scalars + 16, sizeof(u64));
    const u64 need = ((*nbits + 31) / 32) * 4 + 8;
    if (need > c.lit.out_cap) {
        if (c.lit.d_out) TDC_CUDA(cudaFree(c.lit.d_out));
        c.lit.d_out = nullptr;
        c.lit.out_cap = 0;
        TDC_CUDA(cudaMalloc(&c.lit.d_out, need + need / 8));
        c.lit.out_cap = need + need / 8;
    }
    uint8_t* d_out = c.lit.d_out;
    TDC_CUDA(cudaMemsetAsync(d_out, 0, need, st));
    u32* out32 = reinterpret_cast<u32*>(d_out);
    TDC_LAUNCH(lit_header_kernel, 1, 32, 0, st, lead_bits, lead_byte, out32);
    if (ntiles) TDC_LAUNCH(lit_write, ntiles, ENC_THREADS, 0, st, d_in, n, d_code, d_len, (u32*)nullptr, tile_off, out32);
    prof_add_bytes("lit_write", double(n) + double(*nbits) / 8);
    TDC_KCHECK();
    return 0;
}

}  // namespace tdc
