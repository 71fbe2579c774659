// Host <-> device transfers of CALLER buffers (the text in, arrays / factors / archives out).
//
// The reference hands the path a View over pageable memory (a private file mapping or an anonymous copy,
// /root/reference/include/tudocomp/io/RestrictedBuffer.hpp:108-181) and receives results in std::vector / DynamicIntVector
// storage, also pageable.  A plain cudaMemcpy on such memory is staged by the driver through one small bounce buffer by
// one thread.  Here: pinned or registered memory goes directly; pageable memory is cut into chunks that HC_THREADS copy
// threads move through their own pinned double buffers and copy streams, so host memcpy and DMA of different chunks
// overlap and the transfer approaches the PCIe rate.
#include <cstring>
#include <thread>

#include "tdc_ctx.h"

namespace tdc {

void host_copier_free(HostCopier& hc) {
    for (int t = 0; t < HC_THREADS; t++) {
        for (int b = 0; b < 2; b++) {
            if (hc.buf[t][b]) cudaFreeHost(hc.buf[t][b]);
            if (hc.ev[t][b]) cudaEventDestroy(hc.ev[t][b]);
            hc.buf[t][b] = nullptr;
            hc.ev[t][b] = nullptr;
        }
        if (hc.stream[t]) cudaStreamDestroy(hc.stream[t]);
        hc.stream[t] = cudaStream_t(0);
    }
    hc.ready = false;
}

static int host_copier_init(HostCopier& hc) {
    if (hc.ready) return 0;
    for (int t = 0; t < HC_THREADS; t++) {
        TDC_CUDA(cudaStreamCreateWithFlags(&hc.stream[t], cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            TDC_CUDA(cudaMallocHost(&hc.buf[t][b], HC_CHUNK));
            TDC_CUDA(cudaEventCreateWithFlags(&hc.ev[t][b], cudaEventDisableTiming));
        }
    }
    hc.ready = true;
    return 0;
}

static bool is_pageable(const void* p) {
#ifdef TDC_CUSIM
    (void)p;
    return false;
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
#endif
}

// One copy thread: chunks t, t + T, t + 2T, ...  Returns the first CUDA error (as int) or 0.
static void copy_worker(HostCopier* hc, int device, int t, int T, uint8_t* dst, const uint8_t* src, size_t bytes, bool h2d, int* rc) {
    *rc = 0;
    if (cudaSetDevice(device) != cudaSuccess) { *rc = 1; return; }
    const size_t nchunks = (bytes + HC_CHUNK - 1) / HC_CHUNK;
    cudaStream_t st = hc->stream[t];
    int j = 0;
    size_t prev_off = 0, prev_len = 0;
    int prev_b = -1;
    for (size_t k = size_t(t); k < nchunks; k += size_t(T), j++) {
        const size_t off = k * HC_CHUNK, len = std::min(HC_CHUNK, bytes - off);
        const int b = j & 1;
        if (h2d) {
            if (j >= 2 && cudaEventSynchronize(hc->ev[t][b]) != cudaSuccess) { *rc = 1; return; }  // buffer free again
            memcpy(hc->buf[t][b], src + off, len);
            if (cudaMemcpyAsync(dst + off, hc->buf[t][b], len, cudaMemcpyHostToDevice, st) != cudaSuccess) { *rc = 1; return; }
            cudaEventRecord(hc->ev[t][b], st);
        } else {
            // buffer b was drained two iterations ago (its memcpy-out ran before this point)
            if (cudaMemcpyAsync(hc->buf[t][b], src + off, len, cudaMemcpyDeviceToHost, st) != cudaSuccess) { *rc = 1; return; }
            cudaEventRecord(hc->ev[t][b], st);
            if (prev_b >= 0) {
                if (cudaEventSynchronize(hc->ev[t][prev_b]) != cudaSuccess) { *rc = 1; return; }
                memcpy(dst + prev_off, hc->buf[t][prev_b], prev_len);
            }
            prev_b = b;
            prev_off = off;
            prev_len = len;
        }
    }
    if (!h2d && prev_b >= 0) {
        if (cudaEventSynchronize(hc->ev[t][prev_b]) != cudaSuccess) { *rc = 1; return; }
        memcpy(dst + prev_off, hc->buf[t][prev_b], prev_len);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) *rc = 1;
}

// Blocking copy between a caller's HOST buffer and device memory; everything queued on c.stream before the call is
// finished first, and the copy is complete on return.
int host_copy(Ctx& c, void* dst, const void* src, size_t bytes, bool h2d) {
    if (bytes == 0) return 0;
    const void* host = h2d ? src : dst;
    TDC_CUDA(cudaStreamSynchronize(c.stream));
    if (bytes < HC_MIN_STAGED || !is_pageable(host)) {
        TDC_CUDA(cudaMemcpyAsync(dst, src, bytes, h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, c.stream));
        TDC_CUDA(cudaStreamSynchronize(c.stream));
        return 0;
    }
    TDC_TRY(host_copier_init(c.copier));
    const size_t nchunks = (bytes + HC_CHUNK - 1) / HC_CHUNK;
    const int T = int(std::min<size_t>(HC_THREADS, nchunks));
    int rc[HC_THREADS] = {0};
    std::thread th[HC_THREADS];
    for (int t = 1; t < T; t++)
        th[t] = std::thread(copy_worker, &c.copier, c.device, t, T, static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), bytes, h2d, &rc[t]);
    copy_worker(&c.copier, c.device, 0, T, static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), bytes, h2d, &rc[0]);
    for (int t = 1; t < T; t++) th[t].join();
    for (int t = 0; t < T; t++)
        if (rc[t]) { set_error("host_copy: staged %s copy failed: %s", h2d ? "H2D" : "D2H", cudaGetErrorString(cudaGetLastError())); return -1; }
    return 0;
}

}  // namespace tdc
