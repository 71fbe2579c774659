// Rank-to-rank exchange used by the sharded multi-GPU text index.  One rank = one process = one GPU.
// The product implementation is NCCL (grouped ncclSend/ncclRecv = all-to-all over NVLink), loaded with dlopen so that
// libtdcgpu.so itself has no link-time dependency on NCCL (the single-GPU `tdc` driver does not need it).
// The CPU simulator build (tests/sim, test infrastructure only) plugs in callbacks instead, which the CPU tests back
// with torch.distributed/gloo.
#pragma once
#include "tdc_common.cuh"

namespace tdc {

struct Comm {
    int rank = 0, nranks = 1;
    virtual ~Comm() {}
    // every rank contributes `bytes` bytes of HOST memory; recv (host) receives nranks * bytes in rank order
    virtual int allgather_host(const void* send, void* recv, size_t bytes) = 0;
    // DEVICE buffers; byte offsets and byte counts per peer; returns after the data has arrived
    virtual int alltoallv(const void* dsend, const u64* soff, const u64* scnt, void* drecv, const u64* roff, const u64* rcnt,
                          cudaStream_t st) = 0;
    // Peer-memory transport (NVLink / NVSwitch): every rank registers ONE device allocation of the same size (the scratch
    // arena); peers[p] receives a pointer through which rank p's allocation can be read and written from this rank's
    // kernels and copies (peers[rank] = local).  Returns < 0 when the transport is not available (the caller then keeps
    // using alltoallv).  Collective.  close_window unmaps.
    virtual int open_window(void* local, size_t bytes, void** peers) { (void)local; (void)bytes; (void)peers; return -1; }
    virtual void close_window() {}
};

#ifdef TDC_CUSIM
typedef int (*tdcsim_allgather_fn)(void* user, const void* send, void* recv, uint64_t bytes);
typedef int (*tdcsim_alltoallv_fn)(void* user, const void* send, const uint64_t* soff, const uint64_t* scnt, void* recv,
                                   const uint64_t* roff, const uint64_t* rcnt);
Comm* make_callback_comm(int rank, int nranks, tdcsim_allgather_fn ag, tdcsim_alltoallv_fn a2a, void* user);
#else
int nccl_unique_id(uint8_t out[128]);
Comm* make_nccl_comm(int rank, int nranks, const uint8_t id[128], cudaStream_t st);  // nullptr on failure (error set)
#endif

}  // namespace tdc
