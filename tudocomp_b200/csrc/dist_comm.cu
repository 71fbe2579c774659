// Comm implementations (see dist_comm.h).
#include "dist_comm.h"

#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef TDC_CUSIM
#include <dlfcn.h>
#include <nccl.h>  // types only: every entry point is resolved with dlsym
#endif

namespace tdc {

#ifdef TDC_CUSIM
// ---------------------------------------------------------------------------------------------------------------
// simulator: callbacks supplied by the test harness ("device" memory is host memory here)
// ---------------------------------------------------------------------------------------------------------------
struct CallbackComm : Comm {
    tdcsim_allgather_fn ag;
    tdcsim_alltoallv_fn a2a;
    void* user;
    int allgather_host(const void* send, void* recv, size_t bytes) override {
        if (nranks == 1) { memcpy(recv, send, bytes); return 0; }
        if (ag(user, send, recv, bytes) != 0) { set_error("allgather callback failed"); return -1; }
        return 0;
    }
    int alltoallv(const void* dsend, const u64* soff, const u64* scnt, void* drecv, const u64* roff, const u64* rcnt,
                  cudaStream_t) override {
        if (nranks == 1) {
            if (scnt[0]) memmove((char*)drecv + roff[0], (const char*)dsend + soff[0], scnt[0]);
            return 0;
        }
        if (a2a(user, dsend, soff, scnt, drecv, roff, rcnt) != 0) { set_error("alltoallv callback failed"); return -1; }
        return 0;
    }
};
Comm* make_callback_comm(int rank, int nranks, tdcsim_allgather_fn ag, tdcsim_alltoallv_fn a2a, void* user) {
    CallbackComm* c = new CallbackComm();
    c->rank = rank;
    c->nranks = nranks;
    c->ag = ag;
    c->a2a = a2a;
    c->user = user;
    return c;
}
#else
// ---------------------------------------------------------------------------------------------------------------
// NCCL
// ---------------------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return 0;
    // an NCCL already mapped into the process (e.g. the one PyTorch ships) is found by its soname
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("multi-GPU mode needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror()); return -1; }
#define TDC_SYM(field, name)                                                           \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                         \
    if (!g_nccl.field) { set_error("NCCL symbol %s not found", name); dlclose(h); return -1; }
    TDC_SYM(GetUniqueId, "ncclGetUniqueId");
    TDC_SYM(CommInitRank, "ncclCommInitRank");
    TDC_SYM(CommDestroy, "ncclCommDestroy");
    TDC_SYM(Send, "ncclSend");
    TDC_SYM(Recv, "ncclRecv");
    TDC_SYM(AllGather, "ncclAllGather");
    TDC_SYM(GroupStart, "ncclGroupStart");
    TDC_SYM(GroupEnd, "ncclGroupEnd");
    TDC_SYM(GetErrorString, "ncclGetErrorString");
#undef TDC_SYM
    g_nccl.handle = h;
    return 0;
}
#define TDC_NCCL(expr)                                                                                    \
    do {                                                                                                  \
        ncclResult_t r__ = (expr);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__));          \
            return -1;                                                                                    \
        }                                                                                                 \
    } while (0)

int nccl_unique_id(uint8_t out[128]) {
    TDC_TRY(nccl_load());
    ncclUniqueId id;
    TDC_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    memcpy(out, &id, 128);
    return 0;
}

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    cudaStream_t st = 0;
    uint8_t* d_stage = nullptr;  // [(1 + nranks) * STAGE] device staging of the host all-gather
    uint8_t* h_stage = nullptr;  // pinned mirror
    static const size_t STAGE = size_t(1) << 20;
    void* mapped[64] = {nullptr};  // peer mappings of the registered window (cudaIpcOpenMemHandle)
    ~NcclComm() override {
        close_window();
        if (comm) g_nccl.CommDestroy(comm);
        if (d_stage) cudaFree(d_stage);
        if (h_stage) cudaFreeHost(h_stage);
    }
    int allgather_host(const void* send, void* recv, size_t bytes) override {
        if (nranks == 1) { memcpy(recv, send, bytes); return 0; }
        if (bytes > STAGE) { set_error("allgather_host: %zu bytes exceed the staging buffer", bytes); return -1; }
        memcpy(h_stage, send, bytes);
        TDC_CUDA(cudaMemcpyAsync(d_stage, h_stage, bytes, cudaMemcpyHostToDevice, st));
        TDC_NCCL(g_nccl.AllGather(d_stage, d_stage + STAGE, bytes, ncclUint8, comm, st));
        TDC_CUDA(cudaMemcpyAsync(h_stage + STAGE, d_stage + STAGE, bytes * nranks, cudaMemcpyDeviceToHost, st));
        TDC_CUDA(cudaStreamSynchronize(st));
        memcpy(recv, h_stage + STAGE, bytes * nranks);
        return 0;
    }
    int alltoallv(const void* dsend, const u64* soff, const u64* scnt, void* drecv, const u64* roff, const u64* rcnt,
                  cudaStream_t stream) override {
        const char* s = static_cast<const char*>(dsend);
        char* r = static_cast<char*>(drecv);
        if (scnt[rank]) TDC_CUDA(cudaMemcpyAsync(r + roff[rank], s + soff[rank], scnt[rank], cudaMemcpyDeviceToDevice, stream));
        if (nranks > 1) {
            TDC_NCCL(g_nccl.GroupStart());
            for (int p = 0; p < nranks; p++) {
                if (p == rank) continue;
                if (scnt[p]) TDC_NCCL(g_nccl.Send(s + soff[p], scnt[p], ncclUint8, p, comm, stream));
                if (rcnt[p]) TDC_NCCL(g_nccl.Recv(r + roff[p], rcnt[p], ncclUint8, p, comm, stream));
            }
            TDC_NCCL(g_nccl.GroupEnd());
        }
        return 0;  // stream-ordered: consumers run on the same stream
    }
    // CUDA IPC: one process per GPU, so a peer's allocation is mapped through an IPC handle passed over the host all-gather
    int open_window(void* local, size_t bytes, void** peers) override {
        (void)bytes;
        close_window();
        if (nranks == 1) { peers[0] = local; return 0; }
        if (getenv("TDCGPU_DIST_NO_P2P")) return -1;
        struct Item { cudaIpcMemHandle_t h; int dev; int ok; };
        Item mine;
        memset(&mine, 0, sizeof(mine));
        mine.ok = cudaIpcGetMemHandle(&mine.h, local) == cudaSuccess ? 1 : 0;
        cudaGetDevice(&mine.dev);
        std::vector<Item> all(nranks);
        if (allgather_host(&mine, all.data(), sizeof(Item)) < 0) return -1;
        int good = 1;
        for (int p = 0; p < nranks; p++) good &= all[p].ok;
        if (good) {
            for (int p = 0; p < nranks && good; p++) {
                if (p == rank) { peers[p] = local; continue; }
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = 0; break; }
                mapped[p] = ptr;
                peers[p] = ptr;
            }
        }
        cudaGetLastError();  // clear a sticky-free error state of a failed probe
        // all or nothing: a rank that could not map must not leave the others on a different transport
        int flag = good, sum = 0;
        std::vector<int> flags(nranks);
        if (allgather_host(&flag, flags.data(), sizeof(int)) < 0) return -1;
        for (int p = 0; p < nranks; p++) sum += flags[p];
        if (sum != nranks) { close_window(); return -1; }
        return 0;
    }
    void close_window() override {
        for (int p = 0; p < 64; p++)
            if (mapped[p]) { cudaIpcCloseMemHandle(mapped[p]); mapped[p] = nullptr; }
    }
};

Comm* make_nccl_comm(int rank, int nranks, const uint8_t id[128], cudaStream_t st) {
    NcclComm* c = new NcclComm();
    c->rank = rank;
    c->nranks = nranks;
    c->st = st;
    if (nranks > 1) {
        if (nccl_load() < 0) { delete c; return nullptr; }
        ncclUniqueId uid;
        memcpy(&uid, id, 128);
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
        if (r != ncclSuccess) {
            set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, nranks, g_nccl.GetErrorString(r));
            c->comm = nullptr;
            delete c;
            return nullptr;
        }
        if (cudaMalloc(&c->d_stage, NcclComm::STAGE * size_t(1 + nranks)) != cudaSuccess ||
            cudaMallocHost(&c->h_stage, NcclComm::STAGE * size_t(1 + nranks)) != cudaSuccess) {
            set_error("comm staging allocation failed");
            delete c;
            return nullptr;
        }
    }
    return c;
}
#endif

}  // namespace tdc
