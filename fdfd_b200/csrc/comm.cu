// NCCL-backed slab communication (see comm.cuh).  Function pointers are looked up with dlsym.
#include "comm.cuh"
#include "p2p.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

namespace fdfd {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_api;
std::once_flag g_once;
bool g_ok = false;

void load_api() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) return;
#define SYM(field, name) *(void**)(&g_api.field) = dlsym(g_api.lib, name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllReduce && g_api.AllGather &&
           g_api.Send && g_api.Recv && g_api.GroupStart && g_api.GroupEnd;
}

int need_api() {
    std::call_once(g_once, load_api);
    if (!g_ok) { set_error("NCCL (libnccl.so.2) could not be loaded"); return FDFD_ERR_NCCL; }
    return FDFD_OK;
}

int nccl_fail(ncclResult_t r, const char* what) {
    set_error("NCCL failure in %s: %s", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
    return FDFD_ERR_NCCL;
}
#define FDFD_NCCL(call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) return nccl_fail(_r, #call); } while (0)
}  // namespace

int comm_unique_id(void* id128) {
    FDFD_TRY(need_api());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    FDFD_NCCL(g_api.GetUniqueId((ncclUniqueId*)id128));
    return FDFD_OK;
}

int comm_create(Comm** out, int nranks, int rank, const void* id128) {
    FDFD_CHECK_ARG(out && nranks >= 1 && rank >= 0 && rank < nranks, "bad communicator shape");
    Comm* c = new Comm{nranks, rank, nullptr, nullptr};
    if (nranks > 1) {
        int st = need_api();
        if (st != FDFD_OK) { delete c; return st; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclComm_t comm;
        ncclResult_t r = g_api.CommInitRank(&comm, nranks, id, rank);
        if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
        c->nccl = comm;
        // peer-mapped symmetric heap for the data path (collective; falls back to NCCL when unavailable)
        st = p2p_init(c);
        if (st != FDFD_OK) { g_api.CommDestroy(comm); delete c; return st; }
    }
    *out = c;
    return FDFD_OK;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    p2p_shutdown(c);
    if (c->nccl && g_ok) g_api.CommDestroy((ncclComm_t)c->nccl);
    delete c;
}

int comm_allreduce_sum(Comm* c, double* buf, int count, cudaStream_t stream) {
    if (!c || c->nranks == 1) return FDFD_OK;
    if (p2p_enabled(c) && count <= kP2PRedMax) return p2p_allreduce_sum(c, buf, count, stream);
    FDFD_NCCL(g_api.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)c->nccl, stream));
    return FDFD_OK;
}

int comm_allgather(Comm* c, const void* send, void* recv, size_t bytes, cudaStream_t stream) {
    if (!c || c->nranks == 1) {
        if (send != recv) FDFD_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, stream));
        return FDFD_OK;
    }
    return comm_allgather_nccl(c, send, recv, bytes, stream);
}

int comm_allgather_nccl(Comm* c, const void* send, void* recv, size_t bytes, cudaStream_t stream) {
    FDFD_NCCL(g_api.AllGather(send, recv, bytes, ncclChar, (ncclComm_t)c->nccl, stream));
    return FDFD_OK;
}

int comm_sendrecv(Comm* c, const void* send, int dst, void* recv, int src, size_t bytes,
                  cudaStream_t stream) {
    if (!c || c->nranks == 1) return FDFD_OK;
    // inside a group the first failure is remembered and the group is always closed
    ncclResult_t first = ncclSuccess;
    auto keep = [&](ncclResult_t r) { if (first == ncclSuccess && r != ncclSuccess) first = r; };
    FDFD_NCCL(g_api.GroupStart());
    if (send && dst >= 0) keep(g_api.Send(send, bytes, ncclChar, dst, (ncclComm_t)c->nccl, stream));
    if (recv && src >= 0) keep(g_api.Recv(recv, bytes, ncclChar, src, (ncclComm_t)c->nccl, stream));
    keep(g_api.GroupEnd());
    if (first != ncclSuccess) return nccl_fail(first, "ncclSend/ncclRecv group");
    return FDFD_OK;
}

int comm_halo_rows(Comm* c, const void* x, void* halo_south, void* halo_north, int Nx, int Ny,
                   size_t elem, bool periodic, cudaStream_t stream) {
    if (!c || c->nranks == 1) return FDFD_OK;
    const int P = c->nranks, r = c->rank;
    const int south = r > 0 ? r - 1 : (periodic ? P - 1 : -1);
    const int north = r < P - 1 ? r + 1 : (periodic ? 0 : -1);
    const size_t bytes = (size_t)Nx * elem;
    const char* first_row = (const char*)x;
    const char* last_row = (const char*)x + (size_t)(Ny - 1) * bytes;
    ncclComm_t comm = (ncclComm_t)c->nccl;
    // Order matters when both neighbours are the same peer (2 ranks, periodic): NCCL pairs the
    // sends and receives between two ranks in issue order, so "my last row -> north's south ghost"
    // is issued (and received) first on every rank, "my first row -> south's north ghost" second.
    ncclResult_t first = ncclSuccess;
    auto keep = [&](ncclResult_t r) { if (first == ncclSuccess && r != ncclSuccess) first = r; };
    FDFD_NCCL(g_api.GroupStart());
    if (north >= 0) keep(g_api.Send(last_row, bytes, ncclChar, north, comm, stream));
    if (south >= 0) keep(g_api.Recv(halo_south, bytes, ncclChar, south, comm, stream));
    if (south >= 0) keep(g_api.Send(first_row, bytes, ncclChar, south, comm, stream));
    if (north >= 0) keep(g_api.Recv(halo_north, bytes, ncclChar, north, comm, stream));
    keep(g_api.GroupEnd());
    if (first != ncclSuccess) return nccl_fail(first, "halo ncclSend/ncclRecv group");
    return FDFD_OK;
}

}  // namespace fdfd
