// NVLink peer-memory layer (see p2p.cuh): symmetric heap over CUDA IPC, push/flag collectives.
#include "p2p.cuh"
#include "comm.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace fdfd {

struct P2PHeap {
    char* base[kP2PMaxRanks];
    size_t bytes = 0;
    std::map<size_t, size_t> free_list;   // offset -> size
    std::map<size_t, size_t> used;
    GatherPlan* red = nullptr;            // slots of the all-reduce
    std::mutex mu;
};

namespace {

// bootstrap all-gather of host data through NCCL
int host_allgather(Comm* c, const void* mine, size_t bytes, void* all, cudaStream_t st) {
    char* d = nullptr;
    FDFD_CUDA(cudaMalloc(&d, bytes * (size_t)(c->nranks + 1)));
    int s = FDFD_OK;
    cudaError_t e = cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) s = comm_allgather_nccl(c, d, d + bytes, bytes, st);
    if (e == cudaSuccess && s == FDFD_OK)
        e = cudaMemcpyAsync(all, d + bytes, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (s != FDFD_OK) return s;
    if (e != cudaSuccess) return cuda_fail(e, "host_allgather", __FILE__, __LINE__);
    return FDFD_OK;
}

__device__ __forceinline__ unsigned int p_ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void p_st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int p_ld_volatile(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// the last CTA of a launch advances the epoch (all CTAs read it at their start)
__device__ __forceinline__ void p_finish(unsigned int* state, unsigned int ep, unsigned int ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(state + 1, 1u);
        if (prev == ctas - 1) {
            state[1] = 0;
            __threadfence();
            state[0] = ep + 1;
        }
    }
}

// all-gather: CTA g moves slice g of this rank's block to every peer, raises flag (me, g) there, waits for
// flag (r, g) of every other rank here and copies the received slices out of the parity buffer
__global__ void __launch_bounds__(256) p2p_allgather_kernel(GatherPlan p, const uint2* __restrict__ send,
                                                            uint2* __restrict__ recv, size_t words) {
    const unsigned int ep = p_ld_volatile(p.state), par = ep & 1u;
    const int me = p.rank, P = p.nranks, g = blockIdx.x, G = gridDim.x;
    const size_t per = (words + G - 1) / G;
    const size_t w0 = (size_t)g * per, w1 = min(words, w0 + per);
    const size_t capw = p.cap / 8;
    for (size_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
        const uint2 v = send[w];
        recv[(size_t)me * words + w] = v;
        for (int r = 0; r < P; ++r)
            if (r != me) reinterpret_cast<uint2*>(p.rx_peer[r])[((size_t)par * P + me) * capw + w] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < P && (int)threadIdx.x != me) {
        const int r = threadIdx.x;
        __threadfence_system();
        p_st_release_sys(p.flag_peer[r] + ((size_t)par * P + me) * kP2PGatherCtas + g, ep + 1);
        while (p_ld_acquire_sys(p.flag_peer[me] + ((size_t)par * P + r) * kP2PGatherCtas + g) != ep + 1) { }
    }
    __syncthreads();
    const uint2* rx = reinterpret_cast<const uint2*>(p.rx_peer[me]);
    for (int r = 0; r < P; ++r) {
        if (r == me) continue;
        for (size_t w = w0 + threadIdx.x; w < w1; w += blockDim.x)
            recv[(size_t)r * words + w] = __ldcg(&rx[((size_t)par * P + r) * capw + w]);
    }
    p_finish(p.state, ep, G);
}

// all-reduce (sum) of `count` doubles: one CTA; every rank sums the contributions in rank order
__global__ void __launch_bounds__(kP2PRedMax) p2p_allreduce_kernel(GatherPlan p, double* buf, int count) {
    const unsigned int ep = p_ld_volatile(p.state), par = ep & 1u;
    const int me = p.rank, P = p.nranks, t = threadIdx.x;
    const size_t capw = p.cap / 8;
    double v = 0.0;
    if (t < count) {
        v = buf[t];
        for (int r = 0; r < P; ++r)
            if (r != me) reinterpret_cast<double*>(p.rx_peer[r])[((size_t)par * P + me) * capw + t] = v;
    }
    __syncthreads();
    if (t < P && t != me) {
        __threadfence_system();
        p_st_release_sys(p.flag_peer[t] + ((size_t)par * P + me) * kP2PGatherCtas, ep + 1);
        while (p_ld_acquire_sys(p.flag_peer[me] + ((size_t)par * P + t) * kP2PGatherCtas) != ep + 1) { }
    }
    __syncthreads();
    if (t < count) {
        const double* rx = reinterpret_cast<const double*>(p.rx_peer[me]);
        double s = 0.0;
        for (int r = 0; r < P; ++r) s += (r == me) ? v : __ldcg(&rx[((size_t)par * P + r) * capw + t]);
        buf[t] = s;
    }
    p_finish(p.state, ep, 1);
}

// stand-alone ghost-row exchange: segment g (128 elements) of both edge rows
__global__ void __launch_bounds__(128) p2p_halo_kernel(HaloP2P h, const char* __restrict__ x, char* gs, char* gn,
                                                       int Nx, int Ny, int elem) {
    const unsigned int ep = p_ld_volatile(h.state), par = ep & 1u;
    const int wpe = elem / 8;                                   // 8-byte words per element
    const size_t roww = (size_t)Nx * wpe;
    const size_t w0 = (size_t)blockIdx.x * 128 * wpe, w1 = min(roww, w0 + (size_t)128 * wpe);
    const uint2* xw = reinterpret_cast<const uint2*>(x);
    for (size_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
        if (h.push_north) reinterpret_cast<uint2*>(h.push_north)[(size_t)par * roww + w] = xw[(size_t)(Ny - 1) * roww + w];
        if (h.push_south) reinterpret_cast<uint2*>(h.push_south)[(size_t)par * roww + w] = xw[w];
    }
    __syncthreads();
    // a flag covers 32 columns: this CTA owns flags [4 g, 4 g + 4)
    if (threadIdx.x < 4) {
        const int f = blockIdx.x * 4 + threadIdx.x;
        if (f < h.segs) {
            __threadfence_system();
            if (h.flag_push_north) p_st_release_sys(h.flag_push_north + par * h.segs + f, ep + 1);
            if (h.flag_push_south) p_st_release_sys(h.flag_push_south + par * h.segs + f, ep + 1);
            if (h.flag_south) while (p_ld_acquire_sys(h.flag_south + par * h.segs + f) != ep + 1) { }
            if (h.flag_north) while (p_ld_acquire_sys(h.flag_north + par * h.segs + f) != ep + 1) { }
        }
    }
    __syncthreads();
    for (size_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
        if (h.flag_south) reinterpret_cast<uint2*>(gs)[w] = __ldcg(reinterpret_cast<const uint2*>(h.rx_south) + (size_t)par * roww + w);
        if (h.flag_north) reinterpret_cast<uint2*>(gn)[w] = __ldcg(reinterpret_cast<const uint2*>(h.rx_north) + (size_t)par * roww + w);
    }
    p_finish(h.state, ep, gridDim.x);
}

}  // namespace

bool p2p_enabled(const Comm* c) { return c && c->p2p != nullptr; }

int p2p_init(Comm* c) {
    c->p2p = nullptr;
    if (!c || c->nranks <= 1) return FDFD_OK;
    const char* env = getenv("FDFD_P2P");
    if (env && atoi(env) == 0) return FDFD_OK;
    if (c->nranks > kP2PMaxRanks) return FDFD_OK;
    size_t mb = 256;
    if (const char* e = getenv("FDFD_P2P_HEAP_MB")) mb = (size_t)std::max(16, atoi(e));
    struct Boot { cudaIpcMemHandle_t h; int ok; int pad; } mine, all[kP2PMaxRanks];
    memset(&mine, 0, sizeof(mine));
    P2PHeap* H = new P2PHeap();
    H->bytes = mb << 20;
    for (int r = 0; r < kP2PMaxRanks; ++r) H->base[r] = nullptr;
    char* base = nullptr;
    mine.ok = cudaMalloc(&base, H->bytes) == cudaSuccess;
    if (mine.ok) mine.ok = cudaMemset(base, 0, H->bytes) == cudaSuccess;
    if (mine.ok) mine.ok = cudaIpcGetMemHandle(&mine.h, base) == cudaSuccess;
    cudaGetLastError();
    int s = host_allgather(c, &mine, sizeof(Boot), all, 0);
    bool ok = s == FDFD_OK;
    for (int r = 0; ok && r < c->nranks; ++r) ok = all[r].ok != 0;
    int mapped = ok ? 1 : 0;
    if (ok) {
        H->base[c->rank] = base;
        for (int r = 0; r < c->nranks; ++r) {
            if (r == c->rank) continue;
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                mapped = 0;
                break;
            }
            H->base[r] = (char*)ptr;
        }
    }
    int flags[kP2PMaxRanks];
    if (s == FDFD_OK) s = host_allgather(c, &mapped, sizeof(int), flags, 0);
    bool all_mapped = s == FDFD_OK;
    for (int r = 0; all_mapped && r < c->nranks; ++r) all_mapped = flags[r] != 0;
    if (!all_mapped) {
        for (int r = 0; r < c->nranks; ++r)
            if (r != c->rank && H->base[r]) cudaIpcCloseMemHandle(H->base[r]);
        if (base) cudaFree(base);
        delete H;
        if (getenv("FDFD_P2P_VERBOSE")) fprintf(stderr, "[fdfd] rank %d: peer mapping unavailable, NCCL data path\n", c->rank);
        return s;
    }
    H->free_list[0] = H->bytes;
    c->p2p = H;
    s = p2p_gather_create(c, (size_t)kP2PRedMax * sizeof(double), &H->red, 0);
    if (s != FDFD_OK) { p2p_shutdown(c); return s; }
    if (getenv("FDFD_P2P_VERBOSE")) fprintf(stderr, "[fdfd] rank %d: %zu MB symmetric heap mapped to %d peers\n", c->rank, mb, c->nranks - 1);
    return FDFD_OK;
}

void p2p_shutdown(Comm* c) {
    if (!c || !c->p2p) return;
    P2PHeap* H = c->p2p;
    if (H->red) { cudaFree(H->red->state); delete H->red; }
    cudaDeviceSynchronize();
    for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && H->base[r]) cudaIpcCloseMemHandle(H->base[r]);
    cudaFree(H->base[c->rank]);
    delete H;
    c->p2p = nullptr;
}

int p2p_alloc(Comm* c, size_t bytes, size_t* off) {
    P2PHeap* H = c->p2p;
    FDFD_CHECK_ARG(H, "no symmetric heap");
    bytes = (bytes + 255) & ~(size_t)255;
    std::lock_guard<std::mutex> lk(H->mu);
    for (auto it = H->free_list.begin(); it != H->free_list.end(); ++it) {
        if (it->second >= bytes) {
            const size_t o = it->first, rest = it->second - bytes;
            H->free_list.erase(it);
            if (rest) H->free_list[o + bytes] = rest;
            H->used[o] = bytes;
            *off = o;
            FDFD_CUDA(cudaMemset(H->base[c->rank] + o, 0, bytes));
            return FDFD_OK;
        }
    }
    set_error("symmetric heap exhausted (%zu bytes requested; raise FDFD_P2P_HEAP_MB)", bytes);
    return FDFD_ERR_STATE;
}

void p2p_free(Comm* c, size_t off) {
    if (!c || !c->p2p) return;
    P2PHeap* H = c->p2p;
    std::lock_guard<std::mutex> lk(H->mu);
    auto u = H->used.find(off);
    if (u == H->used.end()) return;
    size_t o = off, sz = u->second;
    H->used.erase(u);
    auto nx = H->free_list.lower_bound(o);
    if (nx != H->free_list.end() && o + sz == nx->first) { sz += nx->second; nx = H->free_list.erase(nx); }
    if (nx != H->free_list.begin()) {
        auto pv = std::prev(nx);
        if (pv->first + pv->second == o) { o = pv->first; sz += pv->second; H->free_list.erase(pv); }
    }
    H->free_list[o] = sz;
}

void* p2p_local(Comm* c, size_t off) { return c->p2p->base[c->rank] + off; }

int p2p_exchange(Comm* c, const size_t* mine, int k, void** ptrs, cudaStream_t st) {
    P2PHeap* H = c->p2p;
    FDFD_CHECK_ARG(H && k >= 1 && k <= 16, "p2p_exchange");
    std::vector<size_t> all((size_t)k * c->nranks);
    // the exchange is also the barrier that orders "my buffers are zeroed" before any peer's first push
    FDFD_CUDA(cudaStreamSynchronize(st));
    FDFD_TRY(host_allgather(c, mine, sizeof(size_t) * k, all.data(), st));
    for (int r = 0; r < c->nranks; ++r)
        for (int i = 0; i < k; ++i) ptrs[r * k + i] = H->base[r] + all[(size_t)r * k + i];
    return FDFD_OK;
}

int p2p_gather_create(Comm* c, size_t cap, GatherPlan** out, cudaStream_t st) {
    FDFD_CHECK_ARG(p2p_enabled(c), "no symmetric heap");
    cap = (cap + 15) & ~(size_t)15;
    GatherPlan* p = new GatherPlan();
    p->nranks = c->nranks; p->rank = c->rank; p->cap = cap;
    int s = p2p_alloc(c, 2 * (size_t)c->nranks * cap, &p->off_rx);
    if (s == FDFD_OK) s = p2p_alloc(c, 2 * (size_t)c->nranks * kP2PGatherCtas * sizeof(unsigned int), &p->off_flags);
    if (s == FDFD_OK && cudaMalloc(&p->state, 2 * sizeof(unsigned int)) != cudaSuccess) s = FDFD_ERR_CUDA;
    if (s == FDFD_OK) cudaMemset(p->state, 0, 2 * sizeof(unsigned int));
    void* ptrs[2 * kP2PMaxRanks];
    const size_t mine[2] = {p->off_rx, p->off_flags};
    if (s == FDFD_OK) s = p2p_exchange(c, mine, 2, ptrs, st);
    if (s != FDFD_OK) { delete p; return s; }
    for (int r = 0; r < c->nranks; ++r) {
        p->rx_peer[r] = (char*)ptrs[2 * r];
        p->flag_peer[r] = (unsigned int*)ptrs[2 * r + 1];
    }
    *out = p;
    return FDFD_OK;
}

void p2p_gather_destroy(Comm* c, GatherPlan* p) {
    if (!p) return;
    cudaFree(p->state);
    p2p_free(c, p->off_rx);
    p2p_free(c, p->off_flags);
    delete p;
}

int p2p_allgather(GatherPlan* p, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    FDFD_CHECK_ARG(p && bytes <= p->cap && bytes % 8 == 0, "p2p_allgather: block does not fit the plan");
    const size_t words = bytes / 8;
    int G = (int)((bytes + 8191) / 8192);
    if (G < 1) G = 1;
    if (G > kP2PGatherCtas) G = kP2PGatherCtas;
    p2p_allgather_kernel<<<G, 256, 0, st>>>(*p, (const uint2*)send, (uint2*)recv, words);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

int p2p_allreduce_sum(Comm* c, double* buf, int count, cudaStream_t st) {
    FDFD_CHECK_ARG(p2p_enabled(c) && count >= 1 && count <= kP2PRedMax, "p2p_allreduce_sum: count");
    p2p_allreduce_kernel<<<1, kP2PRedMax, 0, st>>>(*c->p2p->red, buf, count);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

int p2p_halo_create(Comm* c, int Nx, int south, int north, HaloP2P* out, size_t offs[4], cudaStream_t st) {
    FDFD_CHECK_ARG(p2p_enabled(c), "no symmetric heap");
    HaloP2P h;
    h.segs = (Nx + 31) / 32;
    const size_t rowb = 2 * (size_t)Nx * 16, flagb = 2 * (size_t)h.segs * sizeof(unsigned int);
    FDFD_TRY(p2p_alloc(c, rowb, &offs[0]));
    FDFD_TRY(p2p_alloc(c, rowb, &offs[1]));
    FDFD_TRY(p2p_alloc(c, flagb, &offs[2]));
    FDFD_TRY(p2p_alloc(c, flagb, &offs[3]));
    FDFD_CUDA(cudaMalloc(&h.state, 2 * sizeof(unsigned int)));
    FDFD_CUDA(cudaMemset(h.state, 0, 2 * sizeof(unsigned int)));
    void* ptrs[4 * kP2PMaxRanks];
    FDFD_TRY(p2p_exchange(c, offs, 4, ptrs, st));
    h.rx_south = p2p_local(c, offs[0]); h.rx_north = p2p_local(c, offs[1]);
    if (south >= 0) {
        h.flag_south = (const unsigned int*)p2p_local(c, offs[2]);
        h.push_south = ptrs[4 * south + 1];                       // its rx_north
        h.flag_push_south = (unsigned int*)ptrs[4 * south + 3];   // its flag_north
    }
    if (north >= 0) {
        h.flag_north = (const unsigned int*)p2p_local(c, offs[3]);
        h.push_north = ptrs[4 * north + 0];                       // its rx_south
        h.flag_push_north = (unsigned int*)ptrs[4 * north + 2];   // its flag_south
    }
    *out = h;
    return FDFD_OK;
}

void p2p_halo_destroy(Comm* c, HaloP2P* h, const size_t offs[4]) {
    if (!h || !h->state) return;
    cudaFree(h->state);
    for (int i = 0; i < 4; ++i) p2p_free(c, offs[i]);
    *h = HaloP2P();
}

int p2p_halo_exchange(const HaloP2P& h, const void* x, void* gs, void* gn, int Nx, int Ny, size_t elem,
                      cudaStream_t st) {
    p2p_halo_kernel<<<(Nx + 127) / 128, 128, 0, st>>>(h, (const char*)x, (char*)gs, (char*)gn, Nx, Ny, (int)elem);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

}  // namespace fdfd
