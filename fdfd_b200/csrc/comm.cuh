// Multi-GPU plumbing: y-slab neighbours exchange one grid row per apply, Krylov inner products
// are summed across ranks.  NCCL is resolved at run time (dlopen of the libnccl.so.2 the host
// framework already loaded), so the library itself has no link-time NCCL dependency and loads
// on a CPU-only box for the symbol checks.
#pragma once
#include "common.cuh"

namespace fdfd {

struct P2PHeap;   // p2p.cuh

struct Comm {
    int nranks, rank;
    void* nccl;             // ncclComm_t (bootstrap, fallback data path)
    P2PHeap* p2p = nullptr; // peer-mapped symmetric heap (null: NCCL data path)
};

int comm_unique_id(void* id128);
int comm_create(Comm** out, int nranks, int rank, const void* id128);
void comm_destroy(Comm*);

// In-place sum of `count` doubles across ranks (no-op when comm is null / single rank).
int comm_allreduce_sum(Comm* c, double* buf, int count, cudaStream_t stream);

// Ghost-row exchange for a slab of Ny rows of Nx elements of `elem` bytes:
//   halo_south <- last row of rank-1, halo_north <- first row of rank+1
// (cyclic when `periodic`).  x is the local slab.
int comm_halo_rows(Comm* c, const void* x, void* halo_south, void* halo_north, int Nx, int Ny,
                   size_t elem, bool periodic, cudaStream_t stream);
// recv (nranks * bytes) <- concatenation of every rank's `send` (bytes each), rank order (NCCL; set-up paths and
// the fallback data path — the hot paths use the peer-memory plans of p2p.cuh)
int comm_allgather(Comm* c, const void* send, void* recv, size_t bytes, cudaStream_t stream);
int comm_allgather_nccl(Comm* c, const void* send, void* recv, size_t bytes, cudaStream_t stream);
// Generic neighbour exchange of arbitrary rows (used once for static coefficient ghosts).
int comm_sendrecv(Comm* c, const void* send, int dst, void* recv, int src, size_t bytes,
                  cudaStream_t stream);

}  // namespace fdfd
