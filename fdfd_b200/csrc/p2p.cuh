// NVLink peer-memory layer of the y-slab sharding (one process per GPU).
//
// Every rank allocates one "symmetric heap" with cudaMalloc, exports it with CUDA IPC and maps the heaps of
// all other ranks of the node (NVSwitch: every GPU reaches every peer at full bandwidth).  Receive buffers
// and flags of the halo exchange, of the small all-gathers of the partitioned line solver and of the Krylov
// inner-product reductions live in that heap, so the data path is plain stores to peer memory issued by this
// library's own kernels:
//   push  : the producer writes its contribution into the consumer's receive slot, __threadfence_system(),
//           then a release-store of the epoch into the consumer's flag;
//   wait  : the consumer spins on its LOCAL flag with acquire loads, then reads its local slot.
// Slots and flags are double buffered by the parity of a device-resident epoch: a rank can run at most one
// exchange ahead of a peer (it needs the peer's data of exchange e to finish e), so two parities suffice and
// nothing has to be acknowledged.  The epoch lives on the device and is advanced by the kernels themselves,
// which makes every exchange CUDA-graph capturable (the multigrid cycle is replayed as one graph).
// NCCL remains for bootstrap (unique id, exchange of IPC handles and heap offsets) and as the fallback when
// peer mapping is unavailable (FDFD_P2P=0 forces the fallback).
#pragma once
#include "common.cuh"

namespace fdfd {

struct Comm;

constexpr int kP2PMaxRanks = 16;
constexpr int kP2PRedMax = 512;         // doubles per all-reduce
constexpr int kP2PGatherCtas = 32;      // CTAs (and flags per source rank) of an all-gather

// ghost-row exchange of one sharded operator (device pointers; see helm2d.cu)
struct HaloP2P {
    void* push_north = nullptr;                 // north neighbour's rx_south  [2][Nx] elements
    void* push_south = nullptr;                 // south neighbour's rx_north
    unsigned int* flag_push_north = nullptr;    // north neighbour's flag_south [2][segs]
    unsigned int* flag_push_south = nullptr;    // south neighbour's flag_north
    const void* rx_south = nullptr;             // local receive rows [2][Nx] (16 bytes per element reserved)
    const void* rx_north = nullptr;
    const unsigned int* flag_south = nullptr;   // local flags [2][segs]; null = no neighbour on that side
    const unsigned int* flag_north = nullptr;
    unsigned int* state = nullptr;              // local {epoch, done counter}
    int segs = 0;                               // flags per parity (column segments of 32)
};

struct GatherPlan {
    int nranks = 1, rank = 0;
    char* rx_peer[kP2PMaxRanks];                // receive region of every rank, [2][nranks][cap] bytes
    unsigned int* flag_peer[kP2PMaxRanks];      // flags of every rank, [2][nranks][kP2PGatherCtas]
    unsigned int* state = nullptr;              // local {epoch, done counter}
    size_t cap = 0;                             // bytes per source-rank slot
    size_t off_rx = 0, off_flags = 0;           // local heap offsets
};

int p2p_init(Comm* c);                          // leaves c->p2p null when peer mapping is not possible
void p2p_shutdown(Comm* c);
bool p2p_enabled(const Comm* c);
int p2p_alloc(Comm* c, size_t bytes, size_t* off);      // local heap, 256-byte aligned, zero-filled
void p2p_free(Comm* c, size_t off);
void* p2p_local(Comm* c, size_t off);
// collective: every rank contributes k offsets into its own heap; ptrs[r * k + i] = that location of rank r
// mapped into this process
int p2p_exchange(Comm* c, const size_t* mine, int k, void** ptrs, cudaStream_t st);

int p2p_gather_create(Comm* c, size_t cap_bytes_per_rank, GatherPlan** out, cudaStream_t st);
void p2p_gather_destroy(Comm* c, GatherPlan* p);
// recv[r * bytes ...] <- `send` of rank r (bytes <= cap, multiple of 8); recv is ordinary local memory
int p2p_allgather(GatherPlan* p, const void* send, void* recv, size_t bytes, cudaStream_t st);
// in-place sum over ranks of count <= kP2PRedMax doubles, summed in rank order on every rank (bitwise
// identical results everywhere)
int p2p_allreduce_sum(Comm* c, double* buf, int count, cudaStream_t st);

// ghost rows of a sharded operator: buffers + flags in the heap, pointers of both neighbours resolved
int p2p_halo_create(Comm* c, int Nx, int south, int north, HaloP2P* out, size_t offs[4], cudaStream_t st);
void p2p_halo_destroy(Comm* c, HaloP2P* h, const size_t offs[4]);
// stand-alone exchange (for consumers other than the stencil kernel): pushes the first / last row of x
// (Ny rows of Nx elements of `elem` bytes) and copies the received rows into ghost_south / ghost_north
int p2p_halo_exchange(const HaloP2P& h, const void* x, void* ghost_south, void* ghost_north, int Nx, int Ny,
                      size_t elem, cudaStream_t st);

}  // namespace fdfd
