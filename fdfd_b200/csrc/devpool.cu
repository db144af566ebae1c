// Device-memory cache for the library's small work buffers.
//
// cudaMalloc / cudaFree are driver calls that synchronise the device; on this pool's B200s one of them costs
// 0.3 ms on a quiet context and up to 40 ms after cooperative launches or graph instantiation
// (profiles/r02_alloc_cost.txt).  The eigen-solver classes create and destroy operators, Krylov work vectors
// and scalar slots once per frequency / Bloch point, so those costs used to show up as 0.2-0.6 s per solve.
// Blocks of at most kMaxBlock bytes are therefore recycled: dev_free() parks a block in a per-device free
// list, dev_alloc() hands out the smallest parked block that fits without wasting more than half of it.
// Larger blocks (the vectors of the scattering solves: 0.27-4.3 GB each) go straight to the driver.
#include "common.cuh"

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace fdfd {

namespace {
constexpr size_t kMaxBlock = (size_t)64 << 20;
constexpr size_t kMaxCached = (size_t)1 << 30;

struct Pool {
    std::mutex mu;
    std::unordered_map<void*, std::pair<int, size_t>> live;            // pooled blocks handed out: device, size
    std::map<std::pair<int, size_t>, std::vector<void*>> parked;        // (device, size) -> free blocks
    size_t cached = 0;
    ~Pool() { /* process exit: the driver reclaims everything; calling cudaFree here can run after its teardown */ }
};
Pool& pool() { static Pool* p = new Pool(); return *p; }

size_t rounded(size_t bytes) {
    if (bytes < 512) return 512;
    if (bytes < ((size_t)1 << 20)) return (bytes + 511) & ~(size_t)511;
    return (bytes + 65535) & ~(size_t)65535;
}
}  // namespace

cudaError_t dev_alloc(void** out, size_t bytes) {
    *out = nullptr;
    if (bytes == 0) return cudaSuccess;
    if (bytes > kMaxBlock) return cudaMalloc(out, bytes);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t want = rounded(bytes);
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.parked.lower_bound({dev, want});
        while (it != P.parked.end() && it->first.first == dev && it->first.second <= 2 * want) {
            if (!it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                P.cached -= it->first.second;
                P.live[p] = {dev, it->first.second};
                *out = p;
                return cudaSuccess;
            }
            ++it;
        }
    }
    void* p = nullptr;
    e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        // out of memory with blocks parked: give them back to the driver and retry once
        dev_trim();
        cudaGetLastError();
        e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = {dev, want};
    *out = p;
    return cudaSuccess;
}

void dev_free(void* p) {
    if (!p) return;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it != P.live.end()) {
            const std::pair<int, size_t> key = it->second;
            P.live.erase(it);
            if (P.cached + key.second <= kMaxCached) {
                P.parked[key].push_back(p);
                P.cached += key.second;
                return;
            }
        }
    }
    cudaFree(p);           // not pooled (large block), or the cache is full
}

size_t dev_trim() {
    Pool& P = pool();
    std::vector<void*> drop;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        for (auto& kv : P.parked) for (void* p : kv.second) drop.push_back(p);
        P.parked.clear();
        bytes = P.cached;
        P.cached = 0;
    }
    for (void* p : drop) cudaFree(p);
    return bytes;
}

}  // namespace fdfd

// Returns the cached work buffers to the driver (they are otherwise kept for the next operator / solve);
// the return value is the number of bytes released.
extern "C" int64_t fdfd_trim_memory(void) { return (int64_t)fdfd::dev_trim(); }
