// Building blocks of the persistent (one launch = one whole solve) kernels: generation barrier among the
// co-resident CTAs of a cooperative launch and deterministic grid-wide sums.
#pragma once
#include "common.cuh"

namespace fdfd {

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Generation barrier among `count` CTAs.  bar[0] = arrival counter, bar[1] = generation (monotonic over
// the life of the buffer; every CTA read it at kernel start, before anyone could have advanced it).
// Release on arrival (cumulative over the CTA's writes ordered by the bar.sync), acquire on departure.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int count, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gen;
        unsigned int prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == count - 1) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
            st_release_u32(&bar[1], gen);
        } else {
            while (ld_acquire_u32(&bar[1]) != gen) { }
        }
    }
    __syncthreads();
}

// block-wide sum of K doubles -> this CTA's row of the partials array (call before the barrier);
// sm: K * (threads / 32) doubles of shared memory
template <int K, int THREADS>
__device__ __forceinline__ void block_partials(double (&v)[K], double* row, double* sm) {
    constexpr int W = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
        if (lane == 0) sm[k * W + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) s += sm[threadIdx.x * W + w];
        row[threadIdx.x] = s;
    }
}

// after the barrier: every CTA sums the K partials of all CTAs in the same order (identical bits everywhere)
template <int K>
__device__ __forceinline__ void grid_total(const double* partials, int stride, double (&out)[K], double* sm) {
    if (threadIdx.x < 32) {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)b * stride + k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], off);
            if (threadIdx.x == 0) sm[k] = acc[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = sm[k];
    __syncthreads();
}

}  // namespace fdfd
