// K4 — fused BLAS-1 kernels (see blas1.cuh).  Streaming, bandwidth bound: every vector is read
// or written exactly once per kernel with unit-stride 16-byte (c128) / 8-byte (c64) accesses;
// grids are a multiple of the SM count and loop grid-stride.
#include "blas1.cuh"
#include "comm.cuh"

#include <cstdlib>

#include <cuda/ptx>

namespace fdfd {

int Scalars::alloc() {
    FDFD_CUDA(dev_alloc(&slot, sizeof(double2) * kSlots));
    FDFD_CUDA(cudaMemset(slot, 0, sizeof(double2) * kSlots));
    FDFD_CUDA(dev_alloc(&partials, sizeof(double) * 2 * kMaxBlocks * 64));
    FDFD_CUDA(dev_alloc(&counter, sizeof(unsigned int)));
    FDFD_CUDA(cudaMemset(counter, 0, sizeof(unsigned int)));
    return FDFD_OK;
}
void Scalars::release() {
    dev_free(slot); dev_free(partials); dev_free(counter);
    slot = nullptr; partials = nullptr; counter = nullptr;
}

constexpr int kThreads = 256;

static int stream_grid(int64_t n, int per_thread = 4) {
    int64_t b = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    if (b < 1) b = 1;
    if (b > Scalars::kMaxBlocks) b = Scalars::kMaxBlocks;
    return (int)b;
}

// Block-wide sum of K doubles per thread, then cross-block finalisation by the last block to
// arrive (ticket counter).  out[k] receives the total.  Deterministic for a fixed grid.
template <int K>
__device__ __forceinline__ void reduce_finalize(double (&v)[K], double* partials,
                                                unsigned int* counter, double* out) {
    __shared__ double sm[K][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
        if (lane == 0) sm[k][warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = lane < nwarps ? sm[k][lane] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if (lane == 0) partials[(size_t)blockIdx.x * K + k] = s;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)b * K + k]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], off);
        if (lane == 0) sm[k][warp] = acc[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = lane < nwarps ? sm[k][lane] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if (lane == 0) out[k] = s;
        }
        if (lane == 0) *counter = 0u;
    }
}

// (measured on B200: F2F.F64.F32 converts 16 values per clock per SM, DFMA issues 59 per clock per SM;
// an integer-instruction widening was tried and is slower than the hardware conversion)
template <typename C> __device__ __forceinline__ double2 to_d(C a) { return make_double2((double)a.x, (double)a.y); }
template <typename Cout, typename Cin> __device__ __forceinline__ Cout cvt(Cin a) {
    return mk<Cout>((typename real_of<Cout>::type)a.x, (typename real_of<Cout>::type)a.y);
}
template <typename C> __device__ __forceinline__ C from_d(double2 a) {
    return mk<C>((typename real_of<C>::type)a.x, (typename real_of<C>::type)a.y);
}

// ---------------------------------------------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(kThreads) lincomb_kernel(int64_t n, C* out,
        const double2* c0, const C* x0, const double2* c1, const C* x1, const double2* c2,
        const C* x2, int neg_mask) {
    C k0 = czero<C>(), k1 = czero<C>(), k2 = czero<C>();
    if (x0) { double2 t = c0 ? *c0 : make_double2(1.0, 0.0); if (neg_mask & 1) t = cneg(t); k0 = from_d<C>(t); }
    if (x1) { double2 t = c1 ? *c1 : make_double2(1.0, 0.0); if (neg_mask & 2) t = cneg(t); k1 = from_d<C>(t); }
    if (x2) { double2 t = c2 ? *c2 : make_double2(1.0, 0.0); if (neg_mask & 4) t = cneg(t); k2 = from_d<C>(t); }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        C r = czero<C>();
        if (x0) r = cmul(k0, x0[i]);
        if (x1) r = cfma(k1, x1[i], r);
        if (x2) r = cfma(k2, x2[i], r);
        out[i] = r;
    }
}

// The finalising block writes straight into the slots: wrap reduce_finalize with a small adaptor.
template <typename C>
__global__ void __launch_bounds__(kThreads) lincomb_dots_kernel(int64_t n, C* out,
        const double2* c0, const C* x0, const double2* c1, const C* x1, const double2* c2,
        const C* x2, int neg_mask, const C* u, double* scratch_out, double* partials,
        unsigned int* counter) {
    C k0 = czero<C>(), k1 = czero<C>(), k2 = czero<C>();
    if (x0) { double2 t = c0 ? *c0 : make_double2(1.0, 0.0); if (neg_mask & 1) t = cneg(t); k0 = from_d<C>(t); }
    if (x1) { double2 t = c1 ? *c1 : make_double2(1.0, 0.0); if (neg_mask & 2) t = cneg(t); k1 = from_d<C>(t); }
    if (x2) { double2 t = c2 ? *c2 : make_double2(1.0, 0.0); if (neg_mask & 4) t = cneg(t); k2 = from_d<C>(t); }
    double acc[3] = {0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        C r = czero<C>();
        if (x0) r = cmul(k0, x0[i]);
        if (x1) r = cfma(k1, x1[i], r);
        if (x2) r = cfma(k2, x2[i], r);
        out[i] = r;
        const double2 rd = to_d(r);
        if (u) {
            const double2 ud = to_d(u[i]);
            acc[0] += ud.x * rd.x + ud.y * rd.y;
            acc[1] += ud.x * rd.y - ud.y * rd.x;
        }
        acc[2] += rd.x * rd.x + rd.y * rd.y;
    }
    reduce_finalize<3>(acc, partials, counter, scratch_out);
}

// copies the 3 scratch doubles into the requested slots (runs after the reduction kernel)
__global__ void scatter3_kernel(const double* scratch, double2* slots, int dot_u_slot, int dot_oo_slot) {
    if (dot_u_slot >= 0) slots[dot_u_slot] = make_double2(scratch[0], scratch[1]);
    if (dot_oo_slot >= 0) slots[dot_oo_slot] = make_double2(scratch[2], 0.0);
}

template <typename C>
int lincomb(int64_t n, C* out, const double2* c0, const C* x0, const double2* c1, const C* x1,
            const double2* c2, const C* x2, int neg_mask, const C* u, int dot_u_slot,
            int dot_oo_slot, Scalars& sc, cudaStream_t stream) {
    const int grid = stream_grid(n);
    if (dot_u_slot < 0 && dot_oo_slot < 0) {
        lincomb_kernel<C><<<grid, kThreads, 0, stream>>>(n, out, c0, x0, c1, x1, c2, x2, neg_mask);
    } else {
        double* scratch = sc.partials + (size_t)Scalars::kMaxBlocks * 64;  // tail of the buffer
        lincomb_dots_kernel<C><<<grid, kThreads, 0, stream>>>(n, out, c0, x0, c1, x1, c2, x2,
                neg_mask, dot_u_slot >= 0 ? u : nullptr, scratch, sc.partials, sc.counter);
        scatter3_kernel<<<1, 1, 0, stream>>>(scratch, sc.slot, dot_u_slot, dot_oo_slot);
    }
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename C, int K>
struct DotPtrs { const C* x[K]; const C* y[K]; };

template <typename C, int K>
__global__ void __launch_bounds__(kThreads) dots_kernel(int64_t n, DotPtrs<C, K> p, int ndots,
        double* out, double* partials, unsigned int* counter) {
    double acc[2 * K];
#pragma unroll
    for (int k = 0; k < 2 * K; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < ndots) {
                const double2 a = to_d(p.x[k][i]);
                const double2 b = (p.y[k] == p.x[k]) ? a : to_d(p.y[k][i]);
                acc[2 * k] += a.x * b.x + a.y * b.y;
                acc[2 * k + 1] += a.x * b.y - a.y * b.x;
            }
        }
    }
    reduce_finalize<2 * K>(acc, partials, counter, out);
}

template <typename C>
int dots(int64_t n, int ndots, const C* const* xs, const C* const* ys, int s0, Scalars& sc,
         cudaStream_t stream) {
    FDFD_CHECK_ARG(ndots >= 1 && ndots <= 4, "dots: 1..4 inner products per launch");
    const int grid = stream_grid(n);
    // slots are double2 = consecutive (re, im) doubles, so write the totals in place; the kernel
    // is instantiated per count so that exactly `ndots` slots are written
#define LAUNCH_DOTS(K)                                                                            \
    do {                                                                                          \
        DotPtrs<C, K> p;                                                                          \
        for (int k = 0; k < K; ++k) { p.x[k] = xs[k]; p.y[k] = ys[k]; }                           \
        dots_kernel<C, K><<<grid, kThreads, 0, stream>>>(n, p, ndots, (double*)(sc.slot + s0),   \
                                                          sc.partials, sc.counter);               \
    } while (0)
    switch (ndots) {
        case 1: LAUNCH_DOTS(1); break;
        case 2: LAUNCH_DOTS(2); break;
        case 3: LAUNCH_DOTS(3); break;
        default: LAUNCH_DOTS(4); break;
    }
#undef LAUNCH_DOTS
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

// ---------------------------------------------------------------------------------------------
constexpr int kMB = 8;  // vectors per multi-vector pass

// Two consecutive elements per thread and loop trip: a complex64 pair is one 16-byte load, a
// complex128 pair two of them.  With one 8-byte load per vector and trip the complex64 kernels
// kept too few bytes in flight (measured 2.8-3.7 TB/s against 5.6-5.8 TB/s for complex128).
template <typename C> struct Pair { C a, b; };
__device__ __forceinline__ Pair<float2> ld_pair(const float2* p, int64_t i2) {
    const float4 v = reinterpret_cast<const float4*>(p)[i2];
    return Pair<float2>{make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ Pair<double2> ld_pair(const double2* p, int64_t i2) {
    return Pair<double2>{p[2 * i2], p[2 * i2 + 1]};
}
__device__ __forceinline__ void st_pair(float2* p, int64_t i2, float2 a, float2 b) {
    reinterpret_cast<float4*>(p)[i2] = make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_pair(double2* p, int64_t i2, double2 a, double2 b) {
    p[2 * i2] = a; p[2 * i2 + 1] = b;
}
static inline bool pair_ok(const void* V, int64_t ldv, const void* w) {
    return ((uintptr_t)V % 16 == 0) && ((uintptr_t)w % 16 == 0) && (ldv % 2 == 0);
}

// four fused multiply-adds per complex element (the separate multiply / add form costs six FP64
// instructions; these kernels sit close to the FP64 issue limit once the basis is complex64)
__device__ __forceinline__ void dot_acc(double& re, double& im, double2 a, double2 b) {
    re = fma(a.x, b.x, re); re = fma(a.y, b.y, re);
    im = fma(a.x, b.y, im); im = fma(-a.y, b.x, im);
}
template <typename C> __device__ __forceinline__ C cfma4(C b, C c, C a) {
    using R = typename real_of<C>::type;
    R x = fma(b.x, c.x, a.x); x = fma(-b.y, c.y, x);
    R y = fma(b.x, c.y, a.y); y = fma(b.y, c.x, y);
    return mk<C>(x, y);
}

template <typename CV, typename C, int U>
__global__ void __launch_bounds__(kThreads) multi_dot_kernel(int64_t n, int k, const CV* __restrict__ V,
        int64_t ldv, const C* __restrict__ w, double* out, double* partials, unsigned int* counter) {
    double acc[2 * kMB];
#pragma unroll
    for (int q = 0; q < 2 * kMB; ++q) acc[q] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (U == 2) {
        const int64_t n2 = n >> 1;
        for (int64_t i = tid; i < n2; i += stride) {
            const Pair<C> b = ld_pair(w, i);
            const double2 b0 = to_d(b.a), b1 = to_d(b.b);
#pragma unroll
            for (int q = 0; q < kMB; ++q) {
                if (q < k) {
                    const Pair<CV> a = ld_pair(V + (int64_t)q * ldv, i);
                    dot_acc(acc[2 * q], acc[2 * q + 1], to_d(a.a), b0);
                    dot_acc(acc[2 * q], acc[2 * q + 1], to_d(a.b), b1);
                }
            }
        }
        if ((n & 1) && tid == 0) {
            const double2 b = to_d(w[n - 1]);
#pragma unroll
            for (int q = 0; q < kMB; ++q)
                if (q < k) dot_acc(acc[2 * q], acc[2 * q + 1], to_d(V[(int64_t)q * ldv + n - 1]), b);
        }
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            const double2 b = to_d(w[i]);
#pragma unroll
            for (int q = 0; q < kMB; ++q)
                if (q < k) dot_acc(acc[2 * q], acc[2 * q + 1], to_d(V[(int64_t)q * ldv + i]), b);
        }
    }
    reduce_finalize<2 * kMB>(acc, partials, counter, out);
}

// ---- bulk-copy (TMA) staged variants -------------------------------------------------------------
// The register kernels above keep (resident threads) x (loads per trip) bytes in flight; with the
// 74-84 registers the eight accumulator pairs need, that is ~120 KB per SM for a complex128 basis
// (5.6-5.9 TB/s) and half of it for a complex64 basis (3.5 TB/s — no gain from the smaller basis).
// Here one persistent CTA per SM streams tiles of w and of up to eight basis vectors into shared
// memory with cp.async.bulk (completion on an mbarrier, no registers tied up by loads in flight):
// up to ~200 KB per SM are in flight whatever the element type, and the threads only do LDS + DFMA.
namespace ptx = cuda::ptx;
constexpr int kStreamThreads = 512;
constexpr int kMaxStages = 16;
constexpr size_t kStreamSmem = 200 * 1024;

struct StreamPlan { int tile, stages, threads, grid; size_t smem; };
template <typename CV, typename C>
static StreamPlan stream_plan(int k) {
    // tile (elements per vector per stage): about 36-72 KB per stage so that a stage lasts ~1 us per SM.
    // A complex64 basis runs two CTAs per SM (256 threads, half the shared memory each) so that one CTA
    // computes while the other sits in its per-tile barrier; the complex128 stages are too large for that.
    const int ctas = sizeof(CV) == sizeof(float2) ? 2 : 1;
    int tile = k >= 5 ? 512 : (k >= 2 ? 1024 : 2048);
    const size_t stage = (size_t)tile * (sizeof(C) + (size_t)k * sizeof(CV));
    int stages = (int)(kStreamSmem / ctas / stage);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 1) stages = 1;
    return StreamPlan{tile, stages, kStreamThreads / ctas, kNumSMs * ctas, stage * stages};
}

template <typename CV, typename C>
struct StreamPipe {
    unsigned char* smem;
    uint64_t* full;
    int ntiles, my;           // full tiles of the vector; tiles of this CTA
    int k, tile, stages;
    uint32_t wbytes, vbytes, stage_bytes;
    const CV* V; int64_t ldv; const C* w;
    // consumer / producer cursors kept incrementally (no integer divisions on the per-tile path)
    int cs, pi, ps;
    uint32_t cpar;

    __device__ __forceinline__ int64_t tile_elem0(int i) const { return ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * tile; }
    __device__ __forceinline__ C* w_tile(int s) const { return (C*)(smem + (uint32_t)s * stage_bytes); }
    __device__ __forceinline__ CV* v_tile(int s, int q) const {
        return (CV*)(smem + (uint32_t)s * stage_bytes + wbytes + (uint32_t)q * vbytes);
    }
    __device__ __forceinline__ void issue_next() {            // one elected thread
        const int64_t e0 = tile_elem0(pi);
        ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &full[ps], stage_bytes);
        ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, w_tile(ps), w + e0, wbytes, &full[ps]);
        for (int q = 0; q < k; ++q)
            ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, v_tile(ps, q), V + (int64_t)q * ldv + e0, vbytes, &full[ps]);
        ++pi;
        if (++ps == stages) ps = 0;
    }
    __device__ __forceinline__ void init(unsigned char* sm, uint64_t* bars, int64_t n, int k_, int tile_, int stages_,
                                         const CV* V_, int64_t ldv_, const C* w_) {
        smem = sm; full = bars; k = k_; tile = tile_; stages = stages_; V = V_; ldv = ldv_; w = w_;
        wbytes = (uint32_t)(tile * sizeof(C)); vbytes = (uint32_t)(tile * sizeof(CV));
        stage_bytes = wbytes + (uint32_t)k * vbytes;
        ntiles = (int)(n / tile);
        my = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        cs = 0; cpar = 0; pi = 0; ps = 0;
        if (threadIdx.x == 0) {
            for (int s = 0; s < stages; ++s) ptx::mbarrier_init(&full[s], 1);
            ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
            ptx::fence_proxy_async(ptx::space_shared);
        }
        __syncthreads();
        if (threadIdx.x == 0)
            while (pi < my && pi < stages) issue_next();
    }
    // wait for the next tile of this CTA; returns its stage
    __device__ __forceinline__ int wait() const {
        while (!ptx::mbarrier_try_wait_parity(&full[cs], cpar)) {}
        return cs;
    }
    // every thread is done with the current stage: refill it with the next tile not yet requested
    __device__ __forceinline__ void release() {
        __syncthreads();
        if (threadIdx.x == 0 && pi < my) issue_next();
        if (++cs == stages) { cs = 0; cpar ^= 1u; }
    }
};

template <typename CV, typename C>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_dot_kernel(int64_t n, int k, const CV* __restrict__ V,
        int64_t ldv, const C* __restrict__ w, int tile, int stages, double* out, double* partials, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char stream_smem[];
    __shared__ uint64_t bars[kMaxStages];
    StreamPipe<CV, C> p;
    p.init(stream_smem, bars, n, k, tile, stages, V, ldv, w);
    double acc[2 * kMB];
#pragma unroll
    for (int q = 0; q < 2 * kMB; ++q) acc[q] = 0.0;
    for (int i = 0; i < p.my; ++i) {
        const int s = p.wait();
        const C* ws = p.w_tile(s);
        const CV* vs = p.v_tile(s, 0);
        for (int e = threadIdx.x; e < tile; e += blockDim.x) {
            const double2 b = to_d(ws[e]);
#pragma unroll
            for (int q = 0; q < kMB; ++q)
                if (q < k) dot_acc(acc[2 * q], acc[2 * q + 1], to_d(vs[q * tile + e]), b);
        }
        p.release();
    }
    // ragged tail (< tile elements) straight from global memory, spread over the CTAs
    for (int64_t e = (int64_t)p.ntiles * tile + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const double2 b = to_d(w[e]);
#pragma unroll
        for (int q = 0; q < kMB; ++q)
            if (q < k) dot_acc(acc[2 * q], acc[2 * q + 1], to_d(V[(int64_t)q * ldv + e]), b);
    }
    reduce_finalize<2 * kMB>(acc, partials, counter, out);
}

template <typename CV, typename C, bool NORM>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_axpy_kernel(int64_t n, int k, const CV* __restrict__ V,
        int64_t ldv, C* __restrict__ w, const double2* coef, double sign, int tile, int stages, double* out,
        double* partials, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char stream_smem[];
    __shared__ uint64_t bars[kMaxStages];
    StreamPipe<CV, C> p;
    p.init(stream_smem, bars, n, k, tile, stages, V, ldv, w);
    C h[kMB];
#pragma unroll
    for (int q = 0; q < kMB; ++q) h[q] = q < k ? from_d<C>(cscale(sign, coef[q])) : czero<C>();
    double acc[1] = {0.0};
    for (int i = 0; i < p.my; ++i) {
        const int s = p.wait();
        const C* ws = p.w_tile(s);
        const CV* vs = p.v_tile(s, 0);
        C* wg = w + p.tile_elem0(i);
        for (int e = threadIdx.x; e < tile; e += blockDim.x) {
            C r = ws[e];
#pragma unroll
            for (int q = 0; q < kMB; ++q)
                if (q < k) r = cfma4(h[q], cvt<C>(vs[q * tile + e]), r);
            wg[e] = r;
            if (NORM) acc[0] += (double)r.x * (double)r.x + (double)r.y * (double)r.y;
        }
        p.release();
    }
    for (int64_t e = (int64_t)p.ntiles * tile + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        C r = w[e];
#pragma unroll
        for (int q = 0; q < kMB; ++q)
            if (q < k) r = cfma4(h[q], cvt<C>(V[(int64_t)q * ldv + e]), r);
        w[e] = r;
        if (NORM) acc[0] += (double)r.x * (double)r.x + (double)r.y * (double)r.y;
    }
    if (NORM) reduce_finalize<1>(acc, partials, counter, out);
}

// vectors long enough to fill every SM with several stages go through the staged kernels
static inline bool use_stream(int64_t n, const void* V, int64_t ldv, size_t esz_v, const void* w) {
    static const bool off = getenv("FDFD_NO_STREAM_BLAS") != nullptr;
    return !off && n >= ((int64_t)1 << 20) && ((uintptr_t)V % 16 == 0) && ((uintptr_t)w % 16 == 0) &&
           ((ldv * (int64_t)esz_v) % 16 == 0);
}
template <typename K>
static int stream_attr(K kernel) {
    FDFD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamSmem));
    return FDFD_OK;
}

template <typename CV, typename C>
int multi_dot(int64_t n, int k, const CV* V, int64_t ldv, const C* w, int s0, Scalars& sc,
              cudaStream_t stream) {
    const bool pairs = pair_ok(V, ldv, w);
    const bool staged = use_stream(n, V, ldv, sizeof(CV), w);
    if (staged) FDFD_TRY(stream_attr(stream_dot_kernel<CV, C>));
    const int grid = stream_grid(pairs ? (n + 1) / 2 : n);
    for (int base = 0; base < k; base += kMB) {
        const int kk = (k - base) < kMB ? (k - base) : kMB;
        // the kernel writes 2*kMB doubles; keep them inside the slot array
        FDFD_CHECK_ARG(s0 + base + kMB <= Scalars::kSlots, "multi_dot: slot range");
        if (staged) {
            const StreamPlan pl = stream_plan<CV, C>(kk);
            stream_dot_kernel<CV, C><<<pl.grid, pl.threads, pl.smem, stream>>>(n, kk, V + (int64_t)base * ldv, ldv, w,
                    pl.tile, pl.stages, (double*)(sc.slot + s0 + base), sc.partials, sc.counter);
        } else if (pairs)
            multi_dot_kernel<CV, C, 2><<<grid, kThreads, 0, stream>>>(n, kk, V + (int64_t)base * ldv, ldv, w,
                    (double*)(sc.slot + s0 + base), sc.partials, sc.counter);
        else
            multi_dot_kernel<CV, C, 1><<<grid, kThreads, 0, stream>>>(n, kk, V + (int64_t)base * ldv, ldv, w,
                    (double*)(sc.slot + s0 + base), sc.partials, sc.counter);
    }
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename CV, typename C, bool NORM, int U>
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(int64_t n, int k, const CV* __restrict__ V,
        int64_t ldv, C* __restrict__ w, const double2* coef, double sign, double* out,
        double* partials, unsigned int* counter) {
    C h[kMB];
#pragma unroll
    for (int q = 0; q < kMB; ++q) h[q] = q < k ? from_d<C>(cscale(sign, coef[q])) : czero<C>();
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (U == 2) {
        const int64_t n2 = n >> 1;
        for (int64_t i = tid; i < n2; i += stride) {
            Pair<C> r = ld_pair(w, i);
#pragma unroll
            for (int q = 0; q < kMB; ++q) {
                if (q < k) {
                    const Pair<CV> a = ld_pair(V + (int64_t)q * ldv, i);
                    r.a = cfma4(h[q], cvt<C>(a.a), r.a);
                    r.b = cfma4(h[q], cvt<C>(a.b), r.b);
                }
            }
            st_pair(w, i, r.a, r.b);
            if (NORM) acc[0] += (double)r.a.x * (double)r.a.x + (double)r.a.y * (double)r.a.y +
                                (double)r.b.x * (double)r.b.x + (double)r.b.y * (double)r.b.y;
        }
        if ((n & 1) && tid == 0) {
            C r = w[n - 1];
#pragma unroll
            for (int q = 0; q < kMB; ++q)
                if (q < k) r = cfma4(h[q], cvt<C>(V[(int64_t)q * ldv + n - 1]), r);
            w[n - 1] = r;
            if (NORM) acc[0] += (double)r.x * (double)r.x + (double)r.y * (double)r.y;
        }
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            C r = w[i];
#pragma unroll
            for (int q = 0; q < kMB; ++q)
                if (q < k) r = cfma4(h[q], cvt<C>(V[(int64_t)q * ldv + i]), r);
            w[i] = r;
            if (NORM) acc[0] += (double)r.x * (double)r.x + (double)r.y * (double)r.y;
        }
    }
    if (NORM) reduce_finalize<1>(acc, partials, counter, out);
}

__global__ void zero_im_kernel(double2* slot) { slot->y = 0.0; }

template <typename CV, typename C>
int multi_axpy(int64_t n, int k, const CV* V, int64_t ldv, C* w, int s0, double sign,
               int nrm_slot, Scalars& sc, cudaStream_t stream) {
    if (k == 0 && nrm_slot >= 0) {
        const C* xs[1] = {w}; const C* ys[1] = {w};
        return dots<C>(n, 1, xs, ys, nrm_slot, sc, stream);
    }
    const bool pairs = pair_ok(V, ldv, w);
    const bool staged = use_stream(n, V, ldv, sizeof(CV), w);
    if (staged) {
        FDFD_TRY(stream_attr(stream_axpy_kernel<CV, C, true>));
        FDFD_TRY(stream_attr(stream_axpy_kernel<CV, C, false>));
    }
    const int grid = stream_grid(pairs ? (n + 1) / 2 : n);
    for (int base = 0; base < k; base += kMB) {
        const int kk = (k - base) < kMB ? (k - base) : kMB;
        const bool last = base + kMB >= k;
        const CV* Vb = V + (int64_t)base * ldv;
        const double2* cf = sc.slot + s0 + base;
        if (staged) {
            const StreamPlan pl = stream_plan<CV, C>(kk);
            const bool nrm = last && nrm_slot >= 0;
            double* o = nrm ? (double*)(sc.slot + nrm_slot) : nullptr;
            if (nrm) {
                stream_axpy_kernel<CV, C, true><<<pl.grid, pl.threads, pl.smem, stream>>>(n, kk, Vb, ldv, w, cf, sign,
                        pl.tile, pl.stages, o, sc.partials, sc.counter);
                zero_im_kernel<<<1, 1, 0, stream>>>(sc.slot + nrm_slot);
            } else {
                stream_axpy_kernel<CV, C, false><<<pl.grid, pl.threads, pl.smem, stream>>>(n, kk, Vb, ldv, w, cf, sign,
                        pl.tile, pl.stages, o, sc.partials, sc.counter);
            }
        } else if (last && nrm_slot >= 0) {
            double* o = (double*)(sc.slot + nrm_slot);
            if (pairs) multi_axpy_kernel<CV, C, true, 2><<<grid, kThreads, 0, stream>>>(n, kk, Vb, ldv, w, cf, sign, o, sc.partials, sc.counter);
            else multi_axpy_kernel<CV, C, true, 1><<<grid, kThreads, 0, stream>>>(n, kk, Vb, ldv, w, cf, sign, o, sc.partials, sc.counter);
            zero_im_kernel<<<1, 1, 0, stream>>>(sc.slot + nrm_slot);
        } else {
            if (pairs) multi_axpy_kernel<CV, C, false, 2><<<grid, kThreads, 0, stream>>>(n, kk, Vb, ldv, w, cf, sign, nullptr, sc.partials, sc.counter);
            else multi_axpy_kernel<CV, C, false, 1><<<grid, kThreads, 0, stream>>>(n, kk, Vb, ldv, w, cf, sign, nullptr, sc.partials, sc.counter);
        }
    }
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename C, typename C2>
__global__ void __launch_bounds__(kThreads) scale_inv_sqrt_kernel(int64_t n, C* __restrict__ out,
        C2* __restrict__ out2, C2* __restrict__ out3, const C* __restrict__ x, const double2* nrm2) {
    using R = typename real_of<C>::type;
    const R s = (R)(1.0 / sqrt(nrm2->x));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const C v = cscale(s, x[i]);
        if (out) out[i] = v;
        if (out2) { const C2 c = cvt<C2>(v); out2[i] = c; if (out3) out3[i] = c; }
    }
}

template <typename C>
int scale_inv_sqrt(int64_t n, C* out, const C* x, int nrm2_slot, Scalars& sc, cudaStream_t stream) {
    scale_inv_sqrt_kernel<C, C><<<stream_grid(n), kThreads, 0, stream>>>(n, out, (C*)nullptr, (C*)nullptr, x, sc.slot + nrm2_slot);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

int scale_inv_sqrt_dual(int64_t n, double2* out, float2* out_single, float2* out_single2, const double2* x,
                        int nrm2_slot, Scalars& sc, cudaStream_t stream) {
    scale_inv_sqrt_kernel<double2, float2><<<stream_grid(n), kThreads, 0, stream>>>(n, out, out_single, out_single2, x,
                                                                                     sc.slot + nrm2_slot);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename C>
__global__ void __launch_bounds__(kThreads) pointwise_mul_kernel(int64_t n, C* out, const C* x, const C* d) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = cmul(x[i], d[i]);
}
template <typename C>
int pointwise_mul(int64_t n, C* out, const C* x, const C* d, cudaStream_t stream) {
    pointwise_mul_kernel<C><<<stream_grid(n), kThreads, 0, stream>>>(n, out, x, d);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename Cin, typename Cout, int U>
__global__ void __launch_bounds__(kThreads) convert_kernel(int64_t n, const Cin* __restrict__ in, Cout* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (U == 2) {
        const int64_t n2 = n >> 1;
        for (int64_t i = tid; i < n2; i += stride) {
            const Pair<Cin> v = ld_pair(in, i);
            st_pair(out, i, cvt<Cout>(v.a), cvt<Cout>(v.b));
        }
        if ((n & 1) && tid == 0) out[n - 1] = cvt<Cout>(in[n - 1]);
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] = cvt<Cout>(in[i]);
    }
}
template <typename Cin, typename Cout>
int convert(int64_t n, const Cin* in, Cout* out, cudaStream_t stream) {
    if (pair_ok(in, 0, out)) convert_kernel<Cin, Cout, 2><<<stream_grid((n + 1) / 2), kThreads, 0, stream>>>(n, in, out);
    else convert_kernel<Cin, Cout, 1><<<stream_grid(n), kThreads, 0, stream>>>(n, in, out);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}
template int convert<double2, float2>(int64_t, const double2*, float2*, cudaStream_t);
template int convert<float2, double2>(int64_t, const float2*, double2*, cudaStream_t);

int allreduce_slots(Comm* comm, Scalars& sc, int s0, int count, cudaStream_t stream) {
    if (!comm || comm->nranks == 1) return FDFD_OK;
    return comm_allreduce_sum(comm, (double*)(sc.slot + s0), 2 * count, stream);
}

// a / b with 0 for an exactly-zero (or non-finite) denominator: a Krylov recurrence that has
// converged to machine zero then turns into a no-op instead of producing NaN
__device__ __forceinline__ double2 safe_div(double2 a, double2 b) {
    const double d = b.x * b.x + b.y * b.y;
    if (!(d > 0.0) || !isfinite(d)) return make_double2(0.0, 0.0);
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

__global__ void scalar_op_kernel(int op, int a, int b, int c, int d, int e, double p, double q, double2* s) {
    switch (op) {
        case SOP_SET: s[a] = make_double2(p, q); break;
        case SOP_DIV: s[a] = safe_div(s[b], s[c]); break;
        case SOP_BICG_BETA: {
            const double2 beta = cmul(safe_div(s[b], s[c]), safe_div(s[d], s[e]));
            s[a] = beta;
            s[a + 1] = cneg(cmul(beta, s[e]));
            break;
        }
        case SOP_COPY: s[a] = s[b]; break;
        case SOP_NEG: s[a] = cneg(s[b]); break;
    }
}

int scalar_op(int op, int a, int b, int c, int d, int e, double p, double q, Scalars& sc,
              cudaStream_t stream) {
    scalar_op_kernel<<<1, 1, 0, stream>>>(op, a, b, c, d, e, p, q, sc.slot);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

#define INST(C)                                                                                   \
    template int lincomb<C>(int64_t, C*, const double2*, const C*, const double2*, const C*,      \
                            const double2*, const C*, int, const C*, int, int, Scalars&, cudaStream_t); \
    template int dots<C>(int64_t, int, const C* const*, const C* const*, int, Scalars&, cudaStream_t); \
    template int multi_dot<C, C>(int64_t, int, const C*, int64_t, const C*, int, Scalars&, cudaStream_t); \
    template int multi_axpy<C, C>(int64_t, int, const C*, int64_t, C*, int, double, int, Scalars&, cudaStream_t); \
    template int scale_inv_sqrt<C>(int64_t, C*, const C*, int, Scalars&, cudaStream_t);                \
    template int pointwise_mul<C>(int64_t, C*, const C*, const C*, cudaStream_t);
INST(double2)
INST(float2)
#undef INST
// compressed Krylov basis: complex64 storage, complex128 arithmetic
template int multi_dot<float2, double2>(int64_t, int, const float2*, int64_t, const double2*, int, Scalars&, cudaStream_t);
template int multi_axpy<float2, double2>(int64_t, int, const float2*, int64_t, double2*, int, double, int, Scalars&, cudaStream_t);

}  // namespace fdfd
