// K4 — fused BLAS-1 kernels (see blas1.cuh).  Streaming, bandwidth bound: every vector is read
// or written exactly once per kernel with unit-stride 16-byte (c128) / 8-byte (c64) accesses;
// grids are a multiple of the SM count and loop grid-stride.
#include "blas1.cuh"
#include "comm.cuh"

namespace fdfd {

int Scalars::alloc() {
    FDFD_CUDA(cudaMalloc(&slot, sizeof(double2) * kSlots));
    FDFD_CUDA(cudaMemset(slot, 0, sizeof(double2) * kSlots));
    FDFD_CUDA(cudaMalloc(&partials, sizeof(double) * 2 * kMaxBlocks * 64));
    FDFD_CUDA(cudaMalloc(&counter, sizeof(unsigned int)));
    FDFD_CUDA(cudaMemset(counter, 0, sizeof(unsigned int)));
    return FDFD_OK;
}
void Scalars::release() {
    cudaFree(slot); cudaFree(partials); cudaFree(counter);
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

template <typename C> __device__ __forceinline__ double2 to_d(C a) { return make_double2((double)a.x, (double)a.y); }
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

template <typename C>
__global__ void __launch_bounds__(kThreads) multi_dot_kernel(int64_t n, int k, const C* __restrict__ V,
        int64_t ldv, const C* __restrict__ w, double* out, double* partials, unsigned int* counter) {
    double acc[2 * kMB];
#pragma unroll
    for (int q = 0; q < 2 * kMB; ++q) acc[q] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 b = to_d(w[i]);
#pragma unroll
        for (int q = 0; q < kMB; ++q) {
            if (q < k) {
                const double2 a = to_d(V[(int64_t)q * ldv + i]);
                acc[2 * q] += a.x * b.x + a.y * b.y;
                acc[2 * q + 1] += a.x * b.y - a.y * b.x;
            }
        }
    }
    reduce_finalize<2 * kMB>(acc, partials, counter, out);
}

template <typename C>
int multi_dot(int64_t n, int k, const C* V, int64_t ldv, const C* w, int s0, Scalars& sc,
              cudaStream_t stream) {
    const int grid = stream_grid(n);
    for (int base = 0; base < k; base += kMB) {
        const int kk = (k - base) < kMB ? (k - base) : kMB;
        // the kernel writes 2*kMB doubles; keep them inside the slot array
        FDFD_CHECK_ARG(s0 + base + kMB <= Scalars::kSlots, "multi_dot: slot range");
        multi_dot_kernel<C><<<grid, kThreads, 0, stream>>>(n, kk, V + (int64_t)base * ldv, ldv, w,
                (double*)(sc.slot + s0 + base), sc.partials, sc.counter);
    }
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename C, bool NORM>
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(int64_t n, int k, const C* __restrict__ V,
        int64_t ldv, C* __restrict__ w, const double2* coef, double sign, double* out,
        double* partials, unsigned int* counter) {
    C h[kMB];
#pragma unroll
    for (int q = 0; q < kMB; ++q) h[q] = q < k ? from_d<C>(cscale(sign, coef[q])) : czero<C>();
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        C r = w[i];
#pragma unroll
        for (int q = 0; q < kMB; ++q)
            if (q < k) r = cfma(h[q], V[(int64_t)q * ldv + i], r);
        w[i] = r;
        if (NORM) acc[0] += (double)r.x * (double)r.x + (double)r.y * (double)r.y;
    }
    if (NORM) reduce_finalize<1>(acc, partials, counter, out);
}

__global__ void zero_im_kernel(double2* slot) { slot->y = 0.0; }

template <typename C>
int multi_axpy(int64_t n, int k, const C* V, int64_t ldv, C* w, int s0, double sign,
               int nrm_slot, Scalars& sc, cudaStream_t stream) {
    const int grid = stream_grid(n);
    if (k == 0 && nrm_slot >= 0) {
        const C* xs[1] = {w}; const C* ys[1] = {w};
        return dots<C>(n, 1, xs, ys, nrm_slot, sc, stream);
    }
    for (int base = 0; base < k; base += kMB) {
        const int kk = (k - base) < kMB ? (k - base) : kMB;
        const bool last = base + kMB >= k;
        if (last && nrm_slot >= 0) {
            multi_axpy_kernel<C, true><<<grid, kThreads, 0, stream>>>(n, kk, V + (int64_t)base * ldv,
                    ldv, w, sc.slot + s0 + base, sign, (double*)(sc.slot + nrm_slot), sc.partials, sc.counter);
            zero_im_kernel<<<1, 1, 0, stream>>>(sc.slot + nrm_slot);
        } else {
            multi_axpy_kernel<C, false><<<grid, kThreads, 0, stream>>>(n, kk, V + (int64_t)base * ldv,
                    ldv, w, sc.slot + s0 + base, sign, nullptr, sc.partials, sc.counter);
        }
    }
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename C>
__global__ void __launch_bounds__(kThreads) scale_inv_sqrt_kernel(int64_t n, C* __restrict__ out,
        const C* __restrict__ x, const double2* nrm2) {
    using R = typename real_of<C>::type;
    const R s = (R)(1.0 / sqrt(nrm2->x));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = cscale(s, x[i]);
}

template <typename C>
int scale_inv_sqrt(int64_t n, C* out, const C* x, int nrm2_slot, Scalars& sc, cudaStream_t stream) {
    scale_inv_sqrt_kernel<C><<<stream_grid(n), kThreads, 0, stream>>>(n, out, x, sc.slot + nrm2_slot);
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

template <typename Cin, typename Cout>
__global__ void __launch_bounds__(kThreads) convert_kernel(int64_t n, const Cin* __restrict__ in, Cout* __restrict__ out) {
    using R = typename real_of<Cout>::type;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const Cin v = in[i];
        out[i] = mk<Cout>((R)v.x, (R)v.y);
    }
}
template <typename Cin, typename Cout>
int convert(int64_t n, const Cin* in, Cout* out, cudaStream_t stream) {
    convert_kernel<Cin, Cout><<<stream_grid(n), kThreads, 0, stream>>>(n, in, out);
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
    template int multi_dot<C>(int64_t, int, const C*, int64_t, const C*, int, Scalars&, cudaStream_t); \
    template int multi_axpy<C>(int64_t, int, const C*, int64_t, C*, int, double, int, Scalars&, cudaStream_t); \
    template int scale_inv_sqrt<C>(int64_t, C*, const C*, int, Scalars&, cudaStream_t);                \
    template int pointwise_mul<C>(int64_t, C*, const C*, const C*, cudaStream_t);
INST(double2)
INST(float2)
#undef INST

}  // namespace fdfd
