// K2b — device CSR operators for the full-vector mode solver.
//
// The reference forms Omega = P_red @ Q_red explicitly with scipy sparse products
// (Mode_Solver_2D/Mode_Solver_2D.py:752-774) and hands it to ARPACK/SuperLU.  Here the product is
// never formed: Omega x = P (Q x) is two sparse mat-vecs on the device (P, Q have <= 5 entries
// per row), and (Omega - sigma I)^-1 is applied by the K3 Krylov solver.
//
// Rows are short (2-9 entries) and of near-constant length: a group of kCsrLanes = 4 adjacent lanes owns a
// row, so a warp walks 8 consecutive rows and its loads of `data` / `indices` cover one contiguous span of
// the CSR arrays (a thread-per-row kernel strides through them with the row length, a warp-per-row kernel
// idles 27 lanes); the partial sums meet in two shuffle steps.  Algorithmic bytes per application:
// nnz * (sizeof(C) + 4) + rows * (4 + sizeof(C)) + cols * sizeof(C)  (fdfd_linop_bytes).
#include "krylov.cuh"

namespace fdfd {

constexpr int kCsrLanes = 4;

template <typename C>
__global__ void __launch_bounds__(256) csr_spmv_kernel(int nrows, const int32_t* __restrict__ indptr,
        const int32_t* __restrict__ indices, const C* __restrict__ data, const C* __restrict__ x,
        C* y, C shift, const C* xs /* y += shift * xs[row] when non-null */, const C* b /* y = b - y when non-null */) {
    using R = typename real_of<C>::type;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int row = (int)(t / kCsrLanes), sub = (int)(t % kCsrLanes);
    const bool live = row < nrows;
    R ax = 0, ay = 0;
    if (live) {
        const int e0 = indptr[row], e1 = indptr[row + 1];
        for (int e = e0 + sub; e < e1; e += kCsrLanes) {
            const C d = data[e], v = x[indices[e]];
            ax += d.x * v.x - d.y * v.y;
            ay += d.x * v.y + d.y * v.x;
        }
    }
#pragma unroll
    for (int off = kCsrLanes / 2; off > 0; off >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, off);
        ay += __shfl_xor_sync(0xffffffffu, ay, off);
    }
    if (live && sub == 0) {
        C acc = mk<C>(ax, ay);
        if (xs) acc = cfma(shift, xs[row], acc);
        if (b) acc = csub(b[row], acc);
        y[row] = acc;
    }
}

static inline int csr_grid(int nrows) { return ceil_div((int64_t)nrows * kCsrLanes, 256); }

struct CsrDev {
    int nrows = 0, ncols = 0;
    const int32_t* indptr = nullptr;
    const int32_t* indices = nullptr;
    const void* data = nullptr;
};

// y = P (Q x) + shift * x      (P: n x m, Q: m x n)
struct CsrProductOp : LinOp {
    CsrDev P, Q;
    double shift_re = 0.0, shift_im = 0.0;
    void* tmp = nullptr;   // m elements
    ~CsrProductOp() override { cudaFree(tmp); }
    int64_t n_local() const override { return P.nrows; }

    template <typename C>
    int run(const C* x, C* y, const C* b, cudaStream_t st) {
        using R = typename real_of<C>::type;
        csr_spmv_kernel<C><<<csr_grid(Q.nrows), 256, 0, st>>>(Q.nrows, Q.indptr, Q.indices, (const C*)Q.data, x,
                                                                   (C*)tmp, czero<C>(), nullptr, nullptr);
        csr_spmv_kernel<C><<<csr_grid(P.nrows), 256, 0, st>>>(P.nrows, P.indptr, P.indices, (const C*)P.data,
                                                                   (const C*)tmp, y, mk<C>((R)shift_re, (R)shift_im), x, b);
        FDFD_CUDA(cudaGetLastError());
        return FDFD_OK;
    }
    int apply(const void* x, void* y, cudaStream_t st) override {
        if (dtype == FDFD_C128) return run<double2>((const double2*)x, (double2*)y, nullptr, st);
        return run<float2>((const float2*)x, (float2*)y, nullptr, st);
    }
    int residual(const void* x, void* r, const void* b, cudaStream_t st) override {
        if (dtype == FDFD_C128) return run<double2>((const double2*)x, (double2*)r, (const double2*)b, st);
        return run<float2>((const float2*)x, (float2*)r, (const float2*)b, st);
    }
};

// y = A x + shift * x for a square device CSR matrix (3-D pencil matrices A, B; K7 residual bases)
struct CsrOp : LinOp {
    CsrDev A;
    double shift_re = 0.0, shift_im = 0.0;
    int64_t n_local() const override { return A.nrows; }
    template <typename C>
    int run(const C* x, C* y, const C* b, cudaStream_t st) {
        using R = typename real_of<C>::type;
        const bool sh = shift_re != 0.0 || shift_im != 0.0;
        csr_spmv_kernel<C><<<csr_grid(A.nrows), 256, 0, st>>>(A.nrows, A.indptr, A.indices, (const C*)A.data, x, y,
                                                                   mk<C>((R)shift_re, (R)shift_im), sh ? x : nullptr, b);
        FDFD_CUDA(cudaGetLastError());
        return FDFD_OK;
    }
    int apply(const void* x, void* y, cudaStream_t st) override {
        if (dtype == FDFD_C128) return run<double2>((const double2*)x, (double2*)y, nullptr, st);
        return run<float2>((const float2*)x, (float2*)y, nullptr, st);
    }
    int residual(const void* x, void* r, const void* b, cudaStream_t st) override {
        if (dtype == FDFD_C128) return run<double2>((const double2*)x, (double2*)r, (const double2*)b, st);
        return run<float2>((const float2*)x, (float2*)r, (const float2*)b, st);
    }
};

}  // namespace fdfd

using namespace fdfd;

struct fdfd_linop { LinOp* op; };

extern "C" int fdfd_csr_create(fdfd_linop_t** out, int dtype, int n, const int32_t* indptr,
                               const int32_t* indices, const void* data, double shift_re, double shift_im) {
    FDFD_CHECK_ARG(out && indptr && indices && data, "null pointer");
    FDFD_CHECK_ARG(dtype == FDFD_C128 || dtype == FDFD_C64, "dtype must be FDFD_C128 or FDFD_C64");
    FDFD_CHECK_ARG(n >= 1, "bad shape");
    if (fdfd_device_count() <= 0) { set_error("no CUDA device visible: fdfd_b200 has no CPU path"); return FDFD_ERR_CUDA; }
    CsrOp* op = new CsrOp();
    op->dtype = dtype;
    op->A.nrows = n; op->A.ncols = n; op->A.indptr = indptr; op->A.indices = indices; op->A.data = data;
    op->shift_re = shift_re; op->shift_im = shift_im;
    *out = new fdfd_linop{op};
    return FDFD_OK;
}

extern "C" int fdfd_csr_create_rect(fdfd_linop_t** out, int dtype, int nrows, int ncols, const int32_t* indptr,
                                    const int32_t* indices, const void* data) {
    FDFD_CHECK_ARG(out && indptr && indices && data, "null pointer");
    FDFD_CHECK_ARG(dtype == FDFD_C128 || dtype == FDFD_C64, "dtype must be FDFD_C128 or FDFD_C64");
    FDFD_CHECK_ARG(nrows >= 1 && ncols >= 1, "bad shape");
    if (fdfd_device_count() <= 0) { set_error("no CUDA device visible: fdfd_b200 has no CPU path"); return FDFD_ERR_CUDA; }
    CsrOp* op = new CsrOp();
    op->dtype = dtype;
    op->A.nrows = nrows; op->A.ncols = ncols; op->A.indptr = indptr; op->A.indices = indices; op->A.data = data;
    *out = new fdfd_linop{op};
    return FDFD_OK;
}

extern "C" int fdfd_csr_product_create(fdfd_linop_t** out, int dtype, int n, int m,
                                       const int32_t* P_indptr, const int32_t* P_indices, const void* P_data,
                                       const int32_t* Q_indptr, const int32_t* Q_indices, const void* Q_data,
                                       double shift_re, double shift_im) {
    FDFD_CHECK_ARG(out && P_indptr && P_indices && P_data && Q_indptr && Q_indices && Q_data, "null pointer");
    FDFD_CHECK_ARG(dtype == FDFD_C128 || dtype == FDFD_C64, "dtype must be FDFD_C128 or FDFD_C64");
    FDFD_CHECK_ARG(n >= 1 && m >= 1, "bad shape");
    if (fdfd_device_count() <= 0) { set_error("no CUDA device visible: fdfd_b200 has no CPU path"); return FDFD_ERR_CUDA; }
    CsrProductOp* op = new CsrProductOp();
    op->dtype = dtype;
    op->P.nrows = n; op->P.ncols = m; op->P.indptr = P_indptr; op->P.indices = P_indices; op->P.data = P_data;
    op->Q.nrows = m; op->Q.ncols = n; op->Q.indptr = Q_indptr; op->Q.indices = Q_indices; op->Q.data = Q_data;
    op->shift_re = shift_re; op->shift_im = shift_im;
    cudaError_t e = cudaMalloc(&op->tmp, (size_t)m * op->elem());
    if (e != cudaSuccess) { delete op; return cuda_fail(e, "cudaMalloc(csr product tmp)", __FILE__, __LINE__); }
    *out = new fdfd_linop{op};
    return FDFD_OK;
}

extern "C" void fdfd_linop_destroy(fdfd_linop_t* h) {
    if (!h) return;
    delete h->op;
    delete h;
}

extern "C" int fdfd_linop_apply(fdfd_linop_t* h, const void* x, void* y, void* stream) {
    FDFD_CHECK_ARG(h && x && y && x != y, "bad pointers");
    return h->op->apply(x, y, (cudaStream_t)stream);
}

extern "C" int fdfd_linop_solve(fdfd_linop_t* h, const fdfd_krylov_opts* opts, const void* dinv, const void* b,
                                void* x, double* info_out, void* stream) {
    FDFD_CHECK_ARG(h && opts && b && x, "null pointer");
    FDFD_CHECK_ARG(opts->rtol > 0 && opts->maxit >= 0, "bad tolerance / maxit");
    LinOp& A = *h->op;
    cudaStream_t st = (cudaStream_t)stream;
    Precond* M = nullptr;
    if (dinv) FDFD_TRY(make_diag_precond(A.dtype, A.n_local(), dinv, &M));
    SolveInfo info;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    int s;
    if (opts->method == FDFD_KRYLOV_BICGSTAB) s = bicgstab_solve(A, M, *opts, b, x, info, st);
    else if (opts->method == FDFD_KRYLOV_FGMRES) s = fgmres_solve(A, M, *opts, b, x, info, st);
    else { set_error("unknown Krylov method %d", opts->method); s = FDFD_ERR_ARG; }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    delete M;
    if (info_out) {
        info_out[0] = info.iterations; info_out[1] = info.relres; info_out[2] = (double)info.op_applies;
        info_out[3] = (double)info.precond_applies; info_out[4] = ms * 1e-3; info_out[5] = info.breakdowns;
        info_out[6] = 0; info_out[7] = 0;
    }
    return s;
}
