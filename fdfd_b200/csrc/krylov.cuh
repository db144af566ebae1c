// K3 — Krylov solvers on top of the matrix-free operator (K2) and fused BLAS-1 (K4).
#pragma once
#include "blas1.cuh"
#include "helm2d.cuh"

namespace fdfd {

// Abstract linear operator the Krylov solvers iterate on: the matrix-free Helmholtz stencil (K2)
// or a product of device CSR matrices (mode-solver operator Omega = P Q, csr_ops.cu).
struct LinOp {
    int dtype = FDFD_C128;
    Comm* comm = nullptr;
    virtual ~LinOp() {}
    virtual int64_t n_local() const = 0;
    virtual int apply(const void* x, void* y, cudaStream_t stream) = 0;                         // y = A x
    virtual int residual(const void* x, void* r, const void* b, cudaStream_t stream) = 0;       // r = b - A x
    // y = A x with x stored in complex64 and a complex128 operator (flexible-GMRES corrections kept in
    // complex64); operators without such a kernel report false from has_apply_x32()
    virtual bool has_apply_x32() const { return false; }
    virtual int apply_x32(const float2*, double2*, cudaStream_t) { return FDFD_ERR_STATE; }
    size_t elem() const { return dtype == FDFD_C128 ? 16 : 8; }
};

struct HelmOp : LinOp {
    Helm2D& A;
    explicit HelmOp(Helm2D& a) : A(a) { dtype = a.dtype; comm = a.comm; }
    int64_t n_local() const override { return A.n_local(); }
    int apply(const void* x, void* y, cudaStream_t stream) override {
        return helm2d_apply(A, APPLY_AX, x, y, nullptr, 0.0, stream);
    }
    int residual(const void* x, void* r, const void* b, cudaStream_t stream) override {
        return helm2d_apply(A, APPLY_RESID, x, r, b, 0.0, stream);
    }
    bool has_apply_x32() const override { return A.dtype == FDFD_C128; }
    int apply_x32(const float2* x, double2* y, cudaStream_t stream) override { return helm2d_apply_x32(A, x, y, stream); }
};

// z = M^-1 v.  Implementations: identity, Jacobi (1/diag), shifted-Laplacian multigrid (mg.cu).
struct Precond {
    virtual ~Precond() {}
    virtual int apply(const void* v, void* z, cudaStream_t stream) = 0;
    // complex64 in / complex64 out interface of a preconditioner that runs in complex64 inside a complex128
    // solve (mixed-precision multigrid): single_input() is the buffer the caller may fill directly with the
    // complex64 vector (null = none); apply_single(null, z) then uses that buffer's content
    virtual bool single_io() const { return false; }
    virtual float2* single_input() { return nullptr; }
    virtual int apply_single(const float2*, float2*, cudaStream_t) { return FDFD_ERR_STATE; }
    int64_t applies = 0;
};

struct SolveInfo {
    int iterations = 0;
    double relres = 0.0;
    int64_t op_applies = 0;
    int64_t precond_applies = 0;
    double seconds = 0.0;
    int breakdowns = 0;
};

int make_jacobi_precond(Helm2D& A, Precond** out, cudaStream_t stream);
// z = d .* v with a caller-owned device array d (n elements of the operator dtype)
int make_diag_precond(int dtype, int64_t n, const void* d, Precond** out);
int make_mg_precond(Helm2D& A, const fdfd_krylov_opts& opts, Precond** out, cudaStream_t stream);

// scratch vectors + scalar slots of BiCGSTAB, reusable across solves (inner coarse-level solves
// of the multigrid cycle must not allocate)
struct BicgWorkspace {
    Scalars sc;
    void* v[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap = 0;
    int count = 0;
    int ensure(size_t bytes, bool precond);
    ~BicgWorkspace();
};

// fixed_its: run exactly o.maxit iterations without any host synchronisation (inner solves).
int bicgstab_solve(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const void* b, void* x,
                   SolveInfo& info, cudaStream_t stream, BicgWorkspace* ws = nullptr,
                   bool fixed_its = false);
// device memory of the (F)GMRES basis, reusable across solves with the same operator
struct KrylovArena {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    ~KrylovArena() { release(); }
};
int fgmres_solve(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const void* b, void* x,
                 SolveInfo& info, cudaStream_t stream, KrylovArena* arena = nullptr);

}  // namespace fdfd
