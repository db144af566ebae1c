// K3 — Krylov solvers on top of the matrix-free operator (K2) and fused BLAS-1 (K4).
#pragma once
#include "blas1.cuh"
#include "helm2d.cuh"

namespace fdfd {

// z = M^-1 v.  Implementations: identity, Jacobi (1/diag), shifted-Laplacian multigrid (mg.cu).
struct Precond {
    virtual ~Precond() {}
    virtual int apply(const void* v, void* z, cudaStream_t stream) = 0;
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
int bicgstab_solve(Helm2D& A, Precond* M, const fdfd_krylov_opts& o, const void* b, void* x,
                   SolveInfo& info, cudaStream_t stream, BicgWorkspace* ws = nullptr,
                   bool fixed_its = false);
int fgmres_solve(Helm2D& A, Precond* M, const fdfd_krylov_opts& o, const void* b, void* x,
                 SolveInfo& info, cudaStream_t stream);

}  // namespace fdfd
