/*
 * fdfd_b200 — C-ABI of the B200-native FDFD solve engine (libfdfd_b200.so).
 *
 * The reference (SolverNotConverging/FDFD) is pure Python: its "interface" for this path is a
 * set of in-process Python calls, cited per entry point below (file:line relative to the
 * reference root).  A reference maintainer binds this header with ctypes — see
 * INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative fdfd_status otherwise; the message of
 *     the last failure on the calling thread is available from fdfd_last_error().
 *   - *_host entry points take HOST pointers and do their own H2D/D2H copies;
 *     all other data pointers are caller-owned DEVICE pointers (cudaMalloc / torch tensors).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - flat index n = i + Nx*j (x fastest), exactly as the reference
 *     (Scattering_Solver_2D.py:41,61,107; Band_Diagram_Solver.py:286-291).
 *   - complex128 = interleaved (re, im) doubles, complex64 = interleaved floats.
 *   - no CPU fallback anywhere: without a CUDA device every compute call fails with
 *     FDFD_ERR_CUDA.
 */
#ifndef FDFD_B200_H
#define FDFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    FDFD_OK = 0,
    FDFD_ERR_ARG = -1,       /* bad argument (ValueError on the Python side)            */
    FDFD_ERR_CUDA = -2,      /* CUDA runtime failure / no device                        */
    FDFD_ERR_BREAKDOWN = -3, /* Krylov breakdown (rho or omega ~ 0, NaN)                 */
    FDFD_ERR_NOCONV = -4,    /* maxit reached (solution still returned, relres reported) */
    FDFD_ERR_NCCL = -5,
    FDFD_ERR_STATE = -6
} fdfd_status;

enum { FDFD_C128 = 0, FDFD_C64 = 1 };
enum { FDFD_POL_TE = 0, FDFD_POL_TM = 1 };       /* operator type, see fdfd_helm2d_create   */
enum { FDFD_BC_DIRICHLET = 0, FDFD_BC_BLOCH = 1 };/* yeeder2d BC codes (yee_derivative.py:16-19) */
enum { FDFD_KRYLOV_BICGSTAB = 0, FDFD_KRYLOV_FGMRES = 1 };
enum { FDFD_PRECOND_NONE = 0, FDFD_PRECOND_JACOBI = 1, FDFD_PRECOND_MG = 2 };

const char* fdfd_last_error(void);
int fdfd_version(void);
/* sizeof(fdfd_krylov_opts) as compiled into the library: bindings check their struct against it */
int fdfd_krylov_opts_size(void);
/* number of visible CUDA devices (0 when none) — the Python layer refuses to run on 0. */
int fdfd_device_count(void);
/* The library recycles its small device work buffers (<= 64 MB each, <= 1 GB in total) between operators and
 * solves instead of returning them to the driver; this releases them.  Returns the bytes released. */
int64_t fdfd_trim_memory(void);

/* ------------------------------------------------------------------------------------------
 * K1  Yee derivative-operator assembly  — replaces yeeder2d(NS, RES, BC, kinc)
 *     (Scattering/yee_derivative.py:6-74 == Band_Diagram_Solver/yee_derivative.py:6-74).
 *
 * Per axis the operator has at most three distinct values; they are evaluated on the host
 * with the reference's numpy expressions (bit parity of values) and passed in:
 *   diag  = -1/d            (or -1j*k when N==1,  yee_derivative.py:47-48)
 *   fwd   = +1/d
 *   wrap  = exp(-1j*k*N*d)/d   (Bloch axis only, yee_derivative.py:54-57 / :66-68)
 * An axis is complex128 iff N==1 or bc==BLOCH, else float64 — the same per-axis rule scipy's
 * type promotion produces.  Index arrays are int32 (scipy picks int32 while nnz < 2^31).
 * DHX/DHY (CSC) share the index arrays of DEX/DEY (CSR); data = -conj (yee_derivative.py:71-72).
 * ------------------------------------------------------------------------------------------ */
int fdfd_yee2d_csr_nnz(int Nx, int Ny, int bcx, int bcy,
                       int64_t* nnz_dex, int64_t* nnz_dey,
                       int* dex_is_complex, int* dey_is_complex);

/* axis: 0 -> DEX, 1 -> DEY.  vals = {diag_re, diag_im, fwd_re, fwd_im, wrap_re, wrap_im}.
 * e_data / h_data: float64[nnz] or complex128[nnz] per the rule above; h_data may be NULL. */
int fdfd_yee2d_csr_fill(int Nx, int Ny, int axis, int bc, const double vals[6],
                        int32_t* indptr, int32_t* indices, void* e_data, void* h_data,
                        void* stream);
/* same, HOST output buffers (device scratch + D2H inside). */
int fdfd_yee2d_csr_fill_host(int Nx, int Ny, int axis, int bc, const double vals[6],
                             int32_t* indptr, int32_t* indices, void* e_data, void* h_data);

/* ------------------------------------------------------------------------------------------
 * K1b  staggered-grid operators with two entries per row — replaces the Python loops of
 *      Mode_Solver_2D/Mode_Solver_2D.py:482-544 (_difference_matrix_x/_y between the Ez / Ex / Ey / Hz grids)
 *      and Periodic_Solver_3D/Periodic_Solver_3D.py:95-194 (_difference_matrix_x/_y, the periodic z difference
 *      _periodic_difference_matrix_z and the forward average _periodic_forward_average_z).
 *   Grids are (nx, ny, nz) with x fastest (2-D: nz = 1).  Row (i, j, k) of the OUTPUT grid holds
 *       (column (i, j, k) of the input grid, v_self)   and
 *       (column (i+1, j, k) | (i, j+1, k) | (i, j, (k+1) mod nz) for axis 0 | 1 | 2, v_fwd)
 *   in ascending column order, exactly what scipy's coo -> csr conversion produces; int32 index arrays,
 *   float64 data (2 per row).  v_self / v_fwd are evaluated by the caller with the reference's expressions
 *   (-1/d, +1/d or 0.5, 0.5).  h_data (optional) = -data: the CSC arrays of -D^H share indptr / indices.
 * ------------------------------------------------------------------------------------------ */
int fdfd_stagger_csr_fill(const int in_shape[3], const int out_shape[3], int axis, double v_self, double v_fwd,
                          int32_t* indptr, int32_t* indices, double* data, double* h_data, void* stream);
/* same, HOST output buffers */
int fdfd_stagger_csr_fill_host(const int in_shape[3], const int out_shape[3], int axis, double v_self, double v_fwd,
                               int32_t* indptr, int32_t* indices, double* data, double* h_data);

/* ------------------------------------------------------------------------------------------
 * Set-up of the scattering operator on the device (reference: add_object / add_UPML /
 * _build_A_* of Scattering_Solver_2D.py:72-153, :176-188): the three coefficient arrays from an object-id map,
 *   ca[q] = table[0][ids[q]], cb[q] = table[1][ids[q]], cd[q] = table[2][ids[q]]
 * ids: device uint8[n]; table_host: HOST complex128 [3][ntab] evaluated by the caller with the reference's
 * numpy expressions (1/m_a, 1/m_b, diagonal term per object; entry 0 = background).  The UPML frame is
 * overwritten afterwards by the caller with values computed on the O(perimeter) frame.  Synchronises.
 * ------------------------------------------------------------------------------------------ */
int fdfd_scatter_fill(const uint8_t* ids, int64_t n, int ntab, const double* table_host, void* ca, void* cb,
                      void* cd, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  matrix-free complex Helmholtz / curl-curl operator on an Nx x Ny (slab of a) Yee grid
 *     — replaces the explicit sparse products of
 *       _build_A_TE / _build_A_TM   (Scattering_Solver_2D.py:176-188) and
 *       A_tm / A_te of the band solver (Band_Diagram_Solver.py:312,319).
 *
 *   pol = FDFD_POL_TE :  y = sx*DHX diag(ca) DEX x + sy*DHY diag(cb) DEY x + cd.*x
 *   pol = FDFD_POL_TM :  y = sx*DEX diag(ca) DHX x + sy*DEY diag(cb) DHY x + cd.*x
 * with DEX.. the *unit-spacing* yeeder2d operators for (bcx, bcy, phase), i.e. sx = +-1/dx^2.
 * ca, cb, cd are device arrays of Nx*Ny_local complex numbers in `dtype` (cd may be NULL = 0).
 * phase = {phx_re, phx_im, phy_re, phy_im} = exp(-1j*k*N*d) per axis (ignored for Dirichlet).
 *
 * y-slab sharding: rank r of nranks owns rows [j0, j0+Ny_local) of the global
 * Ny_global rows.  comm = opaque communicator from fdfd_comm_create (NULL for nranks==1).
 * ------------------------------------------------------------------------------------------ */
typedef struct fdfd_helm2d fdfd_helm2d_t;
typedef struct fdfd_comm fdfd_comm_t;

int fdfd_helm2d_create(fdfd_helm2d_t** out, int dtype, int pol, int Nx, int Ny_global,
                       int j0, int Ny_local, int bcx, int bcy, const double phase[4],
                       double sx, double sy,
                       const void* ca, const void* cb, const void* cd,
                       fdfd_comm_t* comm, void* stream);
void fdfd_helm2d_destroy(fdfd_helm2d_t*);
/* y = A x  (device pointers, Nx*Ny_local complex each; halo exchange inside when sharded) */
int fdfd_helm2d_apply(fdfd_helm2d_t*, const void* x, void* y, void* stream);
/* HOST-buffer apply of the full (unsharded) operator: uploads x, applies, downloads y. */
int fdfd_helm2d_apply_host(fdfd_helm2d_t*, const void* x_host, void* y_host);
/* algorithmic bytes one apply moves (80 B/pt c128, 40 B/pt c64) */
int64_t fdfd_helm2d_bytes_per_apply(const fdfd_helm2d_t*);
/* selects the kernel variant (0 = default) — benchmarking aid */
int fdfd_helm2d_set_variant(fdfd_helm2d_t*, int variant);

/* ------------------------------------------------------------------------------------------
 * K3-K5  Krylov solve  A x = b  — replaces spla.factorized / spsolve
 *        (Scattering_Solver_2D.py:194-203, :209-217).
 * b, x: device, Nx*Ny_local complex of the operator's dtype; x holds the initial guess.
 * info (host, optional, 8 doubles): {iterations, relres (true), applies, precond_applies,
 *        seconds (device time), breakdowns, 0, 0}.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int method;       /* FDFD_KRYLOV_*                                   */
    int precond;      /* FDFD_PRECOND_*                                  */
    double rtol;      /* stop when ||b - A x|| <= rtol * ||b||           */
    int maxit;
    int restart;      /* FGMRES restart length                           */
    int mg_levels;    /* 0 = automatic                                   */
    int mg_pre, mg_post;   /* smoothing sweeps                           */
    double mg_shift_re, mg_shift_im;  /* complex shift (beta1, beta2) of the preconditioner */
    double mg_omega;  /* Jacobi damping                                  */
    int mg_cycle;     /* 1 = V, 2 = W                                    */
    int mg_coarse_its;/* smoother sweeps / Krylov its on the coarsest level */
    int verbose;      /* print the residual every `verbose` iterations (0 = silent) */
    /* tangential line relaxation inside absorbing layers (UPML): number of columns at the low/high
     * x end and rows at the low/high y end of the GLOBAL grid that get line (tridiagonal) solves
     * instead of point Jacobi; 0 = none.  The Python solver passes its add_UPML widths. */
    int line_x_lo, line_x_hi, line_y_lo, line_y_hi;
    int mg_wfrom;     /* first level (0 = finest) that is cycled as a W instead of a V; <0 = never */
    int mg_min_n;     /* stop coarsening when min(Nx, Ny) of the next level would drop below this */
    int reorth;       /* FGMRES: 1 = second Gram-Schmidt pass                                   */
    int mg_coarse_mode; /* coarsest level: 0 = mg_coarse_its smoothing sweeps, 1 = mg_coarse_its
                         * iterations of BiCGSTAB preconditioned by one smoothing sweep (a chain of ~30
                         * launches per iteration), 2 (default) = the same iterations inside ONE persistent
                         * kernel (csrc/mg_coarse.cuh) whenever the level is unsharded and its relaxation
                         * lines fit in shared memory, else as 1 */
    int mg_coarse_replicate; /* sharded grids: 1 = every rank solves the whole coarsest level, 0 = distributed,
                              * -1 = automatic (replicate while the coarse grid has <= 2^20 points)        */
    int mg_single;    /* 1 = run the multigrid cycle in complex64 inside a complex128 solve (mixed precision) */
    int mg_graph;     /* 1 = replay the multigrid cycle as a captured CUDA graph (default)            */
    double mg_kh_max; /* stop coarsening once k*h = 1/sqrt(min(sx,sy)) of the next level exceeds this */
    int mg_fuse_len;  /* levels whose relaxation lines are at most this long (and unsharded) use the fused one-CTA-per-line kernel */
    int basis_single; /* FGMRES (complex128, preconditioned): keep the Krylov basis V in complex64 (arithmetic and Z stay complex128) */
} fdfd_krylov_opts;

void fdfd_krylov_default_opts(fdfd_krylov_opts*);
/* The preconditioner built by a solve (multigrid hierarchy, line factors, captured graph) is kept in
 * the operator handle and reused by later solves with the same preconditioner options — the
 * counterpart of the reference's cached factorisation.  Call fdfd_helm2d_invalidate after changing
 * the coefficient arrays in place. */
int fdfd_helm2d_invalidate(fdfd_helm2d_t*);
int fdfd_krylov_solve(fdfd_helm2d_t* A, const fdfd_krylov_opts* opts,
                      const void* b, void* x, double* info, void* stream);
/* HOST-buffer solve (full operator): uploads b and x0, solves, downloads x. */
int fdfd_krylov_solve_host(fdfd_helm2d_t* A, const fdfd_krylov_opts* opts,
                           const void* b_host, void* x_host, double* info);

/* ------------------------------------------------------------------------------------------
 * K6  building blocks of the GPU Arnoldi / Krylov-Schur eigen-solver (fused tall-skinny kernels
 *     of K4 exposed for the host-side driver fdfd_b200/eigs.py) — replaces the ARPACK calls
 *     eigs(A, M=..., k, sigma) of Band_Diagram_Solver.py:313,320 and the Gram-Schmidt loops of
 *     refined_shift_invert_arnoldi.py:22-50.
 *   V: device, k vectors of n complex numbers at stride ldv (elements); w: device vector.
 *   fdfd_blas_multi_dot : h[i] = sum conj(V_i) * w          -> host (2k doubles), synchronises
 *   fdfd_blas_multi_axpy: w += sign * sum_i h[i] * V_i      (h from host); optional ||w||^2 -> host
 * ------------------------------------------------------------------------------------------ */
typedef struct fdfd_blas fdfd_blas_t;
int fdfd_blas_create(fdfd_blas_t** out);
void fdfd_blas_destroy(fdfd_blas_t*);
int fdfd_blas_multi_dot(fdfd_blas_t*, int dtype, int64_t n, int k, const void* V, int64_t ldv,
                        const void* w, double* h_host, void* stream);
int fdfd_blas_multi_axpy(fdfd_blas_t*, int dtype, int64_t n, int k, const void* V, int64_t ldv,
                         void* w, const double* h_host, double sign, double* nrm2_host, void* stream);

/* K7  tall-skinny QR of W = (AV - theta * BV)^T (n x m) for the refined Ritz extraction — replaces the n x m SVD
 *     of refined_shift_invert_arnoldi.py:85 (`np.linalg.svd((A - theta B) V)`): the right singular vectors of W are
 *     those of its m x m factor R.  AV, BV: device complex128, m rows of length n, leading dimension ld; BV may be
 *     NULL.  Rt_dev (device, m*m complex128) receives R column by column (Rt[c*m + i] = R[i][c]).  1 <= m <= 64.
 *     Householder panels in shared memory + a tree over the stacked triangles (csrc/tsqr.cu). */
int fdfd_tsqr_r(int m, int64_t n, const void* AV, const void* BV, int64_t ld, double theta_re, double theta_im,
                void* Rt_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6b  batched shift-invert Arnoldi for the Bloch sweep of a unit cell — replaces, for ALL path points at
 *      once, the per-point operator rebuild + `eigs(A, M, k, sigma)` (ARPACK mode 3 + SuperLU) of
 *      Band_Diagram_Solver.py:302-323.  One CTA per path point: bordered block LU of the cyclic
 *      block-tridiagonal A - sigma M in shared memory (never assembled densely), Arnoldi recurrence of
 *      (A - sigma M)^-1 M with the block substitutions and two-pass Gram-Schmidt inside one kernel per
 *      restart cycle; only the (ncv+1) x ncv projected matrices travel to the host.
 *   kind = FDFD_POL_TE / FDFD_POL_TM operator type of K2 with Bloch walls on both axes;
 *   ca, cb, m: device (Ny, Nx) complex128 (m = diagonal of M, real positive); sx, sy as in fdfd_helm2d_create;
 *   sigma: real shift LEFT of the spectrum (A - sigma M Hermitian positive definite: no pivoting);
 *   phases_host: [P][4] = (phx_re, phx_im, phy_re, phy_im) per path point, exp(-1j*k*N*d) as yeeder2d.
 * ------------------------------------------------------------------------------------------ */
typedef struct fdfd_band fdfd_band_t;
int fdfd_band_max_nx(void);      /* widest cell (Nx) whose block factorisation fits in shared memory */
int fdfd_band_create(fdfd_band_t** out, int kind, int Nx, int Ny, int P, const void* ca, const void* cb,
                     const void* m, double sx, double sy, double sigma, const double* phases_host,
                     int ncv, int kmax, void* stream);
void fdfd_band_destroy(fdfd_band_t*);
/* v0: device [P][n] unit start vectors */
int fdfd_band_set_start(fdfd_band_t*, const void* v0, void* stream);
/* path points that are still iterating (host int array) */
int fdfd_band_set_active(fdfd_band_t*, const int* idx_host, int count, void* stream);
/* Arnoldi steps j0 .. j1-1 for the active points; H_host (optional, [P][ncv+1][ncv] complex128) receives the
 * projected matrices of all points (synchronises) */
int fdfd_band_cycle(fdfd_band_t*, int j0, int j1, void* H_host, void* stream);
/* Krylov-Schur restart of the active points (host arrays in active order): V[:keep] = Q V[:ncv],
 * V[keep] = V[ncv], H <- Hnew;  Q: [Pa][keep][ncv], Hnew: [Pa][ncv+1][ncv] */
int fdfd_band_restart(fdfd_band_t*, const void* Q_host, const void* Hnew_host, int keep, void* stream);
/* Ritz vectors (unit rows) of the listed points, S_host: [count][k][ncv], and their true residuals
 * ||A x - lam M x|| / (||A x|| + |lam| ||M x||) for lam_host [count][k] -> res_host [count][k] */
int fdfd_band_ritz(fdfd_band_t*, const int* idx_host, int count, const void* S_host, const void* lam_host,
                   int k, double* res_host, void* stream);
const void* fdfd_band_vectors(fdfd_band_t*);      /* device [P][kmax][n] */
/* y = (A - sigma M) x of path point p with the entries the factorisation uses (validation hook) */
int fdfd_band_apply_shifted(fdfd_band_t*, int p, const void* x, void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2b  full-vector mode-solver operator  Omega = P Q  kept as two device CSR matrices
 *      — replaces the explicit product P_reduced @ Q_reduced and eigs' SuperLU solve
 *        (Mode_Solver_2D/Mode_Solver_2D.py:771-780).
 *   y = P (Q x) + shift * x ;  P: n x m, Q: m x n, int32 CSR arrays and complex data on the DEVICE
 *   (caller-owned, must outlive the handle).  fdfd_linop_solve solves (P Q + shift I) x = b with the
 *   K3 Krylov solvers; dinv (device, n complex, may be NULL) = 1/diag for Jacobi preconditioning.
 * ------------------------------------------------------------------------------------------ */
typedef struct fdfd_linop fdfd_linop_t;
int fdfd_csr_product_create(fdfd_linop_t** out, int dtype, int n, int m,
                            const int32_t* P_indptr, const int32_t* P_indices, const void* P_data,
                            const int32_t* Q_indptr, const int32_t* Q_indices, const void* Q_data,
                            double shift_re, double shift_im);
/* y = A x + shift * x for one square device CSR matrix (the 3-D pencil matrices A and B of
 * Periodic_Solver_3D.py:478-533 and the SpMV / SpMM of refined_shift_invert_arnoldi.py:74-75,96-97) */
int fdfd_csr_create(fdfd_linop_t** out, int dtype, int n, const int32_t* indptr, const int32_t* indices,
                    const void* data, double shift_re, double shift_im);
/* y = A x for a rectangular device CSR matrix (nrows x ncols; x has ncols, y nrows elements): the field
 * reconstruction of the mode solver, H = Q E / sqrt(lambda), Ez / Hz from the curls
 * (Mode_Solver_2D.py:792-805), runs on the device with these */
int fdfd_csr_create_rect(fdfd_linop_t** out, int dtype, int nrows, int ncols, const int32_t* indptr,
                         const int32_t* indices, const void* data);
void fdfd_linop_destroy(fdfd_linop_t*);
int fdfd_linop_apply(fdfd_linop_t*, const void* x, void* y, void* stream);
int fdfd_linop_solve(fdfd_linop_t*, const fdfd_krylov_opts* opts, const void* dinv, const void* b,
                     void* x, double* info, void* stream);

/* z = M^-1 v: one application of the preconditioner `opts->precond` selects (device pointers).
 * Exposed so the preconditioner can be validated on its own (tests/test_mg_gpu.py). */
int fdfd_precond_apply(fdfd_helm2d_t* A, const fdfd_krylov_opts* opts, const void* v, void* z,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * multi-GPU plumbing: one process per GPU; the 128-byte NCCL unique id is produced by rank 0
 * (fdfd_comm_unique_id) and distributed by the host framework (torch.distributed).
 * ------------------------------------------------------------------------------------------ */
int fdfd_comm_unique_id(void* id128);
int fdfd_comm_create(fdfd_comm_t** out, int nranks, int rank, const void* id128);
/* 1 when the data path of the communicator is NVLink peer memory (every rank mapped every other rank's
 * symmetric heap with CUDA IPC: ghost rows are pushed by the stencil kernel itself, small all-gathers and
 * inner-product reductions are flag-based kernels), 0 when it fell back to NCCL (or FDFD_P2P=0). */
int fdfd_comm_peer_mapped(const fdfd_comm_t*);
/* Destroy (or fdfd_helm2d_invalidate) every sharded operator created with the communicator first: a
 * cached multigrid preconditioner holds a CUDA graph with captured NCCL kernels, and NCCL does not
 * release a communicator while such a graph exists (ncclCommDestroy would block). */
void fdfd_comm_destroy(fdfd_comm_t*);

#ifdef __cplusplus
}
#endif
#endif /* FDFD_B200_H */
