// extern "C" surface of libfdfd_b200.so (see include/fdfd_b200.h).
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <vector>

#include "comm.cuh"
#include "helm2d.cuh"
#include "krylov.cuh"

namespace fdfd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    cudaGetLastError();  // clear the sticky flag for non-fatal errors
    return FDFD_ERR_CUDA;
}

}  // namespace fdfd

using namespace fdfd;

namespace { struct HostPipe; }
struct fdfd_helm2d {
    Helm2D op;
    HostPipe* pipe = nullptr;
    // preconditioner kept between solves with the same operator (the analogue of the reference's
    // cached factorisation `_LU_TE/_LU_TM`, Scattering_Solver_2D.py:192-194): hierarchy, line-solver
    // factors and the captured graph are built once
    Precond* cached_M = nullptr;
    fdfd_krylov_opts cached_opts;
    KrylovArena arena;      // Krylov basis memory, kept with the preconditioner
};
struct fdfd_comm { Comm* c; };

extern "C" const char* fdfd_last_error(void) { return g_err; }
extern "C" int fdfd_version(void) { return 200; }   // 200: round-2 ABI (options struct without lgmres_k, band / stagger / tsqr / scatter_fill entry points)
extern "C" int fdfd_krylov_opts_size(void) { return (int)sizeof(fdfd_krylov_opts); }
extern "C" int fdfd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int fdfd_comm_unique_id(void* id128) { return comm_unique_id(id128); }
extern "C" int fdfd_comm_create(fdfd_comm_t** out, int nranks, int rank, const void* id128) {
    FDFD_CHECK_ARG(out, "null out");
    Comm* c = nullptr;
    FDFD_TRY(comm_create(&c, nranks, rank, id128));
    *out = new fdfd_comm{c};
    return FDFD_OK;
}
extern "C" int fdfd_comm_peer_mapped(const fdfd_comm_t* c) { return c && c->c && p2p_enabled(c->c) ? 1 : 0; }
extern "C" void fdfd_comm_destroy(fdfd_comm_t* c) {
    if (!c) return;
    comm_destroy(c->c);
    delete c;
}

extern "C" int fdfd_helm2d_create(fdfd_helm2d_t** out, int dtype, int pol, int Nx, int Ny_global,
                                  int j0, int Ny_local, int bcx, int bcy, const double phase[4],
                                  double sx, double sy, const void* ca, const void* cb,
                                  const void* cd, fdfd_comm_t* comm, void* stream) {
    FDFD_CHECK_ARG(out, "null out");
    FDFD_CHECK_ARG(dtype == FDFD_C128 || dtype == FDFD_C64, "dtype must be FDFD_C128 or FDFD_C64");
    FDFD_CHECK_ARG(pol == FDFD_POL_TE || pol == FDFD_POL_TM, "pol must be TE or TM");
    FDFD_CHECK_ARG(Nx >= 2 && Ny_global >= 2, "matrix-free operator needs Nx, Ny >= 2");
    FDFD_CHECK_ARG(j0 >= 0 && Ny_local >= 1 && j0 + Ny_local <= Ny_global, "bad slab extent");
    FDFD_CHECK_ARG((bcx == 0 || bcx == 1) && (bcy == 0 || bcy == 1), "BC must be 0 or 1");
    FDFD_CHECK_ARG(ca && cb, "null coefficient array");
    FDFD_CHECK_ARG((int64_t)Nx * Ny_local < ((int64_t)1 << 40), "slab too large");
    const bool sharded = Ny_local != Ny_global;
    FDFD_CHECK_ARG(!sharded || (comm && comm->c && comm->c->nranks > 1), "sharded operator needs a communicator");
    if (fdfd_device_count() <= 0) { set_error("no CUDA device visible: fdfd_b200 has no CPU path"); return FDFD_ERR_CUDA; }
    fdfd_helm2d* h = new fdfd_helm2d();
    Helm2D& op = h->op;
    memset(&op, 0, sizeof(op));
    op.dtype = dtype; op.pol = pol; op.Nx = Nx; op.Ny_global = Ny_global; op.j0 = j0; op.Ny = Ny_local;
    op.bcx = bcx; op.bcy = bcy;
    if (phase) memcpy(op.phase, phase, sizeof(op.phase));
    else { op.phase[0] = 1; op.phase[1] = 0; op.phase[2] = 1; op.phase[3] = 0; }
    op.sx = sx; op.sy = sy; op.ca = ca; op.cb = cb; op.cd = cd;
    op.shift_re = 1.0; op.shift_im = 0.0;
    op.comm = comm ? comm->c : nullptr;
    op.variant = 0;
    cudaStream_t st = (cudaStream_t)stream;
    {
        int st_ = helm2d_setup_sharding(op, st);
        if (st_ != FDFD_OK) { fdfd_helm2d_destroy(h); return st_; }
    }
    *out = h;
    return FDFD_OK;
}

static void destroy_pipe(HostPipe* p);
extern "C" void fdfd_helm2d_destroy(fdfd_helm2d_t* h) {
    if (!h) return;
    delete h->cached_M;          // first: a captured cycle may still reference the ghost buffers
    helm2d_release_sharding(h->op);
    cudaFree(h->op.h_x); cudaFree(h->op.h_y);
    destroy_pipe(h->pipe);
    delete h;
}

extern "C" int fdfd_helm2d_apply(fdfd_helm2d_t* h, const void* x, void* y, void* stream) {
    FDFD_CHECK_ARG(h && x && y, "null pointer");
    FDFD_CHECK_ARG(x != y, "apply cannot run in place");
    return helm2d_apply(h->op, APPLY_AX, x, y, nullptr, 0.0, (cudaStream_t)stream);
}

static int ensure_host_scratch(Helm2D& op) {
    const size_t bytes = (size_t)op.n_local() * op.elem();
    if (!op.h_x) FDFD_CUDA(cudaMalloc(&op.h_x, bytes));
    if (!op.h_y) FDFD_CUDA(cudaMalloc(&op.h_y, bytes));
    return FDFD_OK;
}

// Host-buffer apply, pipelined: the grid is cut into row chunks; chunk k+1 is uploaded while chunk k
// is computed and chunk k-1 is downloaded (full-duplex PCIe), each on its own non-blocking stream.
// The stencil of chunk k reads one row of chunk k+1, so its launch waits for that upload.
namespace {
struct HostPipe {
    cudaStream_t up = nullptr, cmp = nullptr, down = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_cmp;
    int ensure(size_t nchunks) {
        if (!up) {
            FDFD_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
            FDFD_CUDA(cudaStreamCreateWithFlags(&cmp, cudaStreamNonBlocking));
            FDFD_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
        }
        while (ev_up.size() < nchunks) {
            cudaEvent_t a, b;
            FDFD_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            FDFD_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            ev_up.push_back(a); ev_cmp.push_back(b);
        }
        return FDFD_OK;
    }
    ~HostPipe() {
        for (auto e : ev_up) cudaEventDestroy(e);
        for (auto e : ev_cmp) cudaEventDestroy(e);
        if (up) cudaStreamDestroy(up);
        if (cmp) cudaStreamDestroy(cmp);
        if (down) cudaStreamDestroy(down);
    }
};

template <typename C>
int apply_host_pipelined(Helm2D& op, HostPipe& pipe, const C* xh, C* yh) {
    const int Nx = op.Nx, Ny = op.Ny;
    // chunk size: the pipeline's fill / drain costs two chunk transfers on top of the streaming time, so
    // chunks are kept small (FDFD_HOST_CHUNK_MB overrides; measured in profiles/r02_host_pipeline_chunks.txt)
    static const int chunk_mb = [] { const char* e = getenv("FDFD_HOST_CHUNK_MB"); int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
    int rows = (int)(((size_t)chunk_mb << 20) / ((size_t)Nx * sizeof(C)));
    rows = rows < 8 ? 8 : (rows / 8) * 8;                                 // multiple of the tile height
    const int nch = (Ny + rows - 1) / rows;
    FDFD_TRY(pipe.ensure((size_t)nch));
    C* xd = (C*)op.h_x;
    C* yd = (C*)op.h_y;
    for (int k = 0; k < nch; ++k) {
        const int r0 = k * rows, r1 = std::min(Ny, r0 + rows);
        FDFD_CUDA(cudaMemcpyAsync(xd + (size_t)r0 * Nx, xh + (size_t)r0 * Nx, (size_t)(r1 - r0) * Nx * sizeof(C),
                                  cudaMemcpyHostToDevice, pipe.up));
        FDFD_CUDA(cudaEventRecord(pipe.ev_up[k], pipe.up));
    }
    HelmArgs<C> a = helm2d_args<C>(op, xd, yd, nullptr, (const C*)op.ca, (const C*)op.cb, (const C*)op.cd,
                                   op.shift_re, op.shift_im, 0.0);
    for (int k = 0; k < nch; ++k) {
        const int r0 = k * rows, r1 = std::min(Ny, r0 + rows);
        // rows of this chunk read x rows r0-1 .. r1: need uploads up to chunk k+1 (and, for a Bloch
        // wrap in y, the first and last chunk)
        FDFD_CUDA(cudaStreamWaitEvent(pipe.cmp, pipe.ev_up[std::min(k + 1, nch - 1)], 0));
        if (op.bcy == FDFD_BC_BLOCH && (k == 0 || k == nch - 1))
            FDFD_CUDA(cudaStreamWaitEvent(pipe.cmp, pipe.ev_up[nch - 1], 0));
        a.row_begin = r0; a.row_end = r1;
        FDFD_TRY(helm2d_launch_args<C>(op, APPLY_AX, a, pipe.cmp));
        FDFD_CUDA(cudaEventRecord(pipe.ev_cmp[k], pipe.cmp));
        FDFD_CUDA(cudaStreamWaitEvent(pipe.down, pipe.ev_cmp[k], 0));
        FDFD_CUDA(cudaMemcpyAsync(yh + (size_t)r0 * Nx, yd + (size_t)r0 * Nx, (size_t)(r1 - r0) * Nx * sizeof(C),
                                  cudaMemcpyDeviceToHost, pipe.down));
    }
    FDFD_CUDA(cudaStreamSynchronize(pipe.down));
    return FDFD_OK;
}
}  // namespace

extern "C" int fdfd_helm2d_apply_host(fdfd_helm2d_t* h, const void* x_host, void* y_host) {
    FDFD_CHECK_ARG(h && x_host && y_host, "null pointer");
    Helm2D& op = h->op;
    FDFD_TRY(ensure_host_scratch(op));
    const size_t bytes = (size_t)op.n_local() * op.elem();
    if (op.Ny == op.Ny_global && op.Ny >= 64) {
        if (!h->pipe) h->pipe = new HostPipe();
        if (op.dtype == FDFD_C128) return apply_host_pipelined<double2>(op, *h->pipe, (const double2*)x_host, (double2*)y_host);
        return apply_host_pipelined<float2>(op, *h->pipe, (const float2*)x_host, (float2*)y_host);
    }
    // sharded slab (ghost rows come from the neighbours) or tiny grid: plain upload / apply / download
    FDFD_CUDA(cudaMemcpyAsync(op.h_x, x_host, bytes, cudaMemcpyHostToDevice, 0));
    FDFD_TRY(helm2d_apply(op, APPLY_AX, op.h_x, op.h_y, nullptr, 0.0, 0));
    FDFD_CUDA(cudaMemcpyAsync(y_host, op.h_y, bytes, cudaMemcpyDeviceToHost, 0));
    FDFD_CUDA(cudaStreamSynchronize(0));
    return FDFD_OK;
}

static void destroy_pipe(HostPipe* p) { delete p; }

extern "C" int64_t fdfd_helm2d_bytes_per_apply(const fdfd_helm2d_t* h) {
    if (!h) return 0;
    const Helm2D& op = h->op;
    const int arrays = 4 + (op.cd ? 1 : 0);   // x, y, ca, cb (+ cd)
    return (int64_t)arrays * op.n_local() * (int64_t)op.elem();
}

extern "C" int fdfd_helm2d_set_variant(fdfd_helm2d_t* h, int variant) {
    FDFD_CHECK_ARG(h, "null operator");
    h->op.variant = variant;
    return FDFD_OK;
}

extern "C" void fdfd_krylov_default_opts(fdfd_krylov_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->method = FDFD_KRYLOV_BICGSTAB;
    o->precond = FDFD_PRECOND_JACOBI;
    o->rtol = 1e-8;
    o->maxit = 100000;
    o->restart = 20;
    o->mg_levels = 0;
    o->mg_pre = 1; o->mg_post = 1;
    o->mg_shift_re = 1.0; o->mg_shift_im = 0.4;
    o->mg_omega = 0.7;
    o->mg_cycle = 1;
    o->mg_coarse_its = 0;          /* 0 = automatic from k*h of the coarsest level */
    o->mg_coarse_mode = 2;         /* persistent kernel where the level fits it, launch chain otherwise */
    o->mg_graph = 1;
    o->mg_coarse_replicate = -1;
    o->mg_single = 0;
    o->restart = 30;
    o->mg_kh_max = 1.3;
    o->mg_fuse_len = 1024;
    o->basis_single = 0;
    o->verbose = 0;
    o->line_x_lo = o->line_x_hi = o->line_y_lo = o->line_y_hi = 0;
    o->mg_wfrom = 3;
    o->mg_min_n = 32;
    o->reorth = 0;
}

static bool same_precond(const fdfd_krylov_opts& a, const fdfd_krylov_opts& b) {
    return a.precond == b.precond && a.mg_levels == b.mg_levels && a.mg_pre == b.mg_pre && a.mg_post == b.mg_post &&
           a.mg_shift_re == b.mg_shift_re && a.mg_shift_im == b.mg_shift_im && a.mg_omega == b.mg_omega &&
           a.mg_cycle == b.mg_cycle && a.mg_coarse_its == b.mg_coarse_its && a.line_x_lo == b.line_x_lo &&
           a.line_x_hi == b.line_x_hi && a.line_y_lo == b.line_y_lo && a.line_y_hi == b.line_y_hi &&
           a.mg_wfrom == b.mg_wfrom && a.mg_min_n == b.mg_min_n && a.mg_coarse_mode == b.mg_coarse_mode &&
           a.mg_coarse_replicate == b.mg_coarse_replicate && a.mg_single == b.mg_single && a.mg_graph == b.mg_graph && a.mg_kh_max == b.mg_kh_max && a.mg_fuse_len == b.mg_fuse_len;
}

extern "C" int fdfd_helm2d_invalidate(fdfd_helm2d_t* h) {
    FDFD_CHECK_ARG(h, "null operator");
    delete h->cached_M;
    h->cached_M = nullptr;
    h->arena.release();
    return FDFD_OK;
}

extern "C" int fdfd_krylov_solve(fdfd_helm2d_t* h, const fdfd_krylov_opts* opts, const void* b,
                                 void* x, double* info_out, void* stream) {
    FDFD_CHECK_ARG(h && opts && b && x, "null pointer");
    FDFD_CHECK_ARG(opts->rtol > 0 && opts->maxit >= 0, "bad tolerance / maxit");
    Helm2D& A = h->op;
    cudaStream_t st = (cudaStream_t)stream;
    Precond* M = nullptr;
    int s = FDFD_OK;
    if (opts->precond == FDFD_PRECOND_NONE) {
        M = nullptr;
    } else if (h->cached_M && same_precond(h->cached_opts, *opts)) {
        M = h->cached_M;
    } else {
        delete h->cached_M;
        h->cached_M = nullptr;
        if (opts->precond == FDFD_PRECOND_JACOBI) s = make_jacobi_precond(A, &M, st);
        else if (opts->precond == FDFD_PRECOND_MG) s = make_mg_precond(A, *opts, &M, st);
        else { set_error("unknown preconditioner %d", opts->precond); s = FDFD_ERR_ARG; }
        if (s != FDFD_OK) return s;
        h->cached_M = M;
        h->cached_opts = *opts;
    }
    SolveInfo info;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    HelmOp op(A);
    if (opts->method == FDFD_KRYLOV_BICGSTAB) s = bicgstab_solve(op, M, *opts, b, x, info, st);
    else if (opts->method == FDFD_KRYLOV_FGMRES) s = fgmres_solve(op, M, *opts, b, x, info, st, &h->arena);
    else { set_error("unknown Krylov method %d", opts->method); s = FDFD_ERR_ARG; }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    info.seconds = ms * 1e-3;
    if (info_out) {
        info_out[0] = info.iterations; info_out[1] = info.relres; info_out[2] = (double)info.op_applies;
        info_out[3] = (double)info.precond_applies; info_out[4] = info.seconds; info_out[5] = info.breakdowns;
        info_out[6] = 0; info_out[7] = 0;
    }
    return s;
}

extern "C" int fdfd_precond_apply(fdfd_helm2d_t* h, const fdfd_krylov_opts* opts, const void* v,
                                  void* z, void* stream) {
    FDFD_CHECK_ARG(h && opts && v && z, "null pointer");
    Helm2D& A = h->op;
    cudaStream_t st = (cudaStream_t)stream;
    Precond* M = nullptr;
    int s = FDFD_OK;
    if (opts->precond == FDFD_PRECOND_JACOBI) s = make_jacobi_precond(A, &M, st);
    else if (opts->precond == FDFD_PRECOND_MG) s = make_mg_precond(A, *opts, &M, st);
    else { set_error("fdfd_precond_apply: no preconditioner selected"); s = FDFD_ERR_ARG; }
    if (s != FDFD_OK) return s;
    s = M->apply(v, z, st);
    cudaStreamSynchronize(st);
    delete M;
    return s;
}

extern "C" int fdfd_krylov_solve_host(fdfd_helm2d_t* h, const fdfd_krylov_opts* opts,
                                      const void* b_host, void* x_host, double* info) {
    FDFD_CHECK_ARG(h && opts && b_host && x_host, "null pointer");
    Helm2D& op = h->op;
    FDFD_CHECK_ARG(op.Ny == op.Ny_global, "host-buffer solve is for the unsharded operator");
    FDFD_TRY(ensure_host_scratch(op));
    const size_t bytes = (size_t)op.n_local() * op.elem();
    FDFD_CUDA(cudaMemcpyAsync(op.h_x, b_host, bytes, cudaMemcpyHostToDevice, 0));
    FDFD_CUDA(cudaMemcpyAsync(op.h_y, x_host, bytes, cudaMemcpyHostToDevice, 0));
    int s = fdfd_krylov_solve(h, opts, op.h_x, op.h_y, info, nullptr);
    if (s != FDFD_OK && s != FDFD_ERR_NOCONV) return s;
    FDFD_CUDA(cudaMemcpyAsync(x_host, op.h_y, bytes, cudaMemcpyDeviceToHost, 0));
    FDFD_CUDA(cudaStreamSynchronize(0));
    return s;
}

// ---- K6 building blocks ------------------------------------------------------------------------
struct fdfd_blas { Scalars sc; };

extern "C" int fdfd_blas_create(fdfd_blas_t** out) {
    FDFD_CHECK_ARG(out, "null out");
    if (fdfd_device_count() <= 0) { set_error("no CUDA device visible: fdfd_b200 has no CPU path"); return FDFD_ERR_CUDA; }
    fdfd_blas* b = new fdfd_blas();
    int s = b->sc.alloc();
    if (s != FDFD_OK) { delete b; return s; }
    *out = b;
    return FDFD_OK;
}
extern "C" void fdfd_blas_destroy(fdfd_blas_t* b) {
    if (!b) return;
    b->sc.release();
    delete b;
}

static const int kBlasSlot0 = 16;   // slots [16, 16+k) hold h, slot 8 the squared norm

extern "C" int fdfd_blas_multi_dot(fdfd_blas_t* b, int dtype, int64_t n, int k, const void* V, int64_t ldv,
                                   const void* w, double* h_host, void* stream) {
    FDFD_CHECK_ARG(b && V && w && h_host, "null pointer");
    FDFD_CHECK_ARG(k >= 1 && k <= Scalars::kSlots - kBlasSlot0 - 8, "k out of range");
    FDFD_CHECK_ARG(n >= 1 && ldv >= n, "bad vector shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FDFD_C128) FDFD_TRY(multi_dot<double2>(n, k, (const double2*)V, ldv, (const double2*)w, kBlasSlot0, b->sc, st));
    else if (dtype == FDFD_C64) FDFD_TRY(multi_dot<float2>(n, k, (const float2*)V, ldv, (const float2*)w, kBlasSlot0, b->sc, st));
    else { set_error("bad dtype"); return FDFD_ERR_ARG; }
    FDFD_CUDA(cudaMemcpyAsync(h_host, b->sc.slot + kBlasSlot0, sizeof(double2) * k, cudaMemcpyDeviceToHost, st));
    FDFD_CUDA(cudaStreamSynchronize(st));
    return FDFD_OK;
}

extern "C" int fdfd_blas_multi_axpy(fdfd_blas_t* b, int dtype, int64_t n, int k, const void* V, int64_t ldv,
                                    void* w, const double* h_host, double sign, double* nrm2_host, void* stream) {
    FDFD_CHECK_ARG(b && V && w && h_host, "null pointer");
    FDFD_CHECK_ARG(k >= 1 && k <= Scalars::kSlots - kBlasSlot0 - 8, "k out of range");
    FDFD_CHECK_ARG(n >= 1 && ldv >= n, "bad vector shape");
    cudaStream_t st = (cudaStream_t)stream;
    FDFD_CUDA(cudaMemcpyAsync(b->sc.slot + kBlasSlot0, h_host, sizeof(double2) * k, cudaMemcpyHostToDevice, st));
    const int nslot = nrm2_host ? 8 : -1;
    if (dtype == FDFD_C128) FDFD_TRY(multi_axpy<double2>(n, k, (const double2*)V, ldv, (double2*)w, kBlasSlot0, sign, nslot, b->sc, st));
    else if (dtype == FDFD_C64) FDFD_TRY(multi_axpy<float2>(n, k, (const float2*)V, ldv, (float2*)w, kBlasSlot0, sign, nslot, b->sc, st));
    else { set_error("bad dtype"); return FDFD_ERR_ARG; }
    if (nrm2_host) {
        double2 t;
        FDFD_CUDA(cudaMemcpyAsync(&t, b->sc.slot + 8, sizeof(double2), cudaMemcpyDeviceToHost, st));
        FDFD_CUDA(cudaStreamSynchronize(st));
        *nrm2_host = t.x;
    }
    return FDFD_OK;
}
