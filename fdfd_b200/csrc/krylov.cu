// K3 — right-preconditioned BiCGSTAB and flexible GMRES.  Replaces the SuperLU direct solve of
// the reference (Scattering_Solver_2D.py:194-203 / :209-217: spla.factorized / spsolve).
//
// All vector work runs in the fused BLAS-1 kernels of blas1.cu with coefficients that stay on
// the device (scalar slots); the host only reads back one squared residual norm per iteration
// (BiCGSTAB) or one Hessenberg column (FGMRES).  Inner products are summed across y-slab ranks
// with one NCCL all-reduce per fused group.
#include "krylov.cuh"
#include "comm.cuh"

#include <cmath>
#include <complex>
#include <type_traits>
#include <vector>

namespace fdfd {

namespace {

struct DeviceBuf {
    void* p = nullptr;
    bool owned = true;
    int alloc(size_t bytes) {
        FDFD_CUDA(dev_alloc(&p, bytes));
        return FDFD_OK;
    }
    // carve from a caller-owned arena (256-byte aligned) when one is given
    int alloc(size_t bytes, KrylovArena* arena, size_t& cursor) {
        if (!arena) return alloc(bytes);
        p = (char*)arena->p + cursor;
        owned = false;
        cursor += (bytes + 255) & ~(size_t)255;
        return FDFD_OK;
    }
    ~DeviceBuf() { if (p && owned) dev_free(p); }
};

int read_slots(Scalars& sc, int s0, int count, double2* host, cudaStream_t stream) {
    FDFD_CUDA(cudaMemcpyAsync(host, sc.slot + s0, sizeof(double2) * count, cudaMemcpyDeviceToHost, stream));
    FDFD_CUDA(cudaStreamSynchronize(stream));
    return FDFD_OK;
}

struct ScalarsGuard {
    Scalars& s;
    ~ScalarsGuard() { s.release(); }
};

struct JacobiPrecond : Precond {
    int dtype = FDFD_C128;
    int64_t n = 0;
    void* dinv = nullptr;
    bool owned = true;
    ~JacobiPrecond() override { if (dinv && owned) dev_free(dinv); }
    int apply(const void* v, void* z, cudaStream_t stream) override {
        ++applies;
        if (dtype == FDFD_C128)
            return pointwise_mul<double2>(n, (double2*)z, (const double2*)v, (const double2*)dinv, stream);
        return pointwise_mul<float2>(n, (float2*)z, (const float2*)v, (const float2*)dinv, stream);
    }
};

}  // namespace

int BicgWorkspace::ensure(size_t bytes, bool precond) {
    if (!sc.slot) FDFD_TRY(sc.alloc());
    const int need = precond ? 7 : 5;
    if (bytes > cap) {
        for (int i = 0; i < 7; ++i) { dev_free(v[i]); v[i] = nullptr; }
        cap = bytes; count = 0;
    }
    for (int i = count; i < need; ++i) FDFD_CUDA(dev_alloc(&v[i], cap));
    if (need > count) count = need;
    return FDFD_OK;
}
BicgWorkspace::~BicgWorkspace() {
    for (int i = 0; i < 7; ++i) dev_free(v[i]);
    sc.release();
}

namespace {

// slot map --------------------------------------------------------------------------------------
enum { S_RHO = 0, S_RHO_OLD, S_ALPHA, S_OMEGA, S_BETA, S_NBO, S_RV, S_TS, S_TT, S_RR, S_BB, S_SS,
       S_TMP0, S_TMP1, S_TMP2, S_TMP3, S_H0 = 32 /* Hessenberg column / GMRES y (<= 200 entries) */ };

template <typename C>
int bicgstab_t(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const C* b, C* x, SolveInfo& info,
               cudaStream_t stream, BicgWorkspace* ext_ws = nullptr, bool fixed_its = false) {
    const int64_t n = A.n_local();
    const size_t bytes = (size_t)n * sizeof(C);
    BicgWorkspace own;
    BicgWorkspace& ws = ext_ws ? *ext_ws : own;
    FDFD_TRY(ws.ensure(bytes, M != nullptr));
    Scalars& sc = ws.sc;
    C *r = (C*)ws.v[0], *rh = (C*)ws.v[1], *p = (C*)ws.v[2], *v = (C*)ws.v[3], *t = (C*)ws.v[4];
    C *ph = p, *sh = r;  // unpreconditioned: p^ = p, s^ = s (= r storage)
    if (M) { ph = (C*)ws.v[5]; sh = (C*)ws.v[6]; }
    Comm* comm = A.comm;
    double2 h[4];
    int status = FDFD_OK;

    auto restart = [&]() -> int {
        // r = b - A x ; r^ = r ; p = v = 0 ; rho_old = alpha = omega = 1
        FDFD_TRY(A.residual(x, r, b, stream)); ++info.op_applies;
        FDFD_CUDA(cudaMemcpyAsync(rh, r, bytes, cudaMemcpyDeviceToDevice, stream));
        FDFD_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
        FDFD_CUDA(cudaMemsetAsync(v, 0, bytes, stream));
        FDFD_TRY(scalar_op(SOP_SET, S_RHO_OLD, 0, 0, 0, 0, 1.0, 0.0, sc, stream));
        FDFD_TRY(scalar_op(SOP_SET, S_ALPHA, 0, 0, 0, 0, 1.0, 0.0, sc, stream));
        FDFD_TRY(scalar_op(SOP_SET, S_OMEGA, 0, 0, 0, 0, 1.0, 0.0, sc, stream));
        const C* xs[2] = {r, b}; const C* ys[2] = {r, b};
        FDFD_TRY(dots<C>(n, 2, xs, ys, S_TMP0, sc, stream));     // ||r||^2, ||b||^2
        FDFD_TRY(allreduce_slots(comm, sc, S_TMP0, 2, stream));
        FDFD_TRY(scalar_op(SOP_COPY, S_RHO, S_TMP0, 0, 0, 0, 0, 0, sc, stream));  // (r^, r) = ||r||^2
        FDFD_TRY(scalar_op(SOP_COPY, S_RR, S_TMP0, 0, 0, 0, 0, 0, sc, stream));
        FDFD_TRY(scalar_op(SOP_COPY, S_BB, S_TMP1, 0, 0, 0, 0, 0, sc, stream));
        return FDFD_OK;
    };
    FDFD_TRY(restart());
    if (fixed_its) {
        // sync-free variant for inner solves: exactly o.maxit iterations, no host reads
        for (int it = 0; it < o.maxit; ++it) {
            FDFD_TRY(scalar_op(SOP_BICG_BETA, S_BETA, S_RHO, S_RHO_OLD, S_ALPHA, S_OMEGA, 0, 0, sc, stream));
            FDFD_TRY(lincomb<C>(n, p, nullptr, r, sc.slot + S_BETA, p, sc.slot + S_NBO, v, 0, nullptr, -1, -1, sc, stream));
            if (M) { FDFD_TRY(M->apply(p, ph, stream)); ++info.precond_applies; }
            FDFD_TRY(A.apply(ph, v, stream)); ++info.op_applies;
            { const C* xs[1] = {rh}; const C* ys[1] = {v}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_RV, sc, stream)); }
            FDFD_TRY(allreduce_slots(comm, sc, S_RV, 1, stream));
            FDFD_TRY(scalar_op(SOP_DIV, S_ALPHA, S_RHO, S_RV, 0, 0, 0, 0, sc, stream));
            FDFD_TRY(lincomb<C>(n, r, nullptr, r, sc.slot + S_ALPHA, v, nullptr, nullptr, 2, nullptr, -1, -1, sc, stream));
            if (M) { FDFD_TRY(M->apply(r, sh, stream)); ++info.precond_applies; }
            FDFD_TRY(A.apply(sh, t, stream)); ++info.op_applies;
            { const C* xs[2] = {t, t}; const C* ys[2] = {r, t}; FDFD_TRY(dots<C>(n, 2, xs, ys, S_TS, sc, stream)); }
            FDFD_TRY(allreduce_slots(comm, sc, S_TS, 2, stream));
            FDFD_TRY(scalar_op(SOP_DIV, S_OMEGA, S_TS, S_TT, 0, 0, 0, 0, sc, stream));
            FDFD_TRY(lincomb<C>(n, x, nullptr, x, sc.slot + S_ALPHA, ph, sc.slot + S_OMEGA, sh, 0, nullptr, -1, -1, sc, stream));
            FDFD_TRY(scalar_op(SOP_COPY, S_RHO_OLD, S_RHO, 0, 0, 0, 0, 0, sc, stream));
            FDFD_TRY(lincomb<C>(n, r, nullptr, r, sc.slot + S_OMEGA, t, nullptr, nullptr, 2, rh, S_RHO, S_RR, sc, stream));
            if (comm && comm->nranks > 1) {
                FDFD_TRY(scalar_op(SOP_COPY, S_TMP0, S_RHO, 0, 0, 0, 0, 0, sc, stream));
                FDFD_TRY(allreduce_slots(comm, sc, S_TMP0, 1, stream));
                FDFD_TRY(scalar_op(SOP_COPY, S_RHO, S_TMP0, 0, 0, 0, 0, 0, sc, stream));
            }
        }
        info.iterations = o.maxit;
        return FDFD_OK;
    }
    FDFD_TRY(read_slots(sc, S_RR, 2, h, stream));
    const double bnorm2 = h[1].x > 0 ? h[1].x : 1.0;
    double rr = h[0].x;
    const double target2 = o.rtol * o.rtol * bnorm2;
    int it = 0;
    double best = rr;
    int since_best = 0;
    double last_true = -1.0;      // true residual of the previous failed confirmation
    int floor_hits = 0;           // confirmations that failed without halving it: the attainable accuracy is reached
    while (rr > target2 && it < o.maxit) {
        ++it;
        // beta, p = r + beta (p - omega v)
        FDFD_TRY(scalar_op(SOP_BICG_BETA, S_BETA, S_RHO, S_RHO_OLD, S_ALPHA, S_OMEGA, 0, 0, sc, stream));
        FDFD_TRY(lincomb<C>(n, p, nullptr, r, sc.slot + S_BETA, p, sc.slot + S_NBO, v, 0, nullptr, -1, -1, sc, stream));
        if (M) { FDFD_TRY(M->apply(p, ph, stream)); ++info.precond_applies; }
        FDFD_TRY(A.apply(ph, v, stream)); ++info.op_applies;
        { const C* xs[1] = {rh}; const C* ys[1] = {v}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_RV, sc, stream)); }
        FDFD_TRY(allreduce_slots(comm, sc, S_RV, 1, stream));
        FDFD_TRY(scalar_op(SOP_DIV, S_ALPHA, S_RHO, S_RV, 0, 0, 0, 0, sc, stream));
        // s = r - alpha v  (in r's storage)
        FDFD_TRY(lincomb<C>(n, r, nullptr, r, sc.slot + S_ALPHA, v, nullptr, nullptr, 2, nullptr, -1, -1, sc, stream));
        if (M) { FDFD_TRY(M->apply(r, sh, stream)); ++info.precond_applies; }
        FDFD_TRY(A.apply(sh, t, stream)); ++info.op_applies;
        { const C* xs[2] = {t, t}; const C* ys[2] = {r, t}; FDFD_TRY(dots<C>(n, 2, xs, ys, S_TS, sc, stream)); }
        FDFD_TRY(allreduce_slots(comm, sc, S_TS, 2, stream));
        FDFD_TRY(scalar_op(SOP_DIV, S_OMEGA, S_TS, S_TT, 0, 0, 0, 0, sc, stream));
        // x += alpha p^ + omega s^
        FDFD_TRY(lincomb<C>(n, x, nullptr, x, sc.slot + S_ALPHA, ph, sc.slot + S_OMEGA, sh, 0, nullptr, -1, -1, sc, stream));
        // r = s - omega t ; rho_old <- rho ; rho = (r^, r) ; rr = ||r||^2
        FDFD_TRY(scalar_op(SOP_COPY, S_RHO_OLD, S_RHO, 0, 0, 0, 0, 0, sc, stream));
        FDFD_TRY(lincomb<C>(n, r, nullptr, r, sc.slot + S_OMEGA, t, nullptr, nullptr, 2, rh, S_RHO, S_RR, sc, stream));
        if (comm && comm->nranks > 1) {
            // S_RHO and S_RR are not adjacent: reduce them through the adjacent temporaries
            FDFD_TRY(scalar_op(SOP_COPY, S_TMP0, S_RHO, 0, 0, 0, 0, 0, sc, stream));
            FDFD_TRY(scalar_op(SOP_COPY, S_TMP1, S_RR, 0, 0, 0, 0, 0, sc, stream));
            FDFD_TRY(allreduce_slots(comm, sc, S_TMP0, 2, stream));
            FDFD_TRY(scalar_op(SOP_COPY, S_RHO, S_TMP0, 0, 0, 0, 0, 0, sc, stream));
            FDFD_TRY(scalar_op(SOP_COPY, S_RR, S_TMP1, 0, 0, 0, 0, 0, sc, stream));
        }
        FDFD_TRY(read_slots(sc, S_RR, 1, h, stream));
        rr = h[0].x;
        if (!std::isfinite(rr)) {
            ++info.breakdowns;
            if (info.breakdowns > 5) { status = FDFD_ERR_BREAKDOWN; set_error("BiCGSTAB breakdown (non-finite residual) at iteration %d", it); break; }
            FDFD_TRY(restart());
            FDFD_TRY(read_slots(sc, S_RR, 1, h, stream));
            rr = h[0].x;
            if (!std::isfinite(rr)) { status = FDFD_ERR_BREAKDOWN; set_error("BiCGSTAB: iterate became non-finite"); break; }
            continue;
        }
        if (o.verbose && (it % o.verbose == 0))
            fprintf(stderr, "[fdfd] bicgstab it %d relres %.3e\n", it, std::sqrt(rr / bnorm2));
        if (rr < best) { best = rr; since_best = 0; }
        else if (++since_best >= 200) {       // stagnation: rebuild the recurrences from the true residual
            ++info.breakdowns; since_best = 0;
            FDFD_TRY(restart());
            FDFD_TRY(read_slots(sc, S_RR, 1, h, stream));
            rr = h[0].x; best = rr;
        }
        if (rr <= target2) {
            // confirm with the true residual before stopping (recurrence drift)
            FDFD_TRY(A.residual(x, t, b, stream)); ++info.op_applies;
            const C* xs[1] = {t}; const C* ys[1] = {t};
            FDFD_TRY(dots<C>(n, 1, xs, ys, S_TMP0, sc, stream));
            FDFD_TRY(allreduce_slots(comm, sc, S_TMP0, 1, stream));
            FDFD_TRY(read_slots(sc, S_TMP0, 1, h, stream));
            if (h[0].x > target2 && it < o.maxit) {
                // the recurrence says converged, the true residual does not: rebuild the recurrences.  When that
                // keeps happening without progress the rounding floor of b - A x has been reached: stop instead
                // of burning maxit iterations (the caller sees the true residual and FDFD_ERR_NOCONV).
                if (last_true > 0 && h[0].x > 0.25 * last_true) ++floor_hits; else floor_hits = 0;
                last_true = h[0].x;
                if (floor_hits >= 3) { rr = h[0].x; break; }
                FDFD_TRY(restart());
                FDFD_TRY(read_slots(sc, S_RR, 1, h, stream));
                rr = h[0].x; best = rr;
            } else {
                rr = h[0].x;
            }
        }
    }
    // true residual
    FDFD_TRY(A.residual(x, t, b, stream)); ++info.op_applies;
    { const C* xs[1] = {t}; const C* ys[1] = {t}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_TMP0, sc, stream)); }
    FDFD_TRY(allreduce_slots(comm, sc, S_TMP0, 1, stream));
    FDFD_TRY(read_slots(sc, S_TMP0, 1, h, stream));
    info.iterations = it;
    info.relres = std::sqrt(h[0].x / bnorm2);
    if (status == FDFD_OK && !(info.relres <= o.rtol * 1.0000001)) {
        set_error("BiCGSTAB did not converge: relres %.3e after %d iterations", info.relres, it);
        status = FDFD_ERR_NOCONV;
    }
    return status;
}

// ------------------------------------------------------------------------------------------------
// FGMRES.  The Hessenberg matrix, the Givens rotations and the rotated right-hand side live on the
// device: one single-thread kernel per iteration folds the Gram-Schmidt coefficients into column j,
// applies / generates the rotations and leaves |g_{j+1}|^2 (the squared residual estimate) in a small
// result record, which is copied to pinned host memory asynchronously.  The host checks the record of
// iteration j-1 only after iteration j has been enqueued, so the device never waits for the host; the
// price is at most one iteration more than strictly needed.  The triangular solve for the update
// coefficients is a one-warp kernel at the end of a restart cycle.
struct GivensDev {
    double2* H = nullptr;    // column j at H + j*(m+1), entries 0..j+1
    double2* cs = nullptr;   // [m]
    double2* sn = nullptr;   // [m]
    double2* g = nullptr;    // [m+1]
    double2* rec = nullptr;  // [m] per-iteration record: (|g_{j+1}|^2, h_{j+1,j}^2)
};

__global__ void fgmres_givens_kernel(int j, int m, int add, int finalize, const double2* __restrict__ hslots,
                                     const double2* __restrict__ nrm2, const double2* __restrict__ rr, GivensDev G) {
    double2* col = G.H + (size_t)j * (m + 1);
    for (int i = 0; i <= j; ++i) {
        const double2 h = hslots[i];
        col[i] = add ? make_double2(col[i].x + h.x, col[i].y + h.y) : h;
    }
    if (!finalize) return;
    const double hn2 = nrm2->x;
    const double hn = sqrt(hn2 > 0.0 ? hn2 : 0.0);
    if (j == 0) G.g[0] = make_double2(sqrt(rr->x), 0.0);
    double2 a = col[0];
    for (int i = 0; i < j; ++i) {
        const double2 c = G.cs[i], s = G.sn[i], bb = col[i + 1];
        col[i] = cadd(cmulc(c, a), cmulc(s, bb));       // conj(c) a + conj(s) b
        a = csub(cmul(c, bb), cmul(s, a));              // -s a + c b
    }
    const double d = sqrt(a.x * a.x + a.y * a.y + hn * hn);
    double2 c = make_double2(1.0, 0.0), s = make_double2(0.0, 0.0);
    if (d > 0.0) { c = make_double2(a.x / d, a.y / d); s = make_double2(hn / d, 0.0); }
    G.cs[j] = c; G.sn[j] = s;
    col[j] = cadd(cmulc(c, a), cscale(hn, cconj(s)));
    col[j + 1] = make_double2(0.0, 0.0);
    const double2 gj = G.g[j];
    G.g[j] = cmulc(c, gj);
    const double2 gn = cneg(cmul(s, gj));
    G.g[j + 1] = gn;
    G.rec[j] = make_double2(gn.x * gn.x + gn.y * gn.y, hn2);
}

// y = R^-1 g for the leading k x k block (upper triangular after the rotations) -> slots[0..k)
__global__ void __launch_bounds__(32) fgmres_ysolve_kernel(int k, int m, GivensDev G, double2* __restrict__ yslots) {
    __shared__ double2 y[160];
    const int lane = threadIdx.x;
    for (int i = k - 1; i >= 0; --i) {
        double sx = 0.0, sy = 0.0;
        for (int q = i + 1 + lane; q < k; q += 32) {
            const double2 h = G.H[(size_t)q * (m + 1) + i], v = y[q];
            sx += h.x * v.x - h.y * v.y;
            sy += h.x * v.y + h.y * v.x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_down_sync(0xffffffffu, sx, off);
            sy += __shfl_down_sync(0xffffffffu, sy, off);
        }
        if (lane == 0) {
            const double2 gi = G.g[i], r = G.H[(size_t)i * (m + 1) + i];
            const double2 num = make_double2(gi.x - sx, gi.y - sy);
            const double dd = r.x * r.x + r.y * r.y;
            y[i] = dd > 0.0 ? make_double2((num.x * r.x + num.y * r.y) / dd, (num.y * r.x - num.x * r.y) / dd)
                            : make_double2(0.0, 0.0);
        }
        __syncwarp();
    }
    for (int i = lane; i < k; i += 32) yslots[i] = y[i];
}

struct FgmresHost {
    GivensDev G;
    double2* pinned = nullptr;           // [m] copies of the per-iteration records
    std::vector<cudaEvent_t> ev;
    int init(int m) {
        FDFD_CUDA(cudaMalloc(&G.H, sizeof(double2) * (size_t)(m + 1) * m));
        FDFD_CUDA(cudaMalloc(&G.cs, sizeof(double2) * m));
        FDFD_CUDA(cudaMalloc(&G.sn, sizeof(double2) * m));
        FDFD_CUDA(cudaMalloc(&G.g, sizeof(double2) * (m + 1)));
        FDFD_CUDA(cudaMalloc(&G.rec, sizeof(double2) * m));
        FDFD_CUDA(cudaMallocHost(&pinned, sizeof(double2) * m));
        ev.resize(m, nullptr);
        for (auto& e : ev) FDFD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return FDFD_OK;
    }
    ~FgmresHost() {
        cudaFree(G.H); cudaFree(G.cs); cudaFree(G.sn); cudaFree(G.g); cudaFree(G.rec);
        if (pinned) cudaFreeHost(pinned);
        for (auto e : ev) if (e) cudaEventDestroy(e);
    }
};

template <typename C>
int fgmres_t(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const C* b, C* x, SolveInfo& info,
             cudaStream_t stream, KrylovArena* arena) {
    const int64_t n = A.n_local();
    const size_t bytes = (size_t)n * sizeof(C);
    int m = o.restart > 0 ? o.restart : 30;
    if (m > 160) m = 160;
    Scalars sc;
    ScalarsGuard guard{sc};
    FDFD_TRY(sc.alloc());
    FgmresHost hs;
    FDFD_TRY(hs.init(m));
    // Compressed basis: V is only read by the Gram-Schmidt kernels, which dominate the traffic of a
    // restart cycle (2 (j+1) vectors per iteration).  Stored in complex64 (arithmetic stays complex128)
    // it perturbs the Arnoldi relation by 6e-8 relative per cycle, far below the ~10x a cycle gains;
    // the residual is recomputed in complex128 at every restart.  Z (the corrections) stays complex128.
    constexpr bool kCanCompress = std::is_same<C, double2>::value;
    const bool sb = kCanCompress && o.basis_single && M != nullptr;
    // ... and with a complex64 preconditioner the corrections Z are complex64 as well: FGMRES applies A to
    // the STORED z_j (K2 reads complex64 x directly), so A Z = V H holds for the rounded vectors and the
    // update x += Z y (accumulated in complex128) is exact for them — no precision conversion passes at all,
    // and the normalisation kernel writes the next basis vector straight into the preconditioner's input
    const bool zs = sb && A.has_apply_x32() && M->single_io();
    DeviceBuf bV, bZ, bw, bvc;
    const size_t vbytes = (sb ? sizeof(float2) : sizeof(C)) * (size_t)n * (m + 1);
    size_t cursor = 0;
    if (arena) {
        // basis / correction vectors of a restart cycle: tens of GB at 4096^2 and beyond; kept in the
        // operator handle between solves (cudaMalloc / cudaFree of that much memory costs 0.1-0.8 s)
        const size_t need = vbytes + bytes * ((sb && !zs ? 1 : 0) + 1) + (M ? (zs ? bytes / 2 : bytes) * m : 0) + 256 * 8;
        FDFD_TRY(arena->ensure(need));
    }
    FDFD_TRY(bV.alloc(vbytes, arena, cursor));
    if (sb && !zs) FDFD_TRY(bvc.alloc(bytes, arena, cursor));
    if (M) FDFD_TRY(bZ.alloc((zs ? bytes / 2 : bytes) * m, arena, cursor));
    FDFD_TRY(bw.alloc(bytes, arena, cursor));
    C* V = (C*)bV.p;
    float2* Vs = (float2*)bV.p;    // the same storage seen as the compressed basis
    C* vcur = (C*)bvc.p;           // complex128 copy of the newest basis vector (input of M)
    C* Z = M ? (C*)bZ.p : V;
    float2* Zs = (float2*)bZ.p;    // the corrections seen as complex64 vectors (zs)
    float2* pin = zs ? M->single_input() : nullptr;   // preconditioner input filled by the normalisation
    C* w = (C*)bw.p;
    Comm* comm = A.comm;
    double2 hbuf[2];
    int status = FDFD_OK;
    int it = 0;
    double bnorm2 = 0, relres = 1.0;
    {
        const C* xs[1] = {b}; const C* ys[1] = {b};
        FDFD_TRY(dots<C>(n, 1, xs, ys, S_BB, sc, stream));
        FDFD_TRY(allreduce_slots(comm, sc, S_BB, 1, stream));
        FDFD_TRY(read_slots(sc, S_BB, 1, hbuf, stream));
        bnorm2 = hbuf[0].x > 0 ? hbuf[0].x : 1.0;
    }
    const double bnorm = std::sqrt(bnorm2);
    const int passes = o.reorth ? 2 : 1;

    // one Arnoldi step, fully asynchronous
    auto enqueue = [&](int j) -> int {
        if constexpr (kCanCompress) {
            if (zs) {
                float2* zj = Zs + (int64_t)j * n;
                FDFD_TRY(M->apply_single(pin ? nullptr : Vs + (int64_t)j * n, zj, stream)); ++info.precond_applies;
                FDFD_TRY(A.apply_x32(zj, (double2*)w, stream)); ++info.op_applies;
            }
        }
        if (!zs) {
            C* vj = sb ? vcur : V + (int64_t)j * n;
            C* zj = Z + (int64_t)j * n;
            if (M) { FDFD_TRY(M->apply(vj, zj, stream)); ++info.precond_applies; }
            FDFD_TRY(A.apply(zj, w, stream)); ++info.op_applies;
        }
        // classical Gram-Schmidt (optionally a second pass), fused multi-vector kernels
        for (int pass = 0; pass < passes; ++pass) {
            const bool last = pass == passes - 1;
            if constexpr (kCanCompress) {
                if (sb) {
                    FDFD_TRY((multi_dot<float2, double2>(n, j + 1, Vs, n, w, S_H0, sc, stream)));
                    FDFD_TRY(allreduce_slots(comm, sc, S_H0, j + 1, stream));
                    FDFD_TRY((multi_axpy<float2, double2>(n, j + 1, Vs, n, w, S_H0, -1.0, last ? S_TMP0 : -1, sc, stream)));
                }
            }
            if (!sb) {
                FDFD_TRY(multi_dot<C>(n, j + 1, V, n, w, S_H0, sc, stream));
                FDFD_TRY(allreduce_slots(comm, sc, S_H0, j + 1, stream));
                FDFD_TRY(multi_axpy<C>(n, j + 1, V, n, w, S_H0, -1.0, last ? S_TMP0 : -1, sc, stream));
            }
            if (last) FDFD_TRY(allreduce_slots(comm, sc, S_TMP0, 1, stream));
            fgmres_givens_kernel<<<1, 1, 0, stream>>>(j, m, pass > 0, last, sc.slot + S_H0, sc.slot + S_TMP0,
                                                       sc.slot + S_RR, hs.G);
        }
        FDFD_CUDA(cudaMemcpyAsync(hs.pinned + j, hs.G.rec + j, sizeof(double2), cudaMemcpyDeviceToHost, stream));
        FDFD_CUDA(cudaEventRecord(hs.ev[j], stream));
        // next basis vector (a zero norm gives a non-finite vector: that step is discarded by the host)
        if constexpr (kCanCompress) {
            if (sb) FDFD_TRY(scale_inv_sqrt_dual(n, zs ? nullptr : vcur, Vs + (int64_t)(j + 1) * n, pin, w, S_TMP0, sc, stream));
        }
        if (!sb) FDFD_TRY(scale_inv_sqrt<C>(n, V + (int64_t)(j + 1) * n, w, S_TMP0, sc, stream));
        return FDFD_OK;
    };
    // outcome of step j once its record has arrived: 0 = continue, 1 = converged, 2 = exact breakdown
    // (h_{j+1,j} = 0: the Krylov space is invariant, the step itself is valid), -1 = non-finite
    auto outcome = [&](int j, double& rel) -> int {
        if (cudaEventSynchronize(hs.ev[j]) != cudaSuccess) return -1;
        const double g2 = hs.pinned[j].x, hn2 = hs.pinned[j].y;
        if (!std::isfinite(g2) || !std::isfinite(hn2)) return -1;
        rel = std::sqrt(g2) / bnorm;
        if (!(hn2 > 0.0)) return 2;
        return rel <= o.rtol ? 1 : 0;
    };

    bool done = false;
    while (!done) {
        // r = b - A x -> V0 (normalised)
        FDFD_TRY(A.residual(x, w, b, stream)); ++info.op_applies;
        { const C* xs[1] = {w}; const C* ys[1] = {w}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_RR, sc, stream)); }
        FDFD_TRY(allreduce_slots(comm, sc, S_RR, 1, stream));
        FDFD_TRY(read_slots(sc, S_RR, 1, hbuf, stream));
        const double beta = std::sqrt(hbuf[0].x);
        relres = beta / bnorm;
        if (!std::isfinite(beta)) { status = FDFD_ERR_BREAKDOWN; set_error("FGMRES: non-finite residual"); break; }
        if (relres <= o.rtol || it >= o.maxit) break;
        if constexpr (kCanCompress) {
            if (sb) FDFD_TRY(scale_inv_sqrt_dual(n, zs ? nullptr : vcur, Vs, pin, w, S_RR, sc, stream));
        }
        if (!sb) FDFD_TRY(scale_inv_sqrt<C>(n, V, w, S_RR, sc, stream));
        int k = 0;            // valid columns of this cycle
        int enq = 0;          // steps enqueued in this cycle
        int why = 0;          // outcome that ended the cycle
        // consume the record of step j; false = the cycle ends here
        auto examine = [&](int j) -> bool {
            double rel = relres;
            why = outcome(j, rel);
            if (why < 0) {
                if (k == 0) { status = FDFD_ERR_BREAKDOWN; set_error("FGMRES: non-finite Hessenberg entry"); }
                return false;                 // keep the k valid columns found so far
            }
            k = j + 1; relres = rel;
            if (o.verbose && ((it - (enq - k)) % o.verbose == 0))
                fprintf(stderr, "[fdfd] fgmres it %d relres %.3e\n", it - (enq - k), relres);
            return why == 0;
        };
        while (true) {
            FDFD_TRY(enqueue(enq));
            ++enq; ++it;
            // the record of the step BEFORE the one just enqueued: the device always has work queued
            if (enq >= 2 && !examine(enq - 2)) {
                if (why == 1) examine(enq - 1);   // converged: the step in flight is valid as well
                break;
            }
            if (enq >= m || it >= o.maxit) { examine(enq - 1); break; }
        }
        if (status != FDFD_OK) break;
        if (it >= o.maxit) done = true;
        // y = R^-1 g, x += Z y
        if (k > 0) {
            fgmres_ysolve_kernel<<<1, 32, 0, stream>>>(k, m, hs.G, sc.slot + S_H0);
            if constexpr (kCanCompress) {
                if (zs) FDFD_TRY((multi_axpy<float2, double2>(n, k, Zs, n, (double2*)x, S_H0, 1.0, -1, sc, stream)));
            }
            if (!zs) FDFD_TRY(multi_axpy<C>(n, k, Z, n, x, S_H0, 1.0, -1, sc, stream));
        }
        FDFD_CUDA(cudaStreamSynchronize(stream));
        if (k == 0) { status = FDFD_ERR_BREAKDOWN; set_error("FGMRES: breakdown in the first step of a cycle"); break; }
    }
    // true residual
    FDFD_TRY(A.residual(x, w, b, stream)); ++info.op_applies;
    { const C* xs[1] = {w}; const C* ys[1] = {w}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_RR, sc, stream)); }
    FDFD_TRY(allreduce_slots(comm, sc, S_RR, 1, stream));
    FDFD_TRY(read_slots(sc, S_RR, 1, hbuf, stream));
    info.iterations = it;
    info.relres = std::sqrt(hbuf[0].x) / bnorm;
    if (status == FDFD_OK && !(info.relres <= o.rtol * 1.0000001)) {
        set_error("FGMRES did not converge: relres %.3e after %d iterations", info.relres, it);
        status = FDFD_ERR_NOCONV;
    }
    return status;
}

}  // namespace

int make_jacobi_precond(Helm2D& A, Precond** out, cudaStream_t stream) {
    JacobiPrecond* P = new JacobiPrecond();
    P->dtype = A.dtype; P->n = A.n_local();
    cudaError_t e = dev_alloc(&P->dinv, (size_t)A.n_local() * A.elem());
    if (e != cudaSuccess) { delete P; return cuda_fail(e, "cudaMalloc(dinv)", __FILE__, __LINE__); }
    // halo not needed: the diagonal only depends on coefficients
    int st;
    if (A.dtype == FDFD_C128)
        st = helm2d_launch<double2>(A, APPLY_DINV, (const double2*)P->dinv, (double2*)P->dinv, nullptr,
                                    (const double2*)A.ca, (const double2*)A.cb, (const double2*)A.cd, 1.0, 0.0, 0.0, stream);
    else
        st = helm2d_launch<float2>(A, APPLY_DINV, (const float2*)P->dinv, (float2*)P->dinv, nullptr,
                                   (const float2*)A.ca, (const float2*)A.cb, (const float2*)A.cd, 1.0, 0.0, 0.0, stream);
    if (st != FDFD_OK) { delete P; return st; }
    *out = P;
    return FDFD_OK;
}

int make_diag_precond(int dtype, int64_t n, const void* d, Precond** out) {
    JacobiPrecond* P = new JacobiPrecond();
    P->dtype = dtype; P->n = n; P->dinv = const_cast<void*>(d); P->owned = false;
    *out = P;
    return FDFD_OK;
}

int bicgstab_solve(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const void* b, void* x,
                   SolveInfo& info, cudaStream_t stream, BicgWorkspace* ws, bool fixed_its) {
    if (A.dtype == FDFD_C128) return bicgstab_t<double2>(A, M, o, (const double2*)b, (double2*)x, info, stream, ws, fixed_its);
    return bicgstab_t<float2>(A, M, o, (const float2*)b, (float2*)x, info, stream, ws, fixed_its);
}
int KrylovArena::ensure(size_t bytes) {
    if (cap >= bytes) return FDFD_OK;
    release();
    FDFD_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return FDFD_OK;
}
void KrylovArena::release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
}

int fgmres_solve(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const void* b, void* x,
                 SolveInfo& info, cudaStream_t stream, KrylovArena* arena) {
    if (A.dtype == FDFD_C128) return fgmres_t<double2>(A, M, o, (const double2*)b, (double2*)x, info, stream, arena);
    return fgmres_t<float2>(A, M, o, (const float2*)b, (float2*)x, info, stream, arena);
}

}  // namespace fdfd
