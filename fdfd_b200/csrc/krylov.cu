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
        FDFD_CUDA(cudaMalloc(&p, bytes));
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
    ~DeviceBuf() { if (p && owned) cudaFree(p); }
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
    ~JacobiPrecond() override { if (dinv && owned) cudaFree(dinv); }
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
        for (int i = 0; i < 7; ++i) { cudaFree(v[i]); v[i] = nullptr; }
        cap = bytes; count = 0;
    }
    for (int i = count; i < need; ++i) FDFD_CUDA(cudaMalloc(&v[i], cap));
    if (need > count) count = need;
    return FDFD_OK;
}
BicgWorkspace::~BicgWorkspace() {
    for (int i = 0; i < 7; ++i) cudaFree(v[i]);
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
template <typename C>
int fgmres_t(LinOp& A, Precond* M, const fdfd_krylov_opts& o, const C* b, C* x, SolveInfo& info,
             cudaStream_t stream, KrylovArena* arena) {
    using cd = std::complex<double>;
    const int64_t n = A.n_local();
    const size_t bytes = (size_t)n * sizeof(C);
    int m = o.restart > 0 ? o.restart : 30;
    if (m > 160) m = 160;
    Scalars sc;
    ScalarsGuard guard{sc};
    FDFD_TRY(sc.alloc());
    // LGMRES-style augmentation: the last `kaug` slots of every cycle are filled with the
    // (normalised) corrections of the previous cycles instead of new Krylov directions; this
    // removes most of the stagnation that restarting causes on the indefinite Helmholtz operator
    int kaug = o.lgmres_k > 0 ? o.lgmres_k : 0;
    if (kaug > m / 2) kaug = m / 2;
    // Compressed basis: V is only read by the Gram-Schmidt kernels, which dominate the traffic of a
    // restart cycle (2 (j+1) vectors per iteration).  Stored in complex64 (arithmetic stays complex128)
    // it perturbs the Arnoldi relation by 6e-8 relative per cycle, far below the ~10x a cycle gains;
    // the residual is recomputed in complex128 at every restart.  Z (the corrections) stays complex128.
    constexpr bool kCanCompress = std::is_same<C, double2>::value;
    const bool sb = kCanCompress && o.basis_single && M != nullptr;
    DeviceBuf bV, bZ, bw, bE, bvc;
    const size_t vbytes = (sb ? sizeof(float2) : sizeof(C)) * (size_t)n * (m + 1);
    size_t cursor = 0;
    if (arena) {
        // basis / correction vectors of a restart cycle: tens of GB at 4096^2 and beyond; kept in the
        // operator handle between solves (cudaMalloc / cudaFree of that much memory costs 0.1-0.8 s)
        const size_t need = vbytes + bytes * ((sb ? 1 : 0) + ((M || kaug) ? m : 0) + kaug + 1) + 256 * 8;
        FDFD_TRY(arena->ensure(need));
    }
    FDFD_TRY(bV.alloc(vbytes, arena, cursor));
    if (sb) FDFD_TRY(bvc.alloc(bytes, arena, cursor));
    if (M || kaug) FDFD_TRY(bZ.alloc(bytes * m, arena, cursor));
    if (kaug) FDFD_TRY(bE.alloc(bytes * kaug, arena, cursor));
    FDFD_TRY(bw.alloc(bytes, arena, cursor));
    C* V = (C*)bV.p;
    float2* Vs = (float2*)bV.p;    // the same storage seen as the compressed basis
    C* vcur = (C*)bvc.p;           // complex128 copy of the newest basis vector (input of M)
    C* Z = (M || kaug) ? (C*)bZ.p : V;
    C* E = (C*)bE.p;
    C* w = (C*)bw.p;
    int naug = 0, aug_next = 0;
    Comm* comm = A.comm;
    std::vector<double2> hbuf(m + 8);
    std::vector<cd> H((size_t)(m + 1) * m), g(m + 1), cs(m), sn(m), y(m);
    int status = FDFD_OK;
    int it = 0;
    double bnorm2 = 0, relres = 1.0;
    {
        const C* xs[1] = {b}; const C* ys[1] = {b};
        FDFD_TRY(dots<C>(n, 1, xs, ys, S_BB, sc, stream));
        FDFD_TRY(allreduce_slots(comm, sc, S_BB, 1, stream));
        FDFD_TRY(read_slots(sc, S_BB, 1, hbuf.data(), stream));
        bnorm2 = hbuf[0].x > 0 ? hbuf[0].x : 1.0;
    }
    const double bnorm = std::sqrt(bnorm2);
    bool done = false;
    while (!done) {
        // r = b - A x -> V0 (normalised)
        FDFD_TRY(A.residual(x, w, b, stream)); ++info.op_applies;
        { const C* xs[1] = {w}; const C* ys[1] = {w}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_RR, sc, stream)); }
        FDFD_TRY(allreduce_slots(comm, sc, S_RR, 1, stream));
        FDFD_TRY(read_slots(sc, S_RR, 1, hbuf.data(), stream));
        const double beta = std::sqrt(hbuf[0].x);
        relres = beta / bnorm;
        if (!std::isfinite(beta)) { status = FDFD_ERR_BREAKDOWN; set_error("FGMRES: non-finite residual"); break; }
        if (relres <= o.rtol || it >= o.maxit) break;
        if constexpr (kCanCompress) {
            if (sb) FDFD_TRY(scale_inv_sqrt_dual(n, vcur, Vs, w, S_RR, sc, stream));
        }
        if (!sb) FDFD_TRY(scale_inv_sqrt<C>(n, V, w, S_RR, sc, stream));
        std::fill(H.begin(), H.end(), cd(0));
        std::fill(g.begin(), g.end(), cd(0));
        g[0] = beta;
        int k = 0;
        for (int j = 0; j < m; ++j) {
            C* vj = sb ? vcur : V + (int64_t)j * n;
            C* zj = Z + (int64_t)j * n;
            if (j >= m - naug) {
                // augmented direction: a previous correction (oldest first)
                const int q = (aug_next + kaug - naug + (j - (m - naug))) % kaug;
                FDFD_CUDA(cudaMemcpyAsync(zj, E + (int64_t)q * n, bytes, cudaMemcpyDeviceToDevice, stream));
            } else if (M) {
                FDFD_TRY(M->apply(vj, zj, stream)); ++info.precond_applies;
            } else if (kaug) {
                FDFD_CUDA(cudaMemcpyAsync(zj, vj, bytes, cudaMemcpyDeviceToDevice, stream));
            }
            FDFD_TRY(A.apply(zj, w, stream)); ++info.op_applies;
            // classical Gram-Schmidt with one re-orthogonalisation pass, fused multi-vector kernels
            const int passes = o.reorth ? 2 : 1;
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
                FDFD_CUDA(cudaMemcpyAsync(hbuf.data(), sc.slot + S_H0, sizeof(double2) * (j + 1), cudaMemcpyDeviceToHost, stream));
                if (last) FDFD_CUDA(cudaMemcpyAsync(hbuf.data() + j + 1, sc.slot + S_TMP0, sizeof(double2), cudaMemcpyDeviceToHost, stream));
                FDFD_CUDA(cudaStreamSynchronize(stream));
                for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] += cd(hbuf[i].x, hbuf[i].y);
            }
            const double hn = std::sqrt(hbuf[j + 1].x);
            H[(size_t)(j + 1) * m + j] = hn;
            if (hn > 0) {
                if constexpr (kCanCompress) {
                    if (sb) FDFD_TRY(scale_inv_sqrt_dual(n, vcur, Vs + (int64_t)(j + 1) * n, w, S_TMP0, sc, stream));
                }
                if (!sb) FDFD_TRY(scale_inv_sqrt<C>(n, V + (int64_t)(j + 1) * n, w, S_TMP0, sc, stream));
            }
            // Givens rotations on the new column
            for (int i = 0; i < j; ++i) {
                const cd a = H[(size_t)i * m + j], bb = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = std::conj(cs[i]) * a + std::conj(sn[i]) * bb;
                H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * bb;
            }
            {
                const cd a = H[(size_t)j * m + j], bb = H[(size_t)(j + 1) * m + j];
                const double d = std::sqrt(std::norm(a) + std::norm(bb));
                if (d == 0) { cs[j] = 1; sn[j] = 0; }
                else { cs[j] = a / d; sn[j] = bb / d; }
                H[(size_t)j * m + j] = std::conj(cs[j]) * a + std::conj(sn[j]) * bb;
                H[(size_t)(j + 1) * m + j] = 0;
                const cd gj = g[j];
                g[j] = std::conj(cs[j]) * gj;
                g[j + 1] = -sn[j] * gj;
            }
            ++it; k = j + 1;
            relres = std::abs(g[j + 1]) / bnorm;
            if (o.verbose && (it % o.verbose == 0))
                fprintf(stderr, "[fdfd] fgmres it %d relres %.3e\n", it, relres);
            if (!std::isfinite(relres)) { status = FDFD_ERR_BREAKDOWN; set_error("FGMRES: non-finite Hessenberg entry"); done = true; break; }
            if (relres <= o.rtol || it >= o.maxit || hn == 0) break;
        }
        if (status != FDFD_OK) break;
        // y = H^-1 g (upper triangular), x += Z y
        for (int i = k - 1; i >= 0; --i) {
            cd s = g[i];
            for (int q = i + 1; q < k; ++q) s -= H[(size_t)i * m + q] * y[q];
            y[i] = s / H[(size_t)i * m + i];
        }
        for (int i = 0; i < k; ++i) hbuf[i] = make_double2(y[i].real(), y[i].imag());
        FDFD_CUDA(cudaMemcpyAsync(sc.slot + S_H0, hbuf.data(), sizeof(double2) * k, cudaMemcpyHostToDevice, stream));
        if (kaug) {
            // dx = Z y kept (normalised) as an augmentation vector for the next cycles
            C* e = E + (int64_t)aug_next * n;
            FDFD_CUDA(cudaMemsetAsync(e, 0, bytes, stream));
            FDFD_TRY(multi_axpy<C>(n, k, Z, n, e, S_H0, 1.0, S_TMP1, sc, stream));
            FDFD_TRY(allreduce_slots(comm, sc, S_TMP1, 1, stream));
            FDFD_TRY(lincomb<C>(n, x, nullptr, x, nullptr, e, nullptr, nullptr, 0, nullptr, -1, -1, sc, stream));
            FDFD_TRY(scale_inv_sqrt<C>(n, e, e, S_TMP1, sc, stream));
            aug_next = (aug_next + 1) % kaug;
            if (naug < kaug) ++naug;
        } else {
            FDFD_TRY(multi_axpy<C>(n, k, Z, n, x, S_H0, 1.0, -1, sc, stream));
        }
        FDFD_CUDA(cudaStreamSynchronize(stream));
    }
    // true residual
    FDFD_TRY(A.residual(x, w, b, stream)); ++info.op_applies;
    { const C* xs[1] = {w}; const C* ys[1] = {w}; FDFD_TRY(dots<C>(n, 1, xs, ys, S_RR, sc, stream)); }
    FDFD_TRY(allreduce_slots(comm, sc, S_RR, 1, stream));
    FDFD_TRY(read_slots(sc, S_RR, 1, hbuf.data(), stream));
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
    cudaError_t e = cudaMalloc(&P->dinv, (size_t)A.n_local() * A.elem());
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
