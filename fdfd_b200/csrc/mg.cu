// K5 — complex shifted-Laplacian multigrid preconditioner for the scattering operator.
//
// M = A with its diagonal term multiplied by (beta1 + j beta2)  (shift folded into the K2
// kernel argument `shift`), approximated by one multigrid cycle:
//   * cell-centred 2x coarsening, re-discretised coarse operators of the same TE/TM stencil
//     type: face coefficients are the arithmetic mean of the two fine faces on the coarse
//     face, the diagonal term the mean of the four fine cells, sx, sy divided by 4;
//   * the Dirichlet wall (TE: far end, TM: near end) keeps its physical position: the coarse
//     wall-face coefficient is scaled by delta_l / delta_{l+1}, delta = distance of the wall
//     cell centre to the wall in units of the level spacing (1 on the finest level,
//     delta_{l+1} = (delta_l + 1/2) / 2);
//   * restriction = mean of the four fine residuals, prolongation = cell-centred bilinear
//     (clamped at the flux-free side, interpolating to 0 at the wall);
//   * smoother = damped point Jacobi, except inside the absorbing layers where the UPML makes
//     the operator strongly anisotropic with complex coefficients of opposite phase (point
//     relaxation diverges there): columns of the x-layers are relaxed by y-line (tridiagonal)
//     solves, then rows of the y-layers by x-line solves (alternating-direction line Jacobi);
//   * the tridiagonal systems are static, so they are factorised once per hierarchy with a
//     partitioned (Wang / SPIKE) scheme: P chunks per line, local LU + spikes, and a dense
//     inverse of the 2P x 2P interface system.  A solve is three fully parallel kernels
//     (chunk-local forward/backward substitution, interface mat-vec, spike correction).
// The algorithm is restated in numpy in oracle/mg_oracle.py, which the GPU tests compare with.
#include "krylov.cuh"
#include "comm.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

namespace fdfd {

namespace {

constexpr int kMaxChunks = 32;

// ---------------------------------------------------------------------------------------------
// matrix entries of the operator at one grid point (Dirichlet / flux-free walls only)
template <typename C>
struct PointCoefs { C kw, ke, ks, kn, dg; };

// y-slab sharding: `open_s` / `open_n` say that local row 0 / Ny-1 has a neighbour row on another
// rank; the face coefficient across that boundary is `cb_ghost` (TE: south rank's last cb row,
// TM: north rank's first cb row — the same buffer K2 uses).
template <typename C>
struct SlabEdge { bool open_s, open_n; const C* cb_ghost; };

template <typename C, bool TM>
__device__ __forceinline__ PointCoefs<C> point_coefs(const C* ca, const C* cb, const C* cd, int Nx, int Ny,
        typename real_of<C>::type sx, typename real_of<C>::type sy, C shift, int i, int j,
        SlabEdge<C> edge = SlabEdge<C>{false, false, nullptr}) {
    PointCoefs<C> p;
    const int64_t n = (int64_t)j * Nx + i;
    C fw, fe, fs, fn;          // face coefficients incl. wall faces
    bool has_w, has_e, has_s, has_n;   // neighbour unknown exists
    if (TM) {
        fw = ca[n]; has_w = i > 0;
        has_e = i < Nx - 1; fe = has_e ? ca[n + 1] : czero<C>();
        fs = cb[n]; has_s = j > 0 || edge.open_s;
        has_n = j < Ny - 1 || edge.open_n;
        fn = j < Ny - 1 ? cb[n + Nx] : (edge.open_n ? edge.cb_ghost[i] : czero<C>());
    } else {
        has_w = i > 0; fw = has_w ? ca[n - 1] : czero<C>();
        fe = ca[n]; has_e = i < Nx - 1;
        has_s = j > 0 || edge.open_s;
        fs = j > 0 ? cb[n - Nx] : (edge.open_s ? edge.cb_ghost[i] : czero<C>());
        fn = cb[n]; has_n = j < Ny - 1 || edge.open_n;
    }
    p.kw = has_w ? cscale(sx, fw) : czero<C>();
    p.ke = has_e ? cscale(sx, fe) : czero<C>();
    p.ks = has_s ? cscale(sy, fs) : czero<C>();
    p.kn = has_n ? cscale(sy, fn) : czero<C>();
    p.dg = cadd(cscale(-sx, cadd(fw, fe)), cscale(-sy, cadd(fs, fn)));
    if (cd) p.dg = cfma(shift, cd[n], p.dg);
    return p;
}

// ---------------------------------------------------------------------------------------------
// line sets
template <typename C>
struct LineSet {
    int dir = 0;          // 0: lines along y inside x-strips ; 1: lines along x inside y-strips
    int lo = 0, hi = 0;   // strip widths (number of lines at the low / high end)
    int nl = 0;           // lo + hi
    int len = 0;          // unknowns per line
    int m = 0, P = 0;     // chunk length, chunks per line ON THIS RANK
    int nranks = 1;       // ranks a line is spread over (y-lines of a sharded grid), else 1
    int Ptot = 0, k0 = 0; // chunks per line over all ranks, global index of this rank's first chunk
    C *fl = nullptr, *fc = nullptr, *fip = nullptr, *sv = nullptr, *sw = nullptr;  // [len][nl]
    C *rinv = nullptr;    // [nl][2Ptot*2Ptot] row-major
    C *g = nullptr;       // [len][nl] right-hand side / solution workspace
    C *ifg = nullptr;     // [nl][2P]   interface values of the local chunks
    C *ifg_all = nullptr; // [nranks][nl][2P] gathered
    C *ifu = nullptr;     // [nl][2Ptot] interface solution (replicated)
    C *tips = nullptr, *tips_all = nullptr;  // spike tips [nl][P][4] / [nranks][nl][P][4]
    GatherPlan* plan = nullptr;   // peer-memory all-gather of the interface values (sharded y-lines)
    Comm* plan_comm = nullptr;
    void release() {
        if (plan) { p2p_gather_destroy(plan_comm, plan); plan = nullptr; }
        cudaFree(fl); cudaFree(fc); cudaFree(fip); cudaFree(sv); cudaFree(sw); cudaFree(rinv);
        cudaFree(g); cudaFree(ifg); cudaFree(ifu); cudaFree(ifg_all); cudaFree(tips); cudaFree(tips_all);
        fl = fc = fip = sv = sw = rinv = g = ifg = ifu = ifg_all = tips = tips_all = nullptr;
    }
};

template <typename C>
struct LineView {
    int dir, lo, hi, nl, len, m, P, Nx, Ny, nranks, Ptot, k0;
    C *fl, *fc, *fip, *sv, *sw, *rinv, *g, *ifg, *ifg_all, *ifu, *tips, *tips_all;
};

template <typename C>
LineView<C> view(const LineSet<C>& s, int Nx, int Ny) {
    LineView<C> v;
    v.dir = s.dir; v.lo = s.lo; v.hi = s.hi; v.nl = s.nl; v.len = s.len; v.m = s.m; v.P = s.P;
    v.Nx = Nx; v.Ny = Ny; v.nranks = s.nranks; v.Ptot = s.Ptot; v.k0 = s.k0;
    v.fl = s.fl; v.fc = s.fc; v.fip = s.fip; v.sv = s.sv; v.sw = s.sw; v.rinv = s.rinv;
    v.g = s.g; v.ifg = s.ifg; v.ifg_all = s.ifg_all; v.ifu = s.ifu; v.tips = s.tips; v.tips_all = s.tips_all;
    return v;
}

// grid coordinate of (line l, position t)
template <typename C>
__device__ __forceinline__ void line_point(const LineView<C>& v, int l, int t, int& i, int& j) {
    if (v.dir == 0) { i = l < v.lo ? l : v.Nx - v.hi + (l - v.lo); j = t; }
    else { j = l < v.lo ? l : v.Ny - v.hi + (l - v.lo); i = t; }
}

// tridiagonal entries of every line: fl <- sub, fip <- diag, fc <- super
template <typename C, bool TM>
__global__ void line_extract_kernel(LineView<C> v, const C* ca, const C* cb, const C* cd,
        typename real_of<C>::type sx, typename real_of<C>::type sy, C shift, SlabEdge<C> edge) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)v.nl * v.len) return;
    const int l = (int)(idx % v.nl), t = (int)(idx / v.nl);
    int i, j;
    line_point(v, l, t, i, j);
    PointCoefs<C> p = point_coefs<C, TM>(ca, cb, cd, v.Nx, v.Ny, sx, sy, shift, i, j, edge);
    v.fl[idx] = v.dir == 0 ? p.ks : p.kw;
    v.fc[idx] = v.dir == 0 ? p.kn : p.ke;
    v.fip[idx] = p.dg;
}

// chunk-local LU (Thomas) + spikes; one thread per (line, chunk)
template <typename C>
__global__ void line_factor_kernel(LineView<C> v) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= v.nl * v.P) return;
    const int l = idx % v.nl, k = idx / v.nl;
    const int t0 = k * v.m, t1 = min(v.len, t0 + v.m);
    if (t0 >= t1) return;
    const int64_t nl = v.nl;
    const C one = mk<C>(1, 0);
    const C a_first = v.fl[t0 * nl + l];
    const C c_last = v.fc[(t1 - 1) * nl + l];
    C piv = v.fip[t0 * nl + l];
    C ip = cdiv(one, piv);
    v.fip[t0 * nl + l] = ip;
    v.fl[t0 * nl + l] = czero<C>();
    // forward elimination; V's forward sweep rides along (rhs = a_first * e_first)
    C yv = a_first;
    v.sv[t0 * nl + l] = yv;
    for (int t = t0 + 1; t < t1; ++t) {
        const C lm = cmul(v.fl[t * nl + l], ip);           // sub / previous pivot
        piv = csub(v.fip[t * nl + l], cmul(lm, v.fc[(t - 1) * nl + l]));
        ip = cdiv(one, piv);
        v.fl[t * nl + l] = lm;
        v.fip[t * nl + l] = ip;
        yv = cneg(cmul(lm, yv));
        v.sv[t * nl + l] = yv;
    }
    // backward substitution for both spikes (W: rhs = c_last * e_last, forward sweep is trivial)
    C ev = cmul(v.sv[(t1 - 1) * nl + l], ip);
    C ew = cmul(c_last, ip);
    v.sv[(t1 - 1) * nl + l] = ev;
    v.sw[(t1 - 1) * nl + l] = ew;
    for (int t = t1 - 2; t >= t0; --t) {
        const C c = v.fc[t * nl + l], q = v.fip[t * nl + l];
        ev = cmul(csub(v.sv[t * nl + l], cmul(c, ev)), q);
        ew = cmul(cneg(cmul(c, ew)), q);
        v.sv[t * nl + l] = ev;
        v.sw[t * nl + l] = ew;
    }
    // spike tips of this chunk, the only entries of the interface system
    C* tip = v.tips + ((int64_t)l * v.P + k) * 4;
    tip[0] = v.sv[t0 * nl + l]; tip[1] = v.sv[(t1 - 1) * nl + l];
    tip[2] = v.sw[t0 * nl + l]; tip[3] = v.sw[(t1 - 1) * nl + l];
}

// interface system R (2P x 2P, unknowns f_0, z_0, f_1, z_1, ...) and its inverse by Gauss-Jordan
// in shared memory; one CTA per line.  R = I + spike tips, diagonally dominant -> no pivoting.
template <typename C>
__global__ void line_reduced_inverse_kernel(LineView<C> v) {
    extern __shared__ unsigned char smem_raw[];
    C* R = reinterpret_cast<C*>(smem_raw);
    const int l = blockIdx.x;
    const int n = 2 * v.Ptot;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) R[e] = (e / n == e % n) ? mk<C>(1, 0) : czero<C>();
    __syncthreads();
    for (int kg = threadIdx.x; kg < v.Ptot; kg += blockDim.x) {
        const int r = kg / v.P, k = kg % v.P;
        const C* tip = v.tips_all + (((int64_t)r * v.nl + l) * v.P + k) * 4;
        if (kg > 0) {                 // coupling to z_{kg-1}
            R[(2 * kg) * n + 2 * (kg - 1) + 1] = tip[0];
            R[(2 * kg + 1) * n + 2 * (kg - 1) + 1] = tip[1];
        }
        if (kg < v.Ptot - 1) {        // coupling to f_{kg+1}
            R[(2 * kg) * n + 2 * (kg + 1)] = tip[2];
            R[(2 * kg + 1) * n + 2 * (kg + 1)] = tip[3];
        }
    }
    __syncthreads();
    __shared__ double2 piv_sh;
    for (int p = 0; p < n; ++p) {
        if (threadIdx.x == 0) {
            const C d = R[p * n + p];
            const C ip = cdiv(mk<C>(1, 0), d);
            piv_sh = make_double2((double)ip.x, (double)ip.y);
            R[p * n + p] = mk<C>(1, 0);
        }
        __syncthreads();
        const C ip = mk<C>((typename real_of<C>::type)piv_sh.x, (typename real_of<C>::type)piv_sh.y);
        for (int c = threadIdx.x; c < n; c += blockDim.x) R[p * n + c] = cmul(R[p * n + c], ip);
        __syncthreads();
        // eliminate column p from every other row: each thread owns rows r (loop) and all columns
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
            if (r == p) continue;
            const C f = R[r * n + p];
            R[r * n + p] = czero<C>();
            for (int c = 0; c < n; ++c) R[r * n + c] = csub(R[r * n + c], cmul(f, R[p * n + c]));
        }
        __syncthreads();
    }
    C* out = v.rinv + (int64_t)l * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) out[e] = R[e];   // row-major
}

// phase A: chunk-local solve, in place in g; interface values to ifg.  The recurrences are
// latency bound (one dependent complex FMA per step), so the operands of kLB steps are fetched
// into registers ahead of the dependent chain.
constexpr int kLB = 16;
template <typename C>
__global__ void __launch_bounds__(64) line_local_solve_kernel(LineView<C> v) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= v.nl * v.P) return;
    const int l = idx % v.nl, k = idx / v.nl;
    const int t0 = k * v.m, t1 = min(v.len, t0 + v.m);
    if (t0 >= t1) return;
    const int64_t nl = v.nl;
    C* g = v.g + l;
    const C* fl = v.fl + l;
    const C* fc = v.fc + l;
    const C* fip = v.fip + l;
    C y = g[t0 * nl];
    for (int tb = t0 + 1; tb < t1; tb += kLB) {
        C gb[kLB], lb[kLB];
#pragma unroll
        for (int q = 0; q < kLB; ++q) {
            const int t = min(tb + q, t1 - 1);
            gb[q] = g[t * nl];
            lb[q] = fl[t * nl];
        }
#pragma unroll
        for (int q = 0; q < kLB; ++q) {
            if (tb + q < t1) {
                y = csub(gb[q], cmul(lb[q], y));
                g[(tb + q) * nl] = y;
            }
        }
    }
    C e = cmul(y, fip[(t1 - 1) * nl]);
    g[(t1 - 1) * nl] = e;
    const C e_last = e;
    for (int tb = t1 - 2; tb >= t0; tb -= kLB) {
        C gb[kLB], cb[kLB], pb[kLB];
#pragma unroll
        for (int q = 0; q < kLB; ++q) {
            const int t = max(tb - q, t0);
            gb[q] = g[t * nl];
            cb[q] = fc[t * nl];
            pb[q] = fip[t * nl];
        }
#pragma unroll
        for (int q = 0; q < kLB; ++q) {
            if (tb - q >= t0) {
                e = cmul(csub(gb[q], cmul(cb[q], e)), pb[q]);
                g[(tb - q) * nl] = e;
            }
        }
    }
    if (v.Ptot > 1) {
        v.ifg[(int64_t)l * 2 * v.P + 2 * k] = e;
        v.ifg[(int64_t)l * 2 * v.P + 2 * k + 1] = e_last;
    }
}

// phase B: u = R^-1 g_interface ; one warp per (line, row), lanes across the row, shuffle reduce
template <typename C>
__global__ void __launch_bounds__(256) line_reduced_kernel(LineView<C> v) {
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n = 2 * v.Ptot, nloc = 2 * v.P;
    if (warp >= (int64_t)v.nl * n) return;
    const int l = (int)(warp / n), row = (int)(warp % n);
    const C* Ri = v.rinv + ((int64_t)l * n + row) * n;
    double ax = 0.0, ay = 0.0;
    for (int q = lane; q < n; q += 32) {
        // gathered interface values are rank-major: [rank][line][2P]
        const C r = Ri[q], g = v.ifg_all[((int64_t)(q / nloc) * v.nl + l) * nloc + (q % nloc)];
        ax += (double)r.x * g.x - (double)r.y * g.y;
        ay += (double)r.x * g.y + (double)r.y * g.x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ax += __shfl_down_sync(0xffffffffu, ax, off);
        ay += __shfl_down_sync(0xffffffffu, ay, off);
    }
    if (lane == 0) v.ifu[(int64_t)l * n + row] = mk<C>((typename real_of<C>::type)ax, (typename real_of<C>::type)ay);
}

// phase C: spike correction; dir 0 also applies the update x += omega * e
template <typename C>
__global__ void line_correct_kernel(LineView<C> v, C* x, typename real_of<C>::type omega) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)v.nl * v.len) return;
    const int l = (int)(idx % v.nl), t = (int)(idx / v.nl);
    C e = v.g[idx];
    if (v.Ptot > 1) {
        const int kg = v.k0 + t / v.m;
        const C* u = v.ifu + (int64_t)l * 2 * v.Ptot;
        if (kg > 0) e = csub(e, cmul(v.sv[idx], u[2 * (kg - 1) + 1]));
        if (kg < v.Ptot - 1) e = csub(e, cmul(v.sw[idx], u[2 * (kg + 1)]));
    }
    if (v.dir == 0) {
        int i, j;
        line_point(v, l, t, i, j);
        const int64_t n = (int64_t)j * v.Nx + i;
        x[n] = cadd(x[n], cscale(omega, e));
    } else {
        v.g[idx] = e;
    }
}

// x-lines: residual of the strip rows, written transposed ([t = i][line = row]) via a smem tile
template <typename C, bool TM>
__global__ void __launch_bounds__(256) strip_rows_resid_kernel(LineView<C> v, const C* x, const C* b,
        const C* ca, const C* cb, const C* cd, typename real_of<C>::type sx,
        typename real_of<C>::type sy, C shift) {
    __shared__ C tile[32][33];
    const int i0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int l = l0 + r, i = i0 + tx;
        C val = czero<C>();
        if (l < v.nl && i < v.Nx) {
            const int j = l < v.lo ? l : v.Ny - v.hi + (l - v.lo);
            PointCoefs<C> p = point_coefs<C, TM>(ca, cb, cd, v.Nx, v.Ny, sx, sy, shift, i, j);
            const int64_t n = (int64_t)j * v.Nx + i;
            C ax = cmul(p.dg, x[n]);
            if (i > 0) ax = cfma(p.kw, x[n - 1], ax);
            if (i < v.Nx - 1) ax = cfma(p.ke, x[n + 1], ax);
            if (j > 0) ax = cfma(p.ks, x[n - v.Nx], ax);
            if (j < v.Ny - 1) ax = cfma(p.kn, x[n + v.Nx], ax);
            val = csub(b[n], ax);
        }
        tile[r][tx] = val;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, l = l0 + tx;
        if (l < v.nl && i < v.Nx) v.g[(int64_t)i * v.nl + l] = tile[tx][r];
    }
}

// x-lines: x[row(l)][i] += omega * e[i][l]   (transposed through a smem tile)
template <typename C>
__global__ void __launch_bounds__(256) strip_rows_scatter_kernel(LineView<C> v, C* x,
        typename real_of<C>::type omega) {
    __shared__ C tile[32][33];
    const int i0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, l = l0 + tx;
        tile[r][tx] = (l < v.nl && i < v.Nx) ? v.g[(int64_t)i * v.nl + l] : czero<C>();
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int l = l0 + r, i = i0 + tx;
        if (l < v.nl && i < v.Nx) {
            const int j = l < v.lo ? l : v.Ny - v.hi + (l - v.lo);
            const int64_t n = (int64_t)j * v.Nx + i;
            x[n] = cadd(x[n], cscale(omega, tile[tx][r]));
        }
    }
}

// Fused line relaxation for short, unsharded lines (coarse levels): one CTA per line does what the
// five kernels above do for that line — (x-lines: residual of the strip row) -> chunk-local solves ->
// interface mat-vec -> spike correction -> x += omega e — with the line staged in shared memory.
// On the coarse levels every one of those kernels is a few microseconds of launch latency for a
// handful of CTAs; a smoothing sweep drops from nine launches to three.
constexpr int kFusedMaxLen = 1024;
constexpr int kFusedThreads = 256;
constexpr int kFusedRows = 2 * kMaxChunks / (kFusedThreads / 32);   // interface rows per warp (8)
constexpr int kFusedPts = kFusedMaxLen / kFusedThreads;             // line points per thread (4)

template <typename C, bool TM, int DIR>
__global__ void __launch_bounds__(kFusedThreads) line_fused_kernel(LineView<C> v, C* x, const C* b,
        const C* ca, const C* cb, const C* cd, typename real_of<C>::type sx, typename real_of<C>::type sy,
        C shift, typename real_of<C>::type omega) {
    // dynamic shared memory: the line (right-hand side, then solution) and its three factor arrays
    extern __shared__ __align__(16) unsigned char fused_smem[];
    __shared__ C sg[2 * kMaxChunks], su[2 * kMaxChunks];
    const int l = blockIdx.x;
    const int64_t nl = v.nl;
    const int len = v.len;
    C* sl = reinterpret_cast<C*>(fused_smem);
    C* s_fl = sl + len; C* s_fc = s_fl + len; C* s_ip = s_fc + len;
    const int fixed = l < v.lo ? l : (DIR == 0 ? v.Nx : v.Ny) - v.hi + (l - v.lo);   // column (DIR 0) / row (DIR 1)
    const int n = 2 * v.Ptot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // everything the later (latency-bound) phases need is requested up front: the factor arrays go to
    // shared memory, this warp's rows of the interface inverse and this thread's spike entries to registers
    for (int t = threadIdx.x; t < len; t += kFusedThreads) {
        s_fl[t] = v.fl[t * nl + l]; s_fc[t] = v.fc[t * nl + l]; s_ip[t] = v.fip[t * nl + l];
    }
    C rreg[kFusedRows][2], svr[kFusedPts], swr[kFusedPts];
    if (v.Ptot > 1) {
#pragma unroll
        for (int r = 0; r < kFusedRows; ++r) {
            const int row = warp + r * (kFusedThreads / 32);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int q = lane + 32 * c;
                rreg[r][c] = (row < n && q < n) ? v.rinv[((int64_t)l * n + row) * n + q] : czero<C>();
            }
        }
#pragma unroll
        for (int q = 0; q < kFusedPts; ++q) {
            const int t = threadIdx.x + q * kFusedThreads;
            svr[q] = t < len ? v.sv[t * nl + l] : czero<C>();
            swr[q] = t < len ? v.sw[t * nl + l] : czero<C>();
        }
    }
    if (DIR == 0) {
        for (int t = threadIdx.x; t < len; t += kFusedThreads) sl[t] = v.g[t * nl + l];   // written by the SMOOTH pass
    } else {
        const int j = fixed;
        for (int i = threadIdx.x; i < len; i += kFusedThreads) {
            PointCoefs<C> p = point_coefs<C, TM>(ca, cb, cd, v.Nx, v.Ny, sx, sy, shift, i, j);
            const int64_t pn = (int64_t)j * v.Nx + i;
            C ax = cmul(p.dg, x[pn]);
            if (i > 0) ax = cfma(p.kw, x[pn - 1], ax);
            if (i < v.Nx - 1) ax = cfma(p.ke, x[pn + 1], ax);
            if (j > 0) ax = cfma(p.ks, x[pn - v.Nx], ax);
            if (j < v.Ny - 1) ax = cfma(p.kn, x[pn + v.Nx], ax);
            sl[i] = csub(b[pn], ax);
        }
    }
    __syncthreads();
    // chunk-local solves, one thread per chunk, operands in shared memory
    if ((int)threadIdx.x < v.P) {
        const int k = threadIdx.x;
        const int t0 = k * v.m, t1 = min(len, t0 + v.m);
        if (t0 < t1) {
            C y = sl[t0];
            for (int t = t0 + 1; t < t1; ++t) { y = csub(sl[t], cmul(s_fl[t], y)); sl[t] = y; }
            C e = cmul(y, s_ip[t1 - 1]);
            sl[t1 - 1] = e;
            const C e_last = e;
            for (int t = t1 - 2; t >= t0; --t) { e = cmul(csub(sl[t], cmul(s_fc[t], e)), s_ip[t]); sl[t] = e; }
            sg[2 * k] = e; sg[2 * k + 1] = e_last;
        }
    }
    __syncthreads();
    if (v.Ptot > 1) {
        // interface mat-vec u = R^-1 g: one warp per row, the row already sits in registers
#pragma unroll
        for (int r = 0; r < kFusedRows; ++r) {
            const int row = warp + r * (kFusedThreads / 32);
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int q = lane + 32 * c;
                if (q < n) {
                    const C rv = rreg[r][c], g = sg[q];
                    ax += (double)rv.x * g.x - (double)rv.y * g.y;
                    ay += (double)rv.x * g.y + (double)rv.y * g.x;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ax += __shfl_down_sync(0xffffffffu, ax, off);
                ay += __shfl_down_sync(0xffffffffu, ay, off);
            }
            if (lane == 0 && row < n) su[row] = mk<C>((typename real_of<C>::type)ax, (typename real_of<C>::type)ay);
        }
        __syncthreads();
    }
    // spike correction and update.  x-lines of one strip are adjacent rows whose residuals were read
    // by the neighbouring CTAs above: their update goes through g and line_fused_apply_kernel, which
    // keeps the Jacobi-between-lines semantics of the separate kernels; y-lines own whole columns of
    // the SMOOTH pass output, which no other CTA reads here.
#pragma unroll
    for (int q = 0; q < kFusedPts; ++q) {
        const int t = threadIdx.x + q * kFusedThreads;
        if (t < len) {
            C e = sl[t];
            if (v.Ptot > 1) {
                const int kg = t / v.m;
                if (kg > 0) e = csub(e, cmul(svr[q], su[2 * (kg - 1) + 1]));
                if (kg < v.Ptot - 1) e = csub(e, cmul(swr[q], su[2 * (kg + 1)]));
            }
            if (DIR == 0) {
                const int64_t pn = (int64_t)t * v.Nx + fixed;
                x[pn] = cadd(x[pn], cscale(omega, e));
            } else {
                v.g[t * nl + l] = e;
            }
        }
    }
}

// x-lines, second phase: x[row(l)][i] += omega * e[i][l] once every line has been solved
template <typename C>
__global__ void __launch_bounds__(kFusedThreads) line_fused_apply_kernel(LineView<C> v, C* x,
        typename real_of<C>::type omega) {
    const int l = blockIdx.x;
    const int j = l < v.lo ? l : v.Ny - v.hi + (l - v.lo);
    for (int i = threadIdx.x; i < v.len; i += blockDim.x) {
        const int64_t n = (int64_t)j * v.Nx + i;
        x[n] = cadd(x[n], cscale(omega, v.g[(int64_t)i * v.nl + l]));
    }
}

#include "mg_coarse.cuh"

// ---------------------------------------------------------------------------------------------
// grid transfer and coarse coefficients
template <typename C, bool TM>
__global__ void coarsen_coefs_kernel(int Nxc, int Nyc, const C* ca, const C* cb, const C* cd,
        C* cac, C* cbc, C* cdc, typename real_of<C>::type wall_mult, bool wall_lo_y, bool wall_hi_y) {
    using R = typename real_of<C>::type;
    const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= Nxc || J >= Nyc) return;
    const int Nx = 2 * Nxc;
    const int ix = TM ? 0 : 1;
    const int64_t f00 = (int64_t)(2 * J) * Nx + 2 * I;
    C a = cscale((R)0.5, cadd(ca[f00 + ix], ca[f00 + Nx + ix]));
    C b = cscale((R)0.5, cadd(cb[f00 + (int64_t)ix * Nx], cb[f00 + (int64_t)ix * Nx + 1]));
    // (y-slab sharding: only the rank that owns the global wall row scales its cb)
    if (TM) { if (I == 0) a = cscale(wall_mult, a); if (J == 0 && wall_lo_y) b = cscale(wall_mult, b); }
    else { if (I == Nxc - 1) a = cscale(wall_mult, a); if (J == Nyc - 1 && wall_hi_y) b = cscale(wall_mult, b); }
    const int64_t c = (int64_t)J * Nxc + I;
    cac[c] = a; cbc[c] = b;
    if (cd) cdc[c] = cscale((R)0.25, cadd(cadd(cd[f00], cd[f00 + 1]), cadd(cd[f00 + Nx], cd[f00 + Nx + 1])));
}

template <typename C>
__global__ void restrict_kernel(int Nxc, int Nyc, const C* rf, C* rc) {
    using R = typename real_of<C>::type;
    const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= Nxc || J >= Nyc) return;
    const int Nx = 2 * Nxc;
    const int64_t f = (int64_t)(2 * J) * Nx + 2 * I;
    rc[(int64_t)J * Nxc + I] = cscale((R)0.25, cadd(cadd(rf[f], rf[f + 1]), cadd(rf[f + Nx], rf[f + Nx + 1])));
}

// x_f += P e_c ; bilinear, clamped on the flux-free side, linear to zero at the wall.
// ec_south / ec_north: coarse ghost rows J = -1 / J = Nyc owned by the neighbouring ranks (null at a
// global boundary, where the wall / clamp rule applies).
template <typename C, bool TM>
__global__ void prolong_add_kernel(int Nxc, int Nyc, const C* ec, C* xf, typename real_of<C>::type wq,
                                   const C* ec_south, const C* ec_north) {
    using R = typename real_of<C>::type;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    const int Nx = 2 * Nxc, Ny = 2 * Nyc;
    if (i >= Nx || j >= Ny) return;
    // 1-D weights: value = w0 * e[I0] + w1 * e[I1]; `ghost` = I1 lies on a neighbouring rank
    auto axis = [&](int f, int Nc, bool ghost_lo, bool ghost_hi, int& I0, int& I1, R& w0, R& w1) {
        const int I = f >> 1;
        I0 = I; w0 = (R)0.75; w1 = (R)0.25;
        I1 = (f & 1) ? I + 1 : I - 1;
        if (I1 < 0 && !ghost_lo) { if (TM) { w0 = wq; w1 = 0; } I1 = 0; if (!TM) { I1 = I; } }
        else if (I1 >= Nc && !ghost_hi) { if (!TM) { w0 = wq; w1 = 0; } I1 = Nc - 1; if (TM) { I1 = I; } }
    };
    int I0, I1, J0, J1; R wx0, wx1, wy0, wy1;
    axis(i, Nxc, false, false, I0, I1, wx0, wx1);
    axis(j, Nyc, ec_south != nullptr, ec_north != nullptr, J0, J1, wy0, wy1);
    const C* row0 = ec + (int64_t)J0 * Nxc;
    const C* row1 = J1 < 0 ? ec_south : (J1 >= Nyc ? ec_north : ec + (int64_t)J1 * Nxc);
    C v = cscale(wy0, cadd(cscale(wx0, row0[I0]), cscale(wx1, row0[I1])));
    v = cadd(v, cscale(wy1, cadd(cscale(wx0, row1[I0]), cscale(wx1, row1[I1]))));
    const int64_t n = (int64_t)j * Nx + i;
    xf[n] = cadd(xf[n], v);
}

// ---------------------------------------------------------------------------------------------
template <typename C>
struct Level {
    Helm2D op;
    C *ca = nullptr, *cb = nullptr, *cd = nullptr;   // owned when own_coefs
    bool own_coefs = false;
    C *x = nullptr, *b = nullptr, *r = nullptr, *t = nullptr;
    C *x0 = nullptr, *r0 = nullptr, *t0 = nullptr;   // allocation-time roles (restored per apply)
    bool own_xb = true;
    double delta = 1.0;
    LineSet<C> ylines, xlines;
    bool first = true, last = true;   // this rank owns the global low-y / high-y wall
    SlabEdge<C> edge() const {
        return SlabEdge<C>{!first, !last, (const C*)op.cb_ghost_buf};
    }
};

template <typename C>
struct MGPrecond : Precond {
    std::vector<Level<C>> L;
    fdfd_krylov_opts o;
    double shift_re = 1.0, shift_im = 0.5;
    int coarse_its = 8;
    BicgWorkspace coarse_ws;
    int nlev = 0;       // levels of the cycle = L[0 .. nlev)
    int rep = -1;       // index in L of the replicated (unsharded) copy of the coarsest level, or -1
    GatherPlan* rep_plan = nullptr;   // peer-memory all-gather of the coarse right-hand side
    Comm* rep_comm = nullptr;
    // CUDA-graph replay of the cycle, one executable graph per (input, output) pointer pair
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaGraphExec_t graph = nullptr;
    C* zero_buf = nullptr;     // all-zero vector of the coarsest level's size (initial guess of its smoother)
    C* in_buf = nullptr;       // fixed right-hand-side buffer of the captured cycle
    C* out_ptr = nullptr;      // where the captured cycle leaves its result (level-0 x after the swaps)

    // preconditioner of the coarsest-level Krylov solve: one smoothing sweep from a zero guess
    struct SmoothPrecond : Precond {
        MGPrecond* mg = nullptr;
        int level = 0;
        int apply(const void* v, void* z, cudaStream_t st) override {
            // one sweep from a zero guess, written straight into z: the sweep reads lv.x (a buffer that
            // stays zero: the kernels only write their output array) and leaves its result in lv.t
            Level<C>& lv = mg->L[level];
            C *sb = lv.b, *sx = lv.x, *stt = lv.t;
            lv.b = (C*)v; lv.x = mg->zero_buf; lv.t = (C*)z;
            int s = mg->smooth(level, st);
            lv.b = sb; lv.x = sx; lv.t = stt;
            return s;
        }
    } coarse_pre;

    ~MGPrecond() override {
        if (graph) cudaGraphExecDestroy(graph);
        if (rep_plan) p2p_gather_destroy(rep_comm, rep_plan);
        cudaFree(in_buf);
        cudaFree(zero_buf);
        cudaFree(coarse_partials); cudaFree(coarse_bar); cudaFree(coarse_extra[0]); cudaFree(coarse_extra[1]);
        cudaFree(coarse_k5); cudaFree(coarse_relax);
        if (coarse_sh.seq) {
            cudaFree(coarse_sh.seq); cudaFree(coarse_tx_table);
            for (int q = 0; q < 6; ++q) p2p_free(coarse_comm, coarse_offs[q]);
        }
        if (gstream) cudaStreamDestroy(gstream);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
        for (size_t l = 0; l < L.size(); ++l) {
            Level<C>& v = L[l];
            if (v.own_coefs) { cudaFree(v.ca); cudaFree(v.cb); cudaFree(v.cd); }
            if (l > 0) cudaFree(v.b);
            cudaFree(v.x0); cudaFree(v.r0); cudaFree(v.t0);
            helm2d_release_sharding(v.op);
            v.ylines.release(); v.xlines.release();
        }
    }

    bool TM() const { return L[0].op.pol == FDFD_POL_TM; }
    C shift() const { return mk<C>((typename real_of<C>::type)shift_re, (typename real_of<C>::type)shift_im); }

    int setup_lines(Level<C>& lv, LineSet<C>& s, int dir, int lo, int hi, cudaStream_t st) {
        using R = typename real_of<C>::type;
        const int Nx = lv.op.Nx, Ny = lv.op.Ny;
        Comm* comm = lv.op.comm;
        const bool sharded = comm && comm->nranks > 1;
        const int across = dir == 0 ? Nx : Ny;
        if (lo + hi > across) {
            FDFD_CHECK_ARG(!(sharded && dir == 1), "absorbing-layer rows exceed the rows of one slab");
            lo = across; hi = 0;
        }
        FDFD_CHECK_ARG(!(sharded && dir == 1 && (lo >= Ny || hi >= Ny)), "absorbing-layer rows exceed the rows of one slab");
        s.dir = dir; s.lo = lo; s.hi = hi; s.nl = lo + hi;
        s.len = dir == 0 ? Ny : Nx;
        // y-lines of a sharded grid run through every rank: each rank owns P chunks of its own rows
        s.nranks = (dir == 0 && sharded) ? comm->nranks : 1;
        if (s.nl == 0) return FDFD_OK;
        const int len_global = s.len * s.nranks;
        int Pglob = len_global >= 512 ? kMaxChunks : (len_global >= 128 ? 8 : 1);
        int P = Pglob / s.nranks;
        if (P < 1) P = 1;
        FDFD_CHECK_ARG(P * s.nranks <= kMaxChunks, "too many ranks for the partitioned line solver");
        s.m = (s.len + P - 1) / P;
        s.P = (s.len + s.m - 1) / s.m;
        s.Ptot = s.P * s.nranks;
        s.k0 = (s.nranks > 1 ? comm->rank : 0) * s.P;
        const size_t e = (size_t)s.len * s.nl * sizeof(C);
        FDFD_CUDA(cudaMalloc(&s.fl, e)); FDFD_CUDA(cudaMalloc(&s.fc, e)); FDFD_CUDA(cudaMalloc(&s.fip, e));
        FDFD_CUDA(cudaMalloc(&s.sv, e)); FDFD_CUDA(cudaMalloc(&s.sw, e)); FDFD_CUDA(cudaMalloc(&s.g, e));
        FDFD_CUDA(cudaMemsetAsync(s.sv, 0, e, st)); FDFD_CUDA(cudaMemsetAsync(s.sw, 0, e, st));
        const size_t nloc = (size_t)s.nl * 2 * s.P, ntot = (size_t)s.nl * 2 * s.Ptot;
        FDFD_CUDA(cudaMalloc(&s.ifg, nloc * sizeof(C)));
        FDFD_CUDA(cudaMalloc(&s.ifg_all, ntot * sizeof(C)));
        FDFD_CUDA(cudaMalloc(&s.ifu, ntot * sizeof(C)));
        FDFD_CUDA(cudaMalloc(&s.rinv, ntot * 2 * s.Ptot * sizeof(C)));
        FDFD_CUDA(cudaMalloc(&s.tips, (size_t)s.nl * s.P * 4 * sizeof(C)));
        FDFD_CUDA(cudaMalloc(&s.tips_all, (size_t)s.nl * s.Ptot * 4 * sizeof(C)));
        LineView<C> v = view(s, Nx, Ny);
        const int64_t tot = (int64_t)s.nl * s.len;
        const int blocks = (int)((tot + 255) / 256);
        if (TM()) line_extract_kernel<C, true><<<blocks, 256, 0, st>>>(v, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), lv.edge());
        else line_extract_kernel<C, false><<<blocks, 256, 0, st>>>(v, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), lv.edge());
        line_factor_kernel<C><<<(s.nl * s.P + 127) / 128, 128, 0, st>>>(v);
        FDFD_CUDA(cudaGetLastError());
        // every rank needs the spike tips of all chunks to build (and invert) the interface system
        FDFD_TRY(comm_allgather(s.nranks > 1 ? comm : nullptr, s.tips, s.tips_all,
                                (size_t)s.nl * s.P * 4 * sizeof(C), st));
        if (s.nranks > 1 && p2p_enabled(comm)) {
            s.plan_comm = comm;
            FDFD_TRY(p2p_gather_create(comm, nloc * sizeof(C), &s.plan, st));
        }
        if (s.Ptot > 1) {
            const size_t smem = (size_t)4 * s.Ptot * s.Ptot * sizeof(C);
            FDFD_CUDA(cudaFuncSetAttribute(line_reduced_inverse_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            line_reduced_inverse_kernel<C><<<s.nl, 128, smem, st>>>(v);
        }
        FDFD_CUDA(cudaGetLastError());
        return FDFD_OK;
    }

    int line_solve(Level<C>& lv, LineSet<C>& s, C* x, double omega, cudaStream_t st) {
        using R = typename real_of<C>::type;
        LineView<C> v = view(s, lv.op.Nx, lv.op.Ny);
        line_local_solve_kernel<C><<<(s.nl * s.P + 63) / 64, 64, 0, st>>>(v);
        if (s.Ptot > 1) {
            if (s.plan) FDFD_TRY(p2p_allgather(s.plan, s.ifg, s.ifg_all, (size_t)s.nl * 2 * s.P * sizeof(C), st));
            else FDFD_TRY(comm_allgather(s.nranks > 1 ? lv.op.comm : nullptr, s.ifg, s.ifg_all,
                                         (size_t)s.nl * 2 * s.P * sizeof(C), st));
            line_reduced_kernel<C><<<(int)(((int64_t)s.nl * 2 * s.Ptot * 32 + 255) / 256), 256, 0, st>>>(v);
        }
        const int64_t tot = (int64_t)s.nl * s.len;
        if (s.Ptot > 1 || s.dir == 0)
            line_correct_kernel<C><<<(int)((tot + 255) / 256), 256, 0, st>>>(v, x, (R)omega);
        if (s.dir == 1) {
            dim3 grid((lv.op.Nx + 31) / 32, (s.nl + 31) / 32);
            strip_rows_scatter_kernel<C><<<grid, 256, 0, st>>>(v, x, (R)omega);
        }
        FDFD_CUDA(cudaGetLastError());
        return FDFD_OK;
    }

    // one smoothing sweep on level l (x <- smoothed x); uses lv.t as ping-pong buffer
    int smooth(int l, cudaStream_t st) {
        using R = typename real_of<C>::type;
        Level<C>& lv = L[l];
        if (!lv.op.halo.state) FDFD_TRY(helm2d_halo_exchange(lv.op, lv.x, st));   // peer memory: inside the launch
        HelmArgs<C> a = helm2d_args<C>(lv.op, lv.x, lv.t, lv.b, lv.ca, lv.cb, lv.cd, shift_re, shift_im, o.mg_omega);
        a.strip_lo = lv.ylines.lo; a.strip_hi = lv.ylines.hi; a.line_rhs = lv.ylines.g; a.row0 = 0;
        FDFD_TRY(helm2d_launch_args<C>(lv.op, APPLY_SMOOTH, a, st));
        if (lv.ylines.nl) {
            if (fused_ok(lv.ylines)) {
                LineView<C> v = view(lv.ylines, lv.op.Nx, lv.op.Ny);
                if (TM()) line_fused_kernel<C, true, 0><<<lv.ylines.nl, kFusedThreads, fused_smem(lv.ylines), st>>>(v, lv.t, lv.b, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), (R)o.mg_omega);
                else line_fused_kernel<C, false, 0><<<lv.ylines.nl, kFusedThreads, fused_smem(lv.ylines), st>>>(v, lv.t, lv.b, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), (R)o.mg_omega);
                FDFD_CUDA(cudaGetLastError());
            } else {
                FDFD_TRY(line_solve(lv, lv.ylines, lv.t, o.mg_omega, st));
            }
        }
        std::swap(lv.x, lv.t);
        if (lv.xlines.nl) {
            LineView<C> v = view(lv.xlines, lv.op.Nx, lv.op.Ny);
            if (fused_ok(lv.xlines)) {
                if (TM()) line_fused_kernel<C, true, 1><<<lv.xlines.nl, kFusedThreads, fused_smem(lv.xlines), st>>>(v, lv.x, lv.b, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), (R)o.mg_omega);
                else line_fused_kernel<C, false, 1><<<lv.xlines.nl, kFusedThreads, fused_smem(lv.xlines), st>>>(v, lv.x, lv.b, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), (R)o.mg_omega);
                line_fused_apply_kernel<C><<<lv.xlines.nl, kFusedThreads, 0, st>>>(v, lv.x, (R)o.mg_omega);
                FDFD_CUDA(cudaGetLastError());
            } else {
                dim3 grid((lv.op.Nx + 31) / 32, (lv.xlines.nl + 31) / 32);
                if (TM()) strip_rows_resid_kernel<C, true><<<grid, 256, 0, st>>>(v, lv.x, lv.b, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift());
                else strip_rows_resid_kernel<C, false><<<grid, 256, 0, st>>>(v, lv.x, lv.b, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift());
                FDFD_TRY(line_solve(lv, lv.xlines, lv.x, o.mg_omega, st));
            }
        }
        return FDFD_OK;
    }

    // short lines that live on one rank go through the fused one-CTA-per-line kernel
    static size_t fused_smem(const LineSet<C>& s) { return (size_t)4 * s.len * sizeof(C); }
    static int fused_attr() {
        const int bytes = 4 * kFusedMaxLen * (int)sizeof(C);
        FDFD_CUDA(cudaFuncSetAttribute(line_fused_kernel<C, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FDFD_CUDA(cudaFuncSetAttribute(line_fused_kernel<C, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FDFD_CUDA(cudaFuncSetAttribute(line_fused_kernel<C, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FDFD_CUDA(cudaFuncSetAttribute(line_fused_kernel<C, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        return FDFD_OK;
    }
    bool fused_ok(const LineSet<C>& s) const {
        static const bool off = getenv("FDFD_NO_FUSED_LINES") != nullptr;
        return !off && s.nranks == 1 && s.len <= kFusedMaxLen && s.len <= o.mg_fuse_len && s.P <= kFusedThreads;
    }

    // ---- mg_coarse_mode = 2: the whole coarsest-level solve as one persistent kernel (mg_coarse.cuh) ----
    double* coarse_partials = nullptr;
    unsigned int* coarse_bar = nullptr;
    C* coarse_extra[2] = {nullptr, nullptr};
    C *coarse_k5 = nullptr, *coarse_relax = nullptr;
    int coarse_grid = 0;
    size_t coarse_smem_launch = 0;
    bool coarse_spikes = false;
    bool coarse_ready = false;
    static size_t coarse_smem_bytes(const Level<C>& lv, bool spikes) {
        const int len = std::max(lv.ylines.nl ? line_pad_len(lv.ylines.len, lv.ylines.m) : 0,
                                 lv.xlines.nl ? line_pad_len(lv.xlines.len, lv.xlines.m) : 0);
        return (line_smem_elems<C>(lv.ylines.nl, lv.ylines.len, lv.ylines.m, lv.ylines.Ptot, spikes) +
                line_smem_elems<C>(lv.xlines.nl, lv.xlines.len, lv.xlines.m, lv.xlines.Ptot, spikes) + (size_t)len) * sizeof(C);
    }
    // shape limits of the persistent kernel; anything else keeps the launch chain of mode 1
    // (every condition is the same on all ranks of a sharded level: equal slabs, global strip widths)
    bool coarse_persistent_shape_ok(const Level<C>& lv) const {
        if (o.mg_coarse_mode != 2) return false;
        const bool sharded = lv.op.Ny != lv.op.Ny_global;
        if (sharded && !(lv.op.comm && p2p_enabled(lv.op.comm))) return false;   // needs the peer-mapped heap
        if ((int64_t)lv.op.Nx * lv.op.Ny >= ((int64_t)1 << 31)) return false;
        if (lv.ylines.nl && lv.ylines.nranks != (sharded ? lv.op.comm->nranks : 1)) return false;
        if (lv.xlines.nl && lv.xlines.nranks != 1) return false;
        if (lv.ylines.Ptot > kMaxChunks || lv.xlines.Ptot > kMaxChunks) return false;
        if (std::max(lv.ylines.P, lv.xlines.P) > kCoarseThreads) return false;
        // shared memory is sized for the widest line sets any rank of the level has (x-lines exist on the first
        // and last rank only, with the global strip widths)
        const size_t worst = (line_smem_elems<C>(1, lv.ylines.len, std::max(lv.ylines.m, 1), std::max(lv.ylines.Ptot, 1), false) +
                              line_smem_elems<C>(1, lv.op.Nx, std::max((lv.op.Nx + kMaxChunks - 1) / kMaxChunks, 1), kMaxChunks, false) +
                              (size_t)line_pad_len(std::max(lv.op.Nx, lv.ylines.len), 1)) * sizeof(C);
        return (sharded ? worst : coarse_smem_bytes(lv, false)) <= (size_t)216 * 1024;
    }
    bool coarse_persistent_ok(const Level<C>& lv) const { return coarse_ready && coarse_persistent_shape_ok(lv); }
    // sharded level: message buffers in the symmetric heap
    CoarseShard<C> coarse_sh;
    size_t coarse_offs[6] = {0, 0, 0, 0, 0, 0};
    Comm* coarse_comm = nullptr;
    void** coarse_tx_table = nullptr;
    int coarse_setup_sharded(const Level<C>& lv, cudaStream_t st) {
        Comm* comm = lv.op.comm;
        const int P = comm->nranks, me = comm->rank, G = coarse_grid;
        const size_t nl = std::max(lv.ylines.nl, 1), nloc = 2 * (size_t)std::max(lv.ylines.P, 1);
        const size_t bytes[6] = {
            (size_t)2 * 2 * lv.op.Nx * sizeof(C),                    // ghost rows
            (size_t)2 * 2 * G * sizeof(unsigned int),                // their flags
            (size_t)2 * P * kCoarseVals * sizeof(double),            // reduction slots
            (size_t)2 * P * sizeof(unsigned int),
            (size_t)2 * P * nl * nloc * sizeof(C),                   // interface values of the y-lines
            (size_t)2 * P * nl * sizeof(unsigned int)};
        coarse_comm = comm;
        for (int q = 0; q < 6; ++q) FDFD_TRY(p2p_alloc(comm, bytes[q], &coarse_offs[q]));
        void* ptrs[6 * kP2PMaxRanks];
        FDFD_TRY(p2p_exchange(comm, coarse_offs, 6, ptrs, st));
        CoarseShard<C>& sh = coarse_sh;
        sh.nranks = P; sh.rank = me;
        sh.open_s = !lv.first; sh.open_n = !lv.last;
        sh.ghost_rx = (C*)ptrs[6 * me + 0]; sh.gflag_rx = (unsigned int*)ptrs[6 * me + 1];
        sh.red_rx = (double*)ptrs[6 * me + 2]; sh.rflag_rx = (unsigned int*)ptrs[6 * me + 3];
        sh.ifg_rx = (C*)ptrs[6 * me + 4]; sh.iflag_rx = (unsigned int*)ptrs[6 * me + 5];
        if (me > 0) { sh.ghost_tx_s = (C*)ptrs[6 * (me - 1) + 0]; sh.gflag_tx_s = (unsigned int*)ptrs[6 * (me - 1) + 1]; }
        if (me < P - 1) { sh.ghost_tx_n = (C*)ptrs[6 * (me + 1) + 0]; sh.gflag_tx_n = (unsigned int*)ptrs[6 * (me + 1) + 1]; }
        void* table[4][kP2PMaxRanks] = {};
        for (int r = 0; r < P; ++r)
            for (int q = 0; q < 4; ++q) table[q][r] = ptrs[6 * r + 2 + q];
        FDFD_CUDA(cudaMalloc(&coarse_tx_table, sizeof(table)));
        FDFD_CUDA(cudaMemcpy(coarse_tx_table, table, sizeof(table), cudaMemcpyHostToDevice));
        sh.red_tx = (double* const*)(coarse_tx_table + 0 * kP2PMaxRanks);
        sh.rflag_tx = (unsigned int* const*)(coarse_tx_table + 1 * kP2PMaxRanks);
        sh.ifg_tx = (C* const*)(coarse_tx_table + 2 * kP2PMaxRanks);
        sh.iflag_tx = (unsigned int* const*)(coarse_tx_table + 3 * kP2PMaxRanks);
        FDFD_CUDA(cudaMalloc(&sh.seq, 4 * sizeof(unsigned int)));
        FDFD_CUDA(cudaMemset(sh.seq, 0, 4 * sizeof(unsigned int)));
        return FDFD_OK;
    }

    using CoarseKernel = void (*)(CoarseArgs<C>);
    CoarseKernel coarse_kernel(bool sharded) const {
        if (sharded) return TM() ? coarse_bicgstab_kernel<C, true, true> : coarse_bicgstab_kernel<C, false, true>;
        return TM() ? coarse_bicgstab_kernel<C, true, false> : coarse_bicgstab_kernel<C, false, false>;
    }

    int coarse_setup(const Level<C>& lv) {
        const bool sharded = lv.op.Ny != lv.op.Ny_global;
        coarse_spikes = !sharded && coarse_smem_bytes(lv, true) <= (size_t)216 * 1024;
        const size_t smem = sharded ? (size_t)216 * 1024 : coarse_smem_bytes(lv, coarse_spikes);
        auto kern = coarse_kernel(sharded);
        FDFD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, dev = 0, sms = 0, coop = 0;
        FDFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCoarseThreads, smem));
        FDFD_CUDA(cudaGetDevice(&dev));
        FDFD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FDFD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        const int owners = std::max(lv.ylines.nl, lv.xlines.nl);
        // (sharded: the decision must be the same on every rank — all B200s of a box, same kernel, same smem)
        if (per_sm < 1 || !coop || (sharded ? std::max(lv.ylines.nl, kMaxChunks) : owners) > sms) return FDFD_OK;   // launch chain
        const int64_t want = ((int64_t)lv.op.Nx * lv.op.Ny + kCoarseThreads - 1) / kCoarseThreads;
        coarse_grid = sharded ? sms : (int)std::max<int64_t>(owners, std::min<int64_t>(sms, want));   // one CTA per SM at most
        coarse_smem_launch = smem;
        const size_t bytes = (size_t)lv.op.n_local() * sizeof(C);
        FDFD_CUDA(cudaMalloc(&coarse_partials, (size_t)coarse_grid * kCoarseVals * sizeof(double)));
        FDFD_CUDA(cudaMalloc(&coarse_bar, 4 * sizeof(unsigned int)));
        FDFD_CUDA(cudaMemset(coarse_bar, 0, 4 * sizeof(unsigned int)));
        FDFD_CUDA(cudaMalloc(&coarse_extra[0], bytes));
        FDFD_CUDA(cudaMalloc(&coarse_extra[1], bytes));
        FDFD_CUDA(cudaMalloc(&coarse_k5, 5 * bytes));
        FDFD_CUDA(cudaMalloc(&coarse_relax, bytes));
        {
            using R = typename real_of<C>::type;
            const int64_t n = lv.op.n_local();
            const int blocks = (int)((n + 255) / 256);
            if (TM()) coarse_coefs_kernel<C, true><<<blocks, 256>>>(lv.op.Nx, lv.op.Ny, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), (R)o.mg_omega, coarse_k5, coarse_relax, lv.edge());
            else coarse_coefs_kernel<C, false><<<blocks, 256>>>(lv.op.Nx, lv.op.Ny, lv.ca, lv.cb, lv.cd, (R)lv.op.sx, (R)lv.op.sy, shift(), (R)o.mg_omega, coarse_k5, coarse_relax, lv.edge());
            FDFD_CUDA(cudaGetLastError());
        }
        if (sharded) FDFD_TRY(coarse_setup_sharded(lv, 0));
        coarse_ready = true;
        return FDFD_OK;
    }
    int coarse_persistent(Level<C>& lv, const C* b, C* x, cudaStream_t st) {
        using R = typename real_of<C>::type;
        CoarseArgs<C> a;
        a.Nx = lv.op.Nx; a.Ny = lv.op.Ny;
        a.ca = lv.ca; a.cb = lv.cb; a.cd = lv.cd;
        a.sx = (R)lv.op.sx; a.sy = (R)lv.op.sy; a.omega = (R)o.mg_omega;
        a.shift = shift();
        a.its = coarse_its;
        a.b = b; a.x = x;
        a.r = (C*)coarse_ws.v[0]; a.rh = (C*)coarse_ws.v[1]; a.p = (C*)coarse_ws.v[2]; a.v = (C*)coarse_ws.v[3];
        a.t = (C*)coarse_ws.v[4]; a.ph = (C*)coarse_ws.v[5]; a.sh = (C*)coarse_ws.v[6];
        a.phf = coarse_extra[0]; a.shf = coarse_extra[1];
        a.yl = view(lv.ylines, lv.op.Nx, lv.op.Ny);
        a.xl = view(lv.xlines, lv.op.Nx, lv.op.Ny);
        a.k5 = coarse_k5; a.relax = coarse_relax;
        a.partials = coarse_partials;
        a.bar = coarse_bar;
        a.owners = std::max(lv.ylines.nl, lv.xlines.nl);
        a.spikes_in_smem = coarse_spikes ? 1 : 0;
        if (lv.op.Ny != lv.op.Ny_global) a.shard = coarse_sh;
        void* args[] = {&a};
        const void* kern = (const void*)coarse_kernel(lv.op.Ny != lv.op.Ny_global);
        // cooperative launch: the runtime guarantees that all CTAs are co-resident (the barriers need it)
        FDFD_CUDA(cudaLaunchCooperativeKernel(kern, dim3(coarse_grid), dim3(kCoarseThreads), args, coarse_smem_launch, st));
        return FDFD_OK;
    }

    int cycle(int l, cudaStream_t st) {
        using R = typename real_of<C>::type;
        Level<C>& lv = L[l];
        const int last = nlev - 1;
        if (l == last && rep >= 0 && o.mg_coarse_mode >= 1) {
            // Sharded grid: the coarsest problem is latency bound (its kernels do not fill one GPU),
            // so every rank gathers the coarse right-hand side and solves the WHOLE coarse problem
            // redundantly with no communication inside, instead of ~90 latency-bound NCCL calls of a
            // distributed BiCGSTAB; each rank keeps its own rows of the result.
            Level<C>& R = L[rep];
            const size_t loc = (size_t)lv.op.n_local() * sizeof(C);
            if (rep_plan) FDFD_TRY(p2p_allgather(rep_plan, lv.b, R.b, loc, st));
            else FDFD_TRY(comm_allgather(lv.op.comm, lv.b, R.b, loc, st));
            if (coarse_persistent_ok(R)) {
                FDFD_TRY(coarse_persistent(R, R.b, R.r, st));
            } else {
                FDFD_CUDA(cudaMemsetAsync(R.r, 0, (size_t)R.op.n_local() * sizeof(C), st));
                fdfd_krylov_opts io = o;
                io.maxit = coarse_its;
                SolveInfo ii;
                coarse_pre.mg = this; coarse_pre.level = rep;
                HelmOp cop(R.op);
                FDFD_TRY(bicgstab_solve(cop, &coarse_pre, io, R.b, R.r, ii, st, &coarse_ws, true));
            }
            FDFD_CUDA(cudaMemcpyAsync(lv.x, R.r + (int64_t)lv.op.j0 * lv.op.Nx, loc, cudaMemcpyDeviceToDevice, st));
            return FDFD_OK;
        }
        if (l == last) {
            if (o.mg_coarse_mode >= 1) {
                // coarse_its sync-free BiCGSTAB iterations on the shifted coarse operator; the
                // solution lives in lv.r (the smoother ping-pongs lv.x / lv.t underneath)
                const size_t bytes = (size_t)lv.op.n_local() * sizeof(C);
                if (coarse_persistent_ok(lv)) {
                    FDFD_TRY(coarse_persistent(lv, lv.b, lv.r, st));      // starts from a zero guess, as lv.x is here
                } else {
                    FDFD_CUDA(cudaMemcpyAsync(lv.r, lv.x, bytes, cudaMemcpyDeviceToDevice, st));
                    fdfd_krylov_opts io = o;
                    io.maxit = coarse_its;
                    SolveInfo ii;
                    coarse_pre.mg = this; coarse_pre.level = l;
                    HelmOp cop(lv.op);
                    FDFD_TRY(bicgstab_solve(cop, &coarse_pre, io, lv.b, lv.r, ii, st, &coarse_ws, true));
                }
                std::swap(lv.x, lv.r);
            } else {
                for (int s = 0; s < coarse_its; ++s) FDFD_TRY(smooth(l, st));
            }
            return FDFD_OK;
        }
        for (int s = 0; s < o.mg_pre; ++s) FDFD_TRY(smooth(l, st));
        if (!lv.op.halo.state) FDFD_TRY(helm2d_halo_exchange(lv.op, lv.x, st));
        FDFD_TRY(helm2d_launch<C>(lv.op, APPLY_RESID, lv.x, lv.r, lv.b, lv.ca, lv.cb, lv.cd, shift_re, shift_im, 0.0, st));
        Level<C>& lc = L[l + 1];
        dim3 bt(32, 8);
        dim3 gc((lc.op.Nx + 31) / 32, (lc.op.Ny + 7) / 8);
        restrict_kernel<C><<<gc, bt, 0, st>>>(lc.op.Nx, lc.op.Ny, lv.r, lc.b);
        FDFD_CUDA(cudaMemsetAsync(lc.x, 0, (size_t)lc.op.n_local() * sizeof(C), st));
        const int gamma = (o.mg_cycle == 2 || (o.mg_wfrom >= 0 && l + 1 >= o.mg_wfrom)) && (l + 1 < last) ? 2 : 1;
        for (int g = 0; g < gamma; ++g) FDFD_TRY(cycle(l + 1, st));
        const double dc = (lv.delta + 0.5) / 2.0;
        const R wq = (R)(lv.delta / (2.0 * dc));
        dim3 gf((lv.op.Nx + 31) / 32, (lv.op.Ny + 7) / 8);
        // bilinear interpolation reads one coarse row beyond the slab on each open side
        FDFD_TRY(helm2d_halo_exchange(lc.op, lc.x, st));
        const C* gs = lc.first ? nullptr : (const C*)lc.op.halo_south;
        const C* gn = lc.last ? nullptr : (const C*)lc.op.halo_north;
        if (TM()) prolong_add_kernel<C, true><<<gf, bt, 0, st>>>(lc.op.Nx, lc.op.Ny, lc.x, lv.x, wq, gs, gn);
        else prolong_add_kernel<C, false><<<gf, bt, 0, st>>>(lc.op.Nx, lc.op.Ny, lc.x, lv.x, wq, gs, gn);
        FDFD_CUDA(cudaGetLastError());
        for (int s = 0; s < o.mg_post; ++s) FDFD_TRY(smooth(l, st));
        return FDFD_OK;
    }

    // the whole cycle as a stream-ordered sequence (no host synchronisation inside); the result
    // is left in L[0].x
    int run_cycle(const C* v, cudaStream_t st) {
        for (auto& lv : L) { lv.x = lv.x0; lv.r = lv.r0; lv.t = lv.t0; }
        Level<C>& l0 = L[0];
        l0.b = (C*)v;
        FDFD_CUDA(cudaMemsetAsync(l0.x, 0, (size_t)l0.op.n_local() * sizeof(C), st));
        return cycle(0, st);
    }

    // graph mode with the right-hand side already sitting in in_buf (written there by the caller)
    int apply_from_input(void* z, cudaStream_t st) {
        ++applies;
        FDFD_TRY(ensure_graph());
        const size_t bytes = (size_t)L[0].op.n_local() * sizeof(C);
        FDFD_CUDA(cudaEventRecord(ev_in, st));
        FDFD_CUDA(cudaStreamWaitEvent(gstream, ev_in, 0));
        FDFD_CUDA(cudaGraphLaunch(graph, gstream));
        FDFD_CUDA(cudaEventRecord(ev_out, gstream));
        FDFD_CUDA(cudaStreamWaitEvent(st, ev_out, 0));
        FDFD_CUDA(cudaMemcpyAsync(z, out_ptr, bytes, cudaMemcpyDeviceToDevice, st));
        return FDFD_OK;
    }

    int ensure_graph() {
        if (graph) return FDFD_OK;
        // capture once: the cycle reads its right-hand side from in_buf and its kernel
        // sequence / buffer roles are identical for every application
        cudaGraph_t g = nullptr;
        FDFD_CUDA(cudaStreamBeginCapture(gstream, cudaStreamCaptureModeRelaxed));
        int s = run_cycle(in_buf, gstream);
        cudaError_t e = cudaStreamEndCapture(gstream, &g);
        if (s != FDFD_OK) { if (g) cudaGraphDestroy(g); return s; }
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
        e = cudaGraphInstantiate(&graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { graph = nullptr; return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
        out_ptr = L[0].x;
        return FDFD_OK;
    }

    int apply(const void* v, void* z, cudaStream_t st) override {
        ++applies;
        Level<C>& l0 = L[0];
        const size_t bytes = (size_t)l0.op.n_local() * sizeof(C);
        if (!o.mg_graph) {
            FDFD_TRY(run_cycle((const C*)v, st));
            FDFD_CUDA(cudaMemcpyAsync(z, l0.x, bytes, cudaMemcpyDeviceToDevice, st));
            return FDFD_OK;
        }
        FDFD_TRY(ensure_graph());
        FDFD_CUDA(cudaMemcpyAsync(in_buf, v, bytes, cudaMemcpyDeviceToDevice, st));
        FDFD_CUDA(cudaEventRecord(ev_in, st));
        FDFD_CUDA(cudaStreamWaitEvent(gstream, ev_in, 0));
        FDFD_CUDA(cudaGraphLaunch(graph, gstream));
        FDFD_CUDA(cudaEventRecord(ev_out, gstream));
        FDFD_CUDA(cudaStreamWaitEvent(st, ev_out, 0));
        FDFD_CUDA(cudaMemcpyAsync(z, out_ptr, bytes, cudaMemcpyDeviceToDevice, st));
        return FDFD_OK;
    }

    int build(Helm2D& A, const fdfd_krylov_opts& opts, cudaStream_t st) {
        using R = typename real_of<C>::type;
        o = opts;
        shift_re = opts.mg_shift_re; shift_im = opts.mg_shift_im;
        FDFD_TRY(fused_attr());
        const bool sharded = A.Ny != A.Ny_global;
        if (sharded) {
            FDFD_CHECK_ARG(A.comm && A.comm->nranks > 1, "sharded operator needs a communicator");
            FDFD_CHECK_ARG((int64_t)A.Ny * A.comm->nranks == A.Ny_global && A.j0 == A.comm->rank * A.Ny,
                           "sharded multigrid needs equal y-slabs in rank order");
            // the captured cycle then contains the NCCL halo / all-gather / all-reduce kernels; every
            // rank captures and replays the same sequence (measured: 4096^2 on 2 GPUs 1.65 s -> 1.42 s)
        }
        FDFD_CHECK_ARG(A.bcx == FDFD_BC_DIRICHLET && A.bcy == FDFD_BC_DIRICHLET, "multigrid preconditioner needs Dirichlet walls");
        FDFD_CHECK_ARG(A.cd != nullptr, "multigrid preconditioner needs the diagonal term");
        const int min_n = opts.mg_min_n > 2 ? opts.mg_min_n : 2;
        int lx_lo = opts.line_x_lo, lx_hi = opts.line_x_hi, ly_lo = opts.line_y_lo, ly_hi = opts.line_y_hi;
        Level<C> f;
        f.op = A; f.op.h_x = f.op.h_y = nullptr;
        f.ca = (C*)A.ca; f.cb = (C*)A.cb; f.cd = (C*)A.cd; f.own_coefs = false; f.delta = 1.0;
        f.op.shift_re = shift_re; f.op.shift_im = shift_im;
        f.first = A.j0 == 0; f.last = A.j0 + A.Ny == A.Ny_global;
        FDFD_TRY(helm2d_setup_sharding(f.op, st));      // the hierarchy owns its ghost buffers on every level
        const double kh_max = opts.mg_kh_max > 0 ? opts.mg_kh_max : 1e30;
        L.push_back(f);
        while (true) {
            Level<C>& cur = L.back();
            const int Nx = cur.op.Nx, Ny = cur.op.Ny;
            const size_t bytes = (size_t)Nx * Ny * sizeof(C);
            FDFD_CUDA(cudaMalloc(&cur.x, bytes)); FDFD_CUDA(cudaMalloc(&cur.r, bytes)); FDFD_CUDA(cudaMalloc(&cur.t, bytes));
            cur.x0 = cur.x; cur.r0 = cur.r; cur.t0 = cur.t;
            if (L.size() > 1) FDFD_CUDA(cudaMalloc(&cur.b, bytes));
            // evaluated on quantities that are equal on every rank (a rank-dependent failure inside the
            // collective set-up below would leave the other ranks waiting)
            FDFD_CHECK_ARG(!(sharded && (ly_lo >= Ny || ly_hi >= Ny)), "absorbing-layer rows exceed the rows of one slab");
            FDFD_TRY(setup_lines(cur, cur.ylines, 0, lx_lo, lx_hi, st));
            FDFD_TRY(setup_lines(cur, cur.xlines, 1, cur.first ? ly_lo : 0, cur.last ? ly_hi : 0, st));
            const bool max_levels = opts.mg_levels > 0 && (int)L.size() >= opts.mg_levels;
            const int Nyg = cur.op.Ny_global;
            const double s_next = std::min(cur.op.sx, cur.op.sy) / 4.0;
            const double kh_next = s_next > 0 ? 1.0 / std::sqrt(std::fabs(s_next)) : 1e30;
            const bool odd = (Nx & 1) || (Ny & 1) || (cur.op.j0 & 1);
            if (max_levels || Nx / 2 < min_n || Nyg / 2 < min_n || kh_next > kh_max) break;
            if (odd) {
                // (same decision on every rank: equal slabs)
                if (L.size() == 1)
                    fprintf(stderr, "[fdfd] warning: %d x %d rows per slab cannot be halved: the multigrid hierarchy has a "
                                    "single level and degrades to its coarse-level solver (use even slab heights)\n", Nx, Ny);
                break;
            }
            Level<C> c;
            c.op = cur.op;
            c.op.Nx = Nx / 2; c.op.Ny = Ny / 2; c.op.Ny_global = Nyg / 2; c.op.j0 = cur.op.j0 / 2;
            c.first = cur.first; c.last = cur.last;
            c.op.sx = cur.op.sx / 4; c.op.sy = cur.op.sy / 4;
            c.delta = (cur.delta + 0.5) / 2.0;
            const size_t cb = (size_t)c.op.Nx * c.op.Ny * sizeof(C);
            FDFD_CUDA(cudaMalloc(&c.ca, cb)); FDFD_CUDA(cudaMalloc(&c.cb, cb)); FDFD_CUDA(cudaMalloc(&c.cd, cb));
            c.own_coefs = true;
            const R mult = (R)(cur.delta / c.delta);
            dim3 bt(32, 8), g((c.op.Nx + 31) / 32, (c.op.Ny + 7) / 8);
            if (TM()) coarsen_coefs_kernel<C, true><<<g, bt, 0, st>>>(c.op.Nx, c.op.Ny, cur.ca, cur.cb, cur.cd, c.ca, c.cb, c.cd, mult, cur.first, cur.last);
            else coarsen_coefs_kernel<C, false><<<g, bt, 0, st>>>(c.op.Nx, c.op.Ny, cur.ca, cur.cb, cur.cd, c.ca, c.cb, c.cd, mult, cur.first, cur.last);
            FDFD_CUDA(cudaGetLastError());
            c.op.ca = c.ca; c.op.cb = c.cb; c.op.cd = c.cd;
            FDFD_TRY(helm2d_setup_sharding(c.op, st));
            lx_lo = (lx_lo + 1) / 2; lx_hi = (lx_hi + 1) / 2; ly_lo = (ly_lo + 1) / 2; ly_hi = (ly_hi + 1) / 2;
            L.push_back(c);
        }
        nlev = (int)L.size();
        // Replicate the coarsest level only while it is small enough to be latency bound on one GPU
        // (<= 2^20 points, measured: 1024^2 coarse grid of an 8192^2 problem); larger coarse grids are
        // bandwidth bound and keep the distributed BiCGSTAB (16384^2: 2048^2 coarse grid).
        const int64_t coarse_pts = (int64_t)L.back().op.Nx * L.back().op.Ny_global;
        const bool replicate = opts.mg_coarse_replicate > 0 || (opts.mg_coarse_replicate < 0 && coarse_pts <= ((int64_t)1 << 20));
        if (sharded && o.mg_coarse_mode >= 1 && replicate) {
            // replicated copy of the coarsest level: gather its coefficient slabs (equal slabs in rank
            // order -> the gathered arrays are the global arrays) and set it up as a single-rank level
            const Level<C> c = L.back();
            Level<C> R;
            R.op = c.op;
            R.op.Ny = c.op.Ny_global; R.op.j0 = 0; R.op.comm = nullptr;
            R.op.halo_south = R.op.halo_north = R.op.cb_ghost_buf = nullptr;
            R.op.halo = HaloP2P();
            R.first = R.last = true; R.delta = c.delta;
            const size_t loc = (size_t)c.op.n_local() * sizeof(C), full = (size_t)R.op.n_local() * sizeof(C);
            FDFD_CUDA(cudaMalloc(&R.ca, full)); FDFD_CUDA(cudaMalloc(&R.cb, full)); FDFD_CUDA(cudaMalloc(&R.cd, full));
            R.own_coefs = true;
            FDFD_TRY(comm_allgather(c.op.comm, c.ca, R.ca, loc, st));
            FDFD_TRY(comm_allgather(c.op.comm, c.cb, R.cb, loc, st));
            FDFD_TRY(comm_allgather(c.op.comm, c.cd, R.cd, loc, st));
            R.op.ca = R.ca; R.op.cb = R.cb; R.op.cd = R.cd;
            FDFD_CUDA(cudaMalloc(&R.x, full)); FDFD_CUDA(cudaMalloc(&R.r, full)); FDFD_CUDA(cudaMalloc(&R.t, full));
            FDFD_CUDA(cudaMalloc(&R.b, full));
            R.x0 = R.x; R.r0 = R.r; R.t0 = R.t;
            // strip widths of this level (halved once per coarsening, as in the loop above)
            int wx_lo = opts.line_x_lo, wx_hi = opts.line_x_hi, wy_lo = opts.line_y_lo, wy_hi = opts.line_y_hi;
            for (int q = 1; q < nlev; ++q) { wx_lo = (wx_lo + 1) / 2; wx_hi = (wx_hi + 1) / 2; wy_lo = (wy_lo + 1) / 2; wy_hi = (wy_hi + 1) / 2; }
            L.push_back(R);
            rep = (int)L.size() - 1;
            if (p2p_enabled(c.op.comm)) {
                rep_comm = c.op.comm;
                FDFD_TRY(p2p_gather_create(rep_comm, loc, &rep_plan, st));
            }
            FDFD_TRY(setup_lines(L[rep], L[rep].ylines, 0, wx_lo, wx_hi, st));
            FDFD_TRY(setup_lines(L[rep], L[rep].xlines, 1, wy_lo, wy_hi, st));
        }
        {
            const Level<C>& c = L[nlev - 1];
            const double kh = 1.0 / std::sqrt(std::fabs(std::min(c.op.sx, c.op.sy)));
            if (opts.mg_coarse_its > 0) coarse_its = opts.mg_coarse_its;
            else {
                coarse_its = (int)std::ceil(10.0 / kh);
                if (coarse_its < 8) coarse_its = 8;
                if (coarse_its > 80) coarse_its = 80;
            }
            if (o.verbose)
                fprintf(stderr, "[fdfd] multigrid: %d levels, coarsest %d x %d (k*h = %.3f), coarse mode %d its %d\n",
                        nlev, c.op.Nx, c.op.Ny_global, kh, o.mg_coarse_mode, coarse_its);
        }
        if (o.mg_coarse_mode >= 1) {
            // the inner solver must not allocate inside the (captured) cycle
            const Level<C>& c = L.back();      // the replicated level when present (the larger one)
            FDFD_TRY(coarse_ws.ensure((size_t)c.op.n_local() * sizeof(C), true));
            FDFD_CUDA(cudaMalloc(&zero_buf, (size_t)c.op.n_local() * sizeof(C)));
            FDFD_CUDA(cudaMemsetAsync(zero_buf, 0, (size_t)c.op.n_local() * sizeof(C), st));
            if (coarse_persistent_shape_ok(c)) FDFD_TRY(coarse_setup(c));
        }
        if (o.mg_graph) {
            FDFD_CUDA(cudaMalloc(&in_buf, (size_t)L[0].op.n_local() * sizeof(C)));
            FDFD_CUDA(cudaStreamCreateWithFlags(&gstream, cudaStreamNonBlocking));
            FDFD_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
            FDFD_CUDA(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
        }
        FDFD_CUDA(cudaStreamSynchronize(st));
        return FDFD_OK;
    }
};

// Mixed precision: the cycle runs in complex64 (half the bytes of every smoothing / transfer pass)
// inside a complex128 flexible Krylov method, which tolerates the inexact preconditioner.
struct MixedMGPrecond : Precond {
    Helm2D A32;
    float2 *ca = nullptr, *cb = nullptr, *cd = nullptr, *v32 = nullptr, *z32 = nullptr;
    MGPrecond<float2>* mg = nullptr;
    int64_t n = 0;
    ~MixedMGPrecond() override {
        delete mg;
        cudaFree(ca); cudaFree(cb); cudaFree(cd); cudaFree(v32); cudaFree(z32);
    }
    int build(Helm2D& A, const fdfd_krylov_opts& opts, cudaStream_t st) {
        n = A.n_local();
        const size_t bytes = (size_t)n * sizeof(float2);
        FDFD_CUDA(cudaMalloc(&ca, bytes)); FDFD_CUDA(cudaMalloc(&cb, bytes)); FDFD_CUDA(cudaMalloc(&cd, bytes));
        FDFD_CUDA(cudaMalloc(&v32, bytes)); FDFD_CUDA(cudaMalloc(&z32, bytes));
        FDFD_TRY((convert<double2, float2>(n, (const double2*)A.ca, ca, st)));
        FDFD_TRY((convert<double2, float2>(n, (const double2*)A.cb, cb, st)));
        FDFD_TRY((convert<double2, float2>(n, (const double2*)A.cd, cd, st)));
        A32 = A;
        A32.dtype = FDFD_C64;
        A32.ca = ca; A32.cb = cb; A32.cd = cd;
        A32.halo_south = A32.halo_north = A32.cb_ghost_buf = nullptr;
        A32.halo = HaloP2P();
        A32.h_x = A32.h_y = nullptr;
        mg = new MGPrecond<float2>();
        return mg->build(A32, opts, st);
    }
    int apply(const void* v, void* z, cudaStream_t st) override {
        ++applies;
        FDFD_TRY((convert<double2, float2>(n, (const double2*)v, v32, st)));
        FDFD_TRY(mg->apply(v32, z32, st));
        return convert<float2, double2>(n, z32, (double2*)z, st);
    }
    // complex64 in / out: no conversion passes (FGMRES keeps basis and corrections in complex64)
    bool single_io() const override { return true; }
    float2* single_input() override { return mg->o.mg_graph ? mg->in_buf : nullptr; }
    int apply_single(const float2* v, float2* z, cudaStream_t st) override {
        ++applies;
        if (!v) return mg->apply_from_input(z, st);
        return mg->apply(v, z, st);
    }
};

}  // namespace

int make_mg_precond(Helm2D& A, const fdfd_krylov_opts& opts, Precond** out, cudaStream_t stream) {
    int s;
    if (A.dtype == FDFD_C128 && opts.mg_single) {
        FDFD_CHECK_ARG(A.cd != nullptr, "multigrid preconditioner needs the diagonal term");
        auto* M = new MixedMGPrecond();
        s = M->build(A, opts, stream);
        if (s != FDFD_OK) { delete M; return s; }
        *out = M;
        return FDFD_OK;
    }
    if (A.dtype == FDFD_C128) {
        auto* M = new MGPrecond<double2>();
        s = M->build(A, opts, stream);
        if (s != FDFD_OK) { delete M; return s; }
        *out = M;
    } else {
        auto* M = new MGPrecond<float2>();
        s = M->build(A, opts, stream);
        if (s != FDFD_OK) { delete M; return s; }
        *out = M;
    }
    return FDFD_OK;
}

}  // namespace fdfd
