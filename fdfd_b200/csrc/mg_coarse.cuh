// Persistent coarsest-level solver of the multigrid cycle (mg_coarse_mode = 2, opt-in).
//
// Included by mg.cu inside its anonymous namespace (uses PointCoefs / point_coefs / LineView).
//
// The coarsest level (512^2 for a 4096^2 grid) is solved by a fixed number of BiCGSTAB iterations
// preconditioned with one smoothing sweep (mode 1: krylov.cu bicgstab_t with fixed_its + MGPrecond::
// smooth).  That is ~20 dependent kernel launches per iteration on a problem that does not fill the
// GPU: ~1.6 ms of the 3.6 ms an outer iteration takes at 4096^2.  Here the same arithmetic runs in ONE
// cooperative kernel: all CTAs stay resident, every vector (2 MB in complex64) stays in L2, phases are
// separated by grid-wide barriers and inner products are reduced through a per-CTA partials array that
// every CTA sums redundantly after the barrier (deterministic, no atomics).
//
// Per iteration (x0 = 0, so the initial residual is b):
//   A   p = r + beta (p - omega v);  p^ = omega_s D^-1 p  (strip columns: line right-hand side)   | pointwise
//   L1  y-lines of p^                                                                              | 1 CTA / line
//   L2  x-lines of p^: row residual, solve                                                         | 1 CTA / line
//   L3  x-lines of p^: update                                                                      | 1 CTA / line
//   V   v = M p^ ;  (r^, v)                                                                        | stencil
//   S   s = r - alpha v ;  s^ = omega_s D^-1 s (as A)                                              | pointwise
//   L1..L3 for s^
//   T   t = M s^ ;  (t, s), (t, t)                                                                 | stencil
//   X   x += alpha p^ + omega s^ ;  r = s - omega t ;  (r^, r)                                     | pointwise
// 11 grid barriers per iteration.
//
// STATUS: written at the end of round 1 without GPU time left — compiled, never run.  Not reachable
// unless mg_coarse_mode == 2 is requested explicitly.
#pragma once

// <cooperative_groups.h> is included by mg.cu at global scope (this file sits inside a namespace)
namespace cg = ::cooperative_groups;

constexpr int kCoarseThreads = 256;
constexpr int kCoarseDots = 4;          // partial sums per CTA: 4 complex numbers

template <typename C>
struct CoarseArgs {
    int Nx, Ny;
    const C *ca, *cb, *cd;
    typename real_of<C>::type sx, sy, omega;
    C shift;
    int its;
    const C* b;                          // right-hand side
    C* x;                                // solution (overwritten; initial guess is zero)
    C *r, *rh, *p, *v, *t, *ph, *sh;     // work vectors
    LineView<C> yl, xl;                  // line sets of the level (nl may be 0)
    double* partials;                    // [gridDim.x][2 * kCoarseDots]
};

__device__ __forceinline__ double2 cs_div(double2 a, double2 b) {     // as blas1.cu safe_div
    const double d = b.x * b.x + b.y * b.y;
    if (!(d > 0.0) || !isfinite(d)) return make_double2(0.0, 0.0);
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ double2 cs_mul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// M x at one point (five-point stencil of the shifted operator)
template <typename C>
__device__ __forceinline__ C coarse_stencil(const PointCoefs<C>& pc, const C* x, int Nx, int Ny, int i, int j, int64_t n) {
    C ax = cmul(pc.dg, x[n]);
    if (i > 0) ax = cfma(pc.kw, x[n - 1], ax);
    if (i < Nx - 1) ax = cfma(pc.ke, x[n + 1], ax);
    if (j > 0) ax = cfma(pc.ks, x[n - Nx], ax);
    if (j < Ny - 1) ax = cfma(pc.kn, x[n + Nx], ax);
    return ax;
}

// block-wide sum of K doubles -> partials row of this CTA
template <int K>
__device__ __forceinline__ void coarse_block_reduce(double (&v)[K], double* row, double* sm /* [K][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
        if (lane == 0) sm[k * 8 + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < kCoarseThreads / 32; ++w) s += sm[threadIdx.x * 8 + w];
        row[threadIdx.x] = s;
    }
    __syncthreads();
}

// after the grid barrier: every CTA sums the K partials of all CTAs (same order everywhere)
template <int K>
__device__ __forceinline__ void coarse_grid_total(const double* partials, int stride, int offset, double (&out)[K], double* sm) {
    if (threadIdx.x < 32) {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int bidx = threadIdx.x; bidx < (int)gridDim.x; bidx += 32) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)bidx * stride + offset + k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], off);
            if (threadIdx.x == 0) sm[k] = acc[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = sm[k];
    __syncthreads();
}

// One relaxation line handled by the calling CTA (copy of line_fused_kernel's body; DIR 0: the update is
// applied to column `fixed` of x; DIR 1: the correction is left in v.g for coarse_xline_apply).
template <typename C, bool TM, int DIR>
__device__ void coarse_line(const LineView<C>& v, int l, C* x, const C* b, const C* ca, const C* cb, const C* cd,
                            typename real_of<C>::type sx, typename real_of<C>::type sy, C shift,
                            typename real_of<C>::type omega, unsigned char* smem, C* sg, C* su) {
    const int64_t nl = v.nl;
    const int len = v.len;
    C* sl = reinterpret_cast<C*>(smem);
    C* s_fl = sl + len; C* s_fc = s_fl + len; C* s_ip = s_fc + len;
    const int fixed = l < v.lo ? l : (DIR == 0 ? v.Nx : v.Ny) - v.hi + (l - v.lo);
    const int n = 2 * v.Ptot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kCoarseThreads / 32;
    for (int t = threadIdx.x; t < len; t += kCoarseThreads) {
        s_fl[t] = v.fl[t * nl + l]; s_fc[t] = v.fc[t * nl + l]; s_ip[t] = v.fip[t * nl + l];
    }
    if (DIR == 0) {
        for (int t = threadIdx.x; t < len; t += kCoarseThreads) sl[t] = v.g[t * nl + l];
    } else {
        const int j = fixed;
        for (int i = threadIdx.x; i < len; i += kCoarseThreads) {
            PointCoefs<C> pc = point_coefs<C, TM>(ca, cb, cd, v.Nx, v.Ny, sx, sy, shift, i, j);
            const int64_t pn = (int64_t)j * v.Nx + i;
            sl[i] = csub(b[pn], coarse_stencil<C>(pc, x, v.Nx, v.Ny, i, j, pn));
        }
    }
    __syncthreads();
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
        for (int row = warp; row < n; row += nw) {
            const C* Ri = v.rinv + ((int64_t)l * n + row) * n;
            double ax = 0.0, ay = 0.0;
            for (int q = lane; q < n; q += 32) {
                const C rv = Ri[q], g = sg[q];
                ax += (double)rv.x * g.x - (double)rv.y * g.y;
                ay += (double)rv.x * g.y + (double)rv.y * g.x;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ax += __shfl_down_sync(0xffffffffu, ax, off);
                ay += __shfl_down_sync(0xffffffffu, ay, off);
            }
            if (lane == 0) su[row] = mk<C>((typename real_of<C>::type)ax, (typename real_of<C>::type)ay);
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < len; t += kCoarseThreads) {
        C e = sl[t];
        if (v.Ptot > 1) {
            const int kg = t / v.m;
            if (kg > 0) e = csub(e, cmul(v.sv[t * nl + l], su[2 * (kg - 1) + 1]));
            if (kg < v.Ptot - 1) e = csub(e, cmul(v.sw[t * nl + l], su[2 * (kg + 1)]));
        }
        if (DIR == 0) {
            const int64_t pn = (int64_t)t * v.Nx + fixed;
            x[pn] = cadd(x[pn], cscale(omega, e));
        } else {
            v.g[t * nl + l] = e;
        }
    }
    __syncthreads();
}

// the three line phases of one smoothing sweep from a zero guess: z holds omega D^-1 rhs outside the
// strip columns and 0 inside, yl.g holds the strip-column right-hand sides
template <typename C, bool TM>
__device__ void coarse_line_phases(const CoarseArgs<C>& a, cg::grid_group& grid, C* z, const C* rhs,
                                   unsigned char* smem, C* sg, C* su) {
    const int l = blockIdx.x;
    if (a.yl.nl > 0) {
        if (l < a.yl.nl) coarse_line<C, TM, 0>(a.yl, l, z, rhs, a.ca, a.cb, a.cd, a.sx, a.sy, a.shift, a.omega, smem, sg, su);
        grid.sync();
    }
    if (a.xl.nl > 0) {
        if (l < a.xl.nl) coarse_line<C, TM, 1>(a.xl, l, z, rhs, a.ca, a.cb, a.cd, a.sx, a.sy, a.shift, a.omega, smem, sg, su);
        grid.sync();
        if (l < a.xl.nl) {
            const int j = l < a.xl.lo ? l : a.Ny - a.xl.hi + (l - a.xl.lo);
            for (int i = threadIdx.x; i < a.xl.len; i += kCoarseThreads) {
                const int64_t pn = (int64_t)j * a.Nx + i;
                z[pn] = cadd(z[pn], cscale(a.omega, a.xl.g[(int64_t)i * a.xl.nl + l]));
            }
        }
        grid.sync();
    }
}

// pointwise part of the zero-guess smoothing sweep for the value `val` at point (i, j)
template <typename C, bool TM>
__device__ __forceinline__ C coarse_point_relax(const CoarseArgs<C>& a, int i, int j, int64_t n, C val) {
    const LineView<C>& yl = a.yl;
    if (yl.nl > 0 && (i < yl.lo || i >= a.Nx - yl.hi)) {
        const int l = i < yl.lo ? i : yl.lo + (i - (a.Nx - yl.hi));
        yl.g[(int64_t)j * yl.nl + l] = val;               // right-hand side of the y-line through this column
        return czero<C>();
    }
    const PointCoefs<C> pc = point_coefs<C, TM>(a.ca, a.cb, a.cd, a.Nx, a.Ny, a.sx, a.sy, a.shift, i, j);
    return cscale(a.omega, cdiv(val, pc.dg));
}

template <typename C, bool TM>
__global__ void __launch_bounds__(kCoarseThreads) coarse_bicgstab_kernel(CoarseArgs<C> a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char coarse_smem[];
    __shared__ C sg[2 * kMaxChunks], su[2 * kMaxChunks];
    __shared__ double red[2 * kCoarseDots * 8];
    const int n = a.Nx * a.Ny;                       // < 2^31 (checked on the host)
    const int tid = blockIdx.x * kCoarseThreads + threadIdx.x;
    const int stride = gridDim.x * kCoarseThreads;
    const int pstride = 2 * kCoarseDots;
    double* prow = a.partials + (size_t)blockIdx.x * pstride;

    // restart: x = 0, r = r^ = b, p = v = 0, rho = (r, r), rho_old = alpha = omega = 1
    double acc1[2] = {0.0, 0.0};
    for (int q = tid; q < n; q += stride) {
        const C bv = a.b[q];
        a.r[q] = bv; a.rh[q] = bv; a.p[q] = czero<C>(); a.v[q] = czero<C>(); a.x[q] = czero<C>();
        acc1[0] += (double)bv.x * bv.x + (double)bv.y * bv.y;
    }
    coarse_block_reduce<2>(acc1, prow, red);
    grid.sync();
    double tot2[2];
    coarse_grid_total<2>(a.partials, pstride, 0, tot2, red);
    double2 rho = make_double2(tot2[0], 0.0), rho_old = make_double2(1.0, 0.0);
    double2 alpha = make_double2(1.0, 0.0), omega = make_double2(1.0, 0.0);

    for (int it = 0; it < a.its; ++it) {
        // A: p = r + beta (p - omega v), p^ pointwise
        const double2 beta = cs_mul(cs_div(rho, rho_old), cs_div(alpha, omega));
        const double2 nbo = make_double2(-(beta.x * omega.x - beta.y * omega.y), -(beta.x * omega.y + beta.y * omega.x));
        const C betaC = mk<C>((typename real_of<C>::type)beta.x, (typename real_of<C>::type)beta.y);
        const C nboC = mk<C>((typename real_of<C>::type)nbo.x, (typename real_of<C>::type)nbo.y);
        for (int q = tid; q < n; q += stride) {
            const int j = q / a.Nx, i = q - j * a.Nx;
            C pv = cfma(nboC, a.v[q], cfma(betaC, a.p[q], a.r[q]));
            a.p[q] = pv;
            a.ph[q] = coarse_point_relax<C, TM>(a, i, j, q, pv);
        }
        grid.sync();
        coarse_line_phases<C, TM>(a, grid, a.ph, a.p, coarse_smem, sg, su);
        // V: v = M p^, (r^, v)
        double accv[2] = {0.0, 0.0};
        for (int q = tid; q < n; q += stride) {
            const int j = q / a.Nx, i = q - j * a.Nx;
            const PointCoefs<C> pc = point_coefs<C, TM>(a.ca, a.cb, a.cd, a.Nx, a.Ny, a.sx, a.sy, a.shift, i, j);
            const C vv = coarse_stencil<C>(pc, a.ph, a.Nx, a.Ny, i, j, q);
            a.v[q] = vv;
            const C rhv = a.rh[q];
            accv[0] += (double)rhv.x * vv.x + (double)rhv.y * vv.y;      // conj(r^) * v
            accv[1] += (double)rhv.x * vv.y - (double)rhv.y * vv.x;
        }
        coarse_block_reduce<2>(accv, prow + 2, red);
        grid.sync();
        coarse_grid_total<2>(a.partials, pstride, 2, tot2, red);
        alpha = cs_div(rho, make_double2(tot2[0], tot2[1]));
        // S: s = r - alpha v (in r), s^ pointwise
        const C alphaC = mk<C>((typename real_of<C>::type)alpha.x, (typename real_of<C>::type)alpha.y);
        for (int q = tid; q < n; q += stride) {
            const int j = q / a.Nx, i = q - j * a.Nx;
            const C sv = csub(a.r[q], cmul(alphaC, a.v[q]));
            a.r[q] = sv;
            a.sh[q] = coarse_point_relax<C, TM>(a, i, j, q, sv);
        }
        grid.sync();
        coarse_line_phases<C, TM>(a, grid, a.sh, a.r, coarse_smem, sg, su);
        // T: t = M s^, (t, s), (t, t)
        double acct[4] = {0.0, 0.0, 0.0, 0.0};
        for (int q = tid; q < n; q += stride) {
            const int j = q / a.Nx, i = q - j * a.Nx;
            const PointCoefs<C> pc = point_coefs<C, TM>(a.ca, a.cb, a.cd, a.Nx, a.Ny, a.sx, a.sy, a.shift, i, j);
            const C tv = coarse_stencil<C>(pc, a.sh, a.Nx, a.Ny, i, j, q);
            a.t[q] = tv;
            const C sv = a.r[q];
            acct[0] += (double)tv.x * sv.x + (double)tv.y * sv.y;        // conj(t) * s
            acct[1] += (double)tv.x * sv.y - (double)tv.y * sv.x;
            acct[2] += (double)tv.x * tv.x + (double)tv.y * tv.y;
        }
        coarse_block_reduce<4>(acct, prow + 4, red);
        grid.sync();
        double tot4[4];
        coarse_grid_total<4>(a.partials, pstride, 4, tot4, red);
        omega = cs_div(make_double2(tot4[0], tot4[1]), make_double2(tot4[2], 0.0));
        // X: x += alpha p^ + omega s^, r = s - omega t, rho = (r^, r)
        const C omegaC = mk<C>((typename real_of<C>::type)omega.x, (typename real_of<C>::type)omega.y);
        double accr[2] = {0.0, 0.0};
        for (int q = tid; q < n; q += stride) {
            a.x[q] = cfma(omegaC, a.sh[q], cfma(alphaC, a.ph[q], a.x[q]));
            const C rv = csub(a.r[q], cmul(omegaC, a.t[q]));
            a.r[q] = rv;
            const C rhv = a.rh[q];
            accr[0] += (double)rhv.x * rv.x + (double)rhv.y * rv.y;
            accr[1] += (double)rhv.x * rv.y - (double)rhv.y * rv.x;
        }
        coarse_block_reduce<2>(accr, prow, red);
        grid.sync();
        coarse_grid_total<2>(a.partials, pstride, 0, tot2, red);
        rho_old = rho;
        rho = make_double2(tot2[0], tot2[1]);
    }
}
