// K6b — batched shift-invert Arnoldi for the Bloch sweep of a photonic-crystal unit cell.
//
// Reference: Band_Diagram_Solver/Band_Diagram_Solver.py:302-323 — for every Bloch vector of the path the
// operators are rebuilt (yeeder2d with BC = (1, 1)), A = -D^H c D is formed with sparse products and
// `eigs(A, M, k, sigma=0)` runs ARPACK mode 3 with a SuperLU factorisation of A - sigma M.
//
// Here ONE CTA owns one path point for the whole eigen-solve and everything is this library's code:
//   * S = A - sigma M of a cell with Bloch walls is cyclic block tridiagonal in the natural ordering
//     n = i + Nx j: Ny block rows of size Nx, cyclic-tridiagonal diagonal blocks T_j (x couplings and the
//     x wrap) and DIAGONAL off-diagonal blocks L_j, U_j (y couplings; the y wrap puts L_0 and U_{Ny-1} in
//     the corners).  band_factor_kernel runs the bordered block LU of that structure in shared memory
//     (last block row / column as the border, Schur complements D_j of Nx x Nx inverted by Gauss-Jordan):
//     4 Nx^3 complex multiply-adds per block row instead of the O(n^3) dense inverse the first version of
//     this path obtained from a library, and no dense n x n matrix is ever assembled.  With the shift left
//     of the spectrum S is Hermitian positive definite, so every Schur complement is and no pivoting is
//     needed.
//   * band_cycle_kernel advances the Arnoldi recurrence of (A - sigma M)^-1 M: block forward / backward
//     substitution with the stored D_j^-1, W_j = D_j^-1 F_j, G_j (dense Nx x Nx mat-vecs, one warp per row),
//     then classical Gram-Schmidt with a second pass against the basis, all steps of a restart cycle in one
//     launch (the points are independent: no grid-wide dependency).
//   * band_combine_kernel forms V <- Q V (Krylov-Schur restart) and the Ritz vectors, band_residual_kernel
//     the true residuals ||A x - lambda M x|| with the matrix-free Bloch stencil.
// Only the (ncv + 1) x ncv projected matrices visit the host (dense eigen / Schur decompositions of order ~20).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace fdfd {

namespace {

using Z = double2;

struct BandGeom {
    int kind;                 // FDFD_POL_TE / FDFD_POL_TM operator type (see helm2d.cu)
    int Nx, Ny;
    double sx, sy, sigma;
    const Z *ca, *cb, *m;     // (Ny, Nx) device arrays: face coefficients and the diagonal of M
};

struct Entries { Z kw, ke, ks, kn, dg; };

// matrix entries of S = sx D ca D + sy D cb D - sigma M at (i, j), Bloch wrap in both directions, exactly as
// helm2d_march_kernel applies them (wrapped faces: neighbour value times the phase, centre weight |ph|^2)
__device__ __forceinline__ Entries band_entries(const BandGeom& g, Z phx, Z phy, int i, int j) {
    const int Nx = g.Nx, Ny = g.Ny;
    const bool TM = g.kind == FDFD_POL_TM;
    const int iw = i == 0 ? Nx - 1 : i - 1, ie = i == Nx - 1 ? 0 : i + 1;
    const int js = j == 0 ? Ny - 1 : j - 1, jn = j == Ny - 1 ? 0 : j + 1;
    const double phx2 = phx.x * phx.x + phx.y * phx.y, phy2 = phy.x * phy.x + phy.y * phy.y;
    const int64_t row = (int64_t)j * Nx;
    Z ca_w, ca_e, c_s, c_n;
    double w_c2 = 1, e_c2 = 1, s_c2 = 1, n_c2 = 1;
    if (TM) {
        ca_w = g.ca[row + i]; ca_e = g.ca[row + ie];
        c_s = g.cb[row + i]; c_n = g.cb[(int64_t)jn * Nx + i];
        if (i == Nx - 1) e_c2 = phx2;
        if (j == Ny - 1) n_c2 = phy2;
    } else {
        ca_w = g.ca[row + iw]; ca_e = g.ca[row + i];
        c_s = g.cb[(int64_t)js * Nx + i]; c_n = g.cb[row + i];
        if (i == 0) w_c2 = phx2;
        if (j == 0) s_c2 = phy2;
    }
    Entries e;
    e.ke = cscale(g.sx, i == Nx - 1 ? cmul(ca_e, phx) : ca_e);
    e.kw = cscale(g.sx, i == 0 ? cmulc(phx, ca_w) : ca_w);         // conj(phx) * ca
    e.kn = cscale(g.sy, j == Ny - 1 ? cmul(c_n, phy) : c_n);
    e.ks = cscale(g.sy, j == 0 ? cmulc(phy, c_s) : c_s);
    Z dg = cadd(cscale(-g.sx, cadd(cscale(e_c2, ca_e), cscale(w_c2, ca_w))),
                cscale(-g.sy, cadd(cscale(n_c2, c_n), cscale(s_c2, c_s))));
    const Z mv = g.m[row + i];
    e.dg = mk<Z>(dg.x - g.sigma * mv.x, dg.y - g.sigma * mv.y);
    return e;
}

// ---- dense Nx x Nx helpers in shared memory (row-major, whole CTA) ---------------------------------------
// C = A * B
__device__ void sm_gemm(int N, const Z* A, const Z* B, Z* C) {
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int a = e / N, b = e - a * N;
        Z acc = czero<Z>();
        for (int c = 0; c < N; ++c) acc = cfma(A[a * N + c], B[c * N + b], acc);
        C[e] = acc;
    }
    __syncthreads();
}

// in-place inverse by Gauss-Jordan without pivoting (Hermitian positive definite Schur complements)
__device__ void sm_invert(int N, Z* A, Z* colk) {
    for (int k = 0; k < N; ++k) {
        const Z p = cdiv(mk<Z>(1, 0), A[k * N + k]);
        __syncthreads();
        for (int r = threadIdx.x; r < N; r += blockDim.x) colk[r] = A[r * N + k];
        __syncthreads();
        for (int c = threadIdx.x; c < N; c += blockDim.x) A[k * N + c] = c == k ? p : cmul(A[k * N + c], p);
        __syncthreads();
        for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
            const int r = e / N, c = e - r * N;
            if (r == k) continue;
            const Z f = colk[r];
            A[e] = c == k ? cneg(cmul(f, p)) : csub(A[e], cmul(f, A[k * N + c]));
        }
        __syncthreads();
    }
}

__device__ void sm_store(int N, const Z* A, Z* out) {
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) out[e] = A[e];
}

// Bordered block LU of the cyclic block-tridiagonal S of one path point (see the header comment):
//   D_0 = T_0, F_0 = L_0 (corner), G_0 = U_m (corner), m = Ny - 1
//   D_j = T_j - L_j D_{j-1}^-1 U_{j-1},  F_j = -L_j W_{j-1} (+ U_{m-1} for j = m-1),  W_j = D_j^-1 F_j
//   G_j = -G_{j-1} D_{j-1}^-1 U_{j-1} (+ L_m for j = m-1),   D_m = T_m - sum_j G_j W_j
// stored: Dinv[j] (j <= m), W[j], G[j] (j < m), each Nx x Nx row-major.
__global__ void __launch_bounds__(512) band_factor_kernel(BandGeom g, const Z* __restrict__ phases, Z* Dinv, Z* W, Z* G) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    const int N = g.Nx, m = g.Ny - 1, NN = N * N;
    const int p = blockIdx.x;
    const Z phx = phases[2 * p], phy = phases[2 * p + 1];
    Z* E = reinterpret_cast<Z*>(band_smem);      // D_{j-1}^-1
    Z* Dw = E + NN;                              // work: D_j
    Z* Wp = Dw + NN;                             // W_{j-1}
    Z* Gp = Wp + NN;                             // G_{j-1}
    Z* T = Gp + NN;                              // temporary
    Z* colk = T + NN;                            // N
    Z* dl = colk + N;                            // L_j (N), U_{j-1} (N), corner diagonals
    Z* du = dl + N;
    Z* Dp = Dinv + (size_t)p * (m + 1) * NN;
    Z* Wo = W + (size_t)p * m * NN;
    Z* Go = G + (size_t)p * m * NN;
    Z* acc = Dp + (size_t)m * NN;                // D_m accumulates in its own output slot

    auto build_T = [&](int j, Z* out) {
        for (int e = threadIdx.x; e < NN; e += blockDim.x) out[e] = czero<Z>();
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const Entries en = band_entries(g, phx, phy, i, j);
            const int iw = i == 0 ? N - 1 : i - 1, ie = i == N - 1 ? 0 : i + 1;
            // (+= : for N = 2 both x neighbours are the same unknown)
            out[i * N + i] = cadd(out[i * N + i], en.dg);
            out[i * N + iw] = cadd(out[i * N + iw], en.kw);
            out[i * N + ie] = cadd(out[i * N + ie], en.ke);
        }
        __syncthreads();
    };

    // ---- j = 0 ----
    build_T(0, E);
    sm_invert(N, E, colk);
    sm_store(N, E, Dp);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        dl[i] = band_entries(g, phx, phy, i, 0).ks;      // L_0: coupling of row 0 to x_m (corner)
        du[i] = band_entries(g, phx, phy, i, m).kn;      // U_m: coupling of row m to x_0 (corner)
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
        const int a = e / N, b = e - a * N;
        Wp[e] = cmul(E[e], dl[b]);                       // W_0 = D_0^-1 diag(L_0)
        Gp[e] = a == b ? du[a] : czero<Z>();             // G_0 = diag(U_m)
    }
    __syncthreads();
    if (m == 1) {                                        // two block rows: both corner couplings hit the same pair
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const int a = e / N, b = e - a * N;
            const Z u0 = band_entries(g, phx, phy, b, 0).kn, l1 = a == b ? band_entries(g, phx, phy, a, 1).ks : czero<Z>();
            Wp[e] = cadd(Wp[e], cmul(E[e], u0));
            Gp[e] = cadd(Gp[e], l1);
        }
        __syncthreads();
    }
    sm_store(N, Wp, Wo);
    sm_store(N, Gp, Go);
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {      // acc = G_0 W_0 (G_0 diagonal unless m == 1)
        const int a = e / N;
        acc[e] = m == 1 ? czero<Z>() : cmul(du[a], Wp[e]);
    }
    __syncthreads();
    if (m == 1) {
        sm_gemm(N, Gp, Wp, T);
        sm_store(N, T, acc);
        __syncthreads();
    }
    // ---- j = 1 .. m-1 ----
    for (int j = 1; j < m; ++j) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            dl[i] = band_entries(g, phx, phy, i, j).ks;          // L_j
            du[i] = band_entries(g, phx, phy, i, j - 1).kn;      // U_{j-1}
        }
        __syncthreads();
        sm_gemm(N, Gp, E, T);                                    // T = G_{j-1} D_{j-1}^-1
        build_T(j, Dw);
        const bool lastj = j == m - 1;
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const int a = e / N, b = e - a * N;
            Z gv = cneg(cmul(T[e], du[b]));                      // G_j = -T diag(U_{j-1})
            if (lastj && a == b) gv = cadd(gv, band_entries(g, phx, phy, a, m).ks);      // + L_m
            Gp[e] = gv;
            Dw[e] = csub(Dw[e], cmul(cmul(dl[a], E[e]), du[b])); // D_j = T_j - L_j E U_{j-1}
        }
        __syncthreads();
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const int a = e / N, b = e - a * N;
            Z fv = cneg(cmul(dl[a], Wp[e]));                     // F_j = -L_j W_{j-1}
            if (lastj && a == b) fv = cadd(fv, band_entries(g, phx, phy, a, m - 1).kn);  // + U_{m-1}
            T[e] = fv;
        }
        __syncthreads();
        sm_invert(N, Dw, colk);                                  // Dw = D_j^-1
        sm_gemm(N, Dw, T, Wp);                                   // W_j = D_j^-1 F_j
        sm_gemm(N, Gp, Wp, T);                                   // G_j W_j
        for (int e = threadIdx.x; e < NN; e += blockDim.x) acc[e] = cadd(acc[e], T[e]);
        sm_store(N, Dw, Dp + (size_t)j * NN);
        sm_store(N, Wp, Wo + (size_t)j * NN);
        sm_store(N, Gp, Go + (size_t)j * NN);
        __syncthreads();
        Z* sw = E; E = Dw; Dw = sw;
    }
    // ---- border ----
    build_T(m, Dw);
    for (int e = threadIdx.x; e < NN; e += blockDim.x) Dw[e] = csub(Dw[e], acc[e]);
    __syncthreads();
    sm_invert(N, Dw, colk);
    sm_store(N, Dw, acc);
}

// y[a] (+)= sign * sum_b Mtx[a][b] * scale[b] * u[b]  for a < N; one warp per row (Mtx in global memory)
template <bool SCALE>
__device__ __forceinline__ void warp_rows_matvec(int N, const Z* __restrict__ Mtx, const Z* u, const Z* scale,
                                                 Z* y, double sign, bool accumulate) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int a = warp; a < N; a += nw) {
        double sx = 0.0, sy = 0.0;
        for (int b = lane; b < N; b += 32) {
            const Z mv = Mtx[(size_t)a * N + b];
            Z uv = u[b];
            if (SCALE) uv = cmul(scale[b], uv);
            sx += mv.x * uv.x - mv.y * uv.y;
            sy += mv.x * uv.y + mv.y * uv.x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_down_sync(0xffffffffu, sx, off);
            sy += __shfl_down_sync(0xffffffffu, sy, off);
        }
        if (lane == 0) {
            const Z prev = accumulate ? y[a] : czero<Z>();
            y[a] = mk<Z>(prev.x + sign * sx, prev.y + sign * sy);
        }
    }
}

__device__ __forceinline__ void prefetch_block(const Z* p, int NN) {
    const char* c = reinterpret_cast<const char*>(p);
    for (int o = threadIdx.x * 128; o < NN * (int)sizeof(Z); o += blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
}

constexpr int kBandMaxNcv = 40;

// Arnoldi steps j0 .. j1-1 of OP = S^-1 M for the active points (one CTA each).  V: [P][ncv+1][n],
// H: [P][ncv+1][ncv] row-major.  w lives in shared memory.
__global__ void __launch_bounds__(256) band_cycle_kernel(BandGeom g, const Z* __restrict__ phases, const int* __restrict__ active,
        const Z* __restrict__ Dinv, const Z* __restrict__ W, const Z* __restrict__ G, Z* V, Z* H, int ncv, int j0, int j1) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    const int N = g.Nx, Ny = g.Ny, m = Ny - 1, NN = N * N;
    const int n = N * Ny;
    const int p = active[blockIdx.x];
    const Z phx = phases[2 * p], phy = phases[2 * p + 1];
    Z* w = reinterpret_cast<Z*>(band_smem);          // n
    Z* du = w + n;                                   // n: U_j couplings (kn), then scratch
    Z* dl = du + n;                                  // n: L_j couplings (ks)
    Z* ym = dl + n;                                  // N
    Z* tmp = ym + N;                                 // N
    Z* hs = tmp + N;                                 // kBandMaxNcv + 1
    __shared__ double red[8];
    const Z* Dp = Dinv + (size_t)p * (m + 1) * NN;
    const Z* Wo = W + (size_t)p * m * NN;
    const Z* Go = G + (size_t)p * m * NN;
    Z* Vp = V + (size_t)p * (ncv + 1) * n;
    Z* Hp = H + (size_t)p * (ncv + 1) * ncv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;

    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int j = e / N, i = e - j * N;
        const Entries en = band_entries(g, phx, phy, i, j);
        dl[e] = en.ks; du[e] = en.kn;
    }
    __syncthreads();

    for (int js = j0; js < j1; ++js) {
        // r = M v_js
        const Z* vj = Vp + (size_t)js * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) w[e] = cmul(g.m[e], vj[e]);
        prefetch_block(Dp, NN);
        __syncthreads();
        // ---- forward: z_j = D_j^-1 (r_j - L_j z_{j-1}), in place ----
        for (int j = 0; j < m; ++j) {
            if (j + 1 <= m) prefetch_block(Dp + (size_t)(j + 1) * NN, NN);
            if (j > 0) {
                for (int a = threadIdx.x; a < N; a += blockDim.x)
                    tmp[a] = csub(w[j * N + a], cmul(dl[j * N + a], w[(j - 1) * N + a]));
            } else {
                for (int a = threadIdx.x; a < N; a += blockDim.x) tmp[a] = w[a];
            }
            __syncthreads();
            warp_rows_matvec<false>(N, Dp + (size_t)j * NN, tmp, nullptr, w + j * N, 1.0, false);
            __syncthreads();
        }
        // y_m = r_m - sum_j G_j z_j   (all z_j known: the blocks are independent -> parallel over (j, row))
        for (int a = threadIdx.x; a < N; a += blockDim.x) ym[a] = w[m * N + a];
        __syncthreads();
        for (int jr = warp; jr < m * N; jr += nw) {
            const int j = jr / N, a = jr - j * N;
            const Z* Gr = Go + (size_t)j * NN + (size_t)a * N;
            double sx = 0.0, sy = 0.0;
            for (int b = lane; b < N; b += 32) {
                const Z mv = Gr[b], uv = w[j * N + b];
                sx += mv.x * uv.x - mv.y * uv.y;
                sy += mv.x * uv.y + mv.y * uv.x;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sx += __shfl_down_sync(0xffffffffu, sx, off);
                sy += __shfl_down_sync(0xffffffffu, sy, off);
            }
            if (lane == 0) { atomicAdd(&ym[a].x, -sx); atomicAdd(&ym[a].y, -sy); }
        }
        __syncthreads();
        warp_rows_matvec<false>(N, Dp + (size_t)m * NN, ym, nullptr, w + m * N, 1.0, false);     // x_m
        __syncthreads();
        // t_j = z_j - W_j x_m for all j < m (parallel over (j, row))
        for (int jr = warp; jr < m * N; jr += nw) {
            const int j = jr / N, a = jr - j * N;
            const Z* Wr = Wo + (size_t)j * NN + (size_t)a * N;
            double sx = 0.0, sy = 0.0;
            for (int b = lane; b < N; b += 32) {
                const Z mv = Wr[b], uv = w[m * N + b];
                sx += mv.x * uv.x - mv.y * uv.y;
                sy += mv.x * uv.y + mv.y * uv.x;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sx += __shfl_down_sync(0xffffffffu, sx, off);
                sy += __shfl_down_sync(0xffffffffu, sy, off);
            }
            if (lane == 0) { w[j * N + a].x -= sx; w[j * N + a].y -= sy; }
        }
        __syncthreads();
        // ---- backward: x_j = t_j - D_j^-1 (U_j x_{j+1}),  j = m-2 .. 0 (x_{m-1} = t_{m-1}) ----
        for (int j = m - 2; j >= 0; --j) {
            if (j > 0) prefetch_block(Dp + (size_t)(j - 1) * NN, NN);
            warp_rows_matvec<true>(N, Dp + (size_t)j * NN, w + (j + 1) * N, du + j * N, w + j * N, -1.0, true);
            __syncthreads();
        }
        // ---- classical Gram-Schmidt, two passes, against V_0 .. V_js ----
        const int nb = js + 1;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) Hp[(size_t)i * ncv + js] = czero<Z>();
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = warp; i < nb; i += nw) {
                const Z* vi = Vp + (size_t)i * n;
                double sx = 0.0, sy = 0.0;
                for (int e = lane; e < n; e += 32) {
                    const Z a = vi[e], b = w[e];
                    sx += a.x * b.x + a.y * b.y;          // conj(v) * w
                    sy += a.x * b.y - a.y * b.x;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    sx += __shfl_down_sync(0xffffffffu, sx, off);
                    sy += __shfl_down_sync(0xffffffffu, sy, off);
                }
                if (lane == 0) hs[i] = mk<Z>(sx, sy);
            }
            __syncthreads();
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                Z acc = w[e];
                for (int i = 0; i < nb; ++i) acc = csub(acc, cmul(hs[i], Vp[(size_t)i * n + e]));
                w[e] = acc;
            }
            for (int i = threadIdx.x; i < nb; i += blockDim.x) {
                const Z h = Hp[(size_t)i * ncv + js];
                Hp[(size_t)i * ncv + js] = cadd(h, hs[i]);
            }
            __syncthreads();
        }
        // ---- norm, next basis vector ----
        double s2 = 0.0;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s2 += w[e].x * w[e].x + w[e].y * w[e].y;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s2 += __shfl_down_sync(0xffffffffu, s2, off);
        if (lane == 0) red[warp] = s2;
        __syncthreads();
        double tot = 0.0;
        for (int q = 0; q < nw; ++q) tot += red[q];
        const double beta = sqrt(tot);
        if (threadIdx.x == 0) Hp[(size_t)(js + 1) * ncv + js] = mk<Z>(beta, 0.0);
        const double inv = beta > 1e-300 ? 1.0 / beta : 0.0;
        Z* vn = Vp + (size_t)(js + 1) * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) vn[e] = cscale(inv, w[e]);
        __syncthreads();
    }
}

// out[r] = sum_c Q[r][c] V[c] for r < rows (Q: [Pa][rows][ncv] row-major, per ACTIVE slot), element-wise in
// place when out == V (restart: rows = keep, then V[rows] = V[ncv]); normalise = unit rows (Ritz vectors)
__global__ void __launch_bounds__(256) band_combine_kernel(const int* __restrict__ active, const Z* __restrict__ Q, Z* V,
        Z* out, int64_t out_stride, int n, int ncv, int rows, int move_last, int normalise) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    Z* q = reinterpret_cast<Z*>(band_smem);                 // rows * ncv
    __shared__ double red[8 * kBandMaxNcv];
    const int slot = blockIdx.x, p = active[slot];
    Z* Vp = V + (size_t)p * (ncv + 1) * n;
    Z* Op = out ? out + (size_t)p * out_stride : Vp;
    for (int e = threadIdx.x; e < rows * ncv; e += blockDim.x) q[e] = Q[(size_t)slot * rows * ncv + e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    double nrm[kBandMaxNcv];
    for (int r = 0; r < rows; ++r) nrm[r] = 0.0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        Z v[kBandMaxNcv];
        for (int c = 0; c < ncv; ++c) v[c] = Vp[(size_t)c * n + e];
        const Z last = Vp[(size_t)ncv * n + e];
        for (int r = 0; r < rows; ++r) {
            Z acc = czero<Z>();
            for (int c = 0; c < ncv; ++c) acc = cfma(q[r * ncv + c], v[c], acc);
            Op[(size_t)r * n + e] = acc;
            if (normalise) nrm[r] += acc.x * acc.x + acc.y * acc.y;
        }
        if (move_last) Vp[(size_t)rows * n + e] = last;
    }
    if (normalise) {
        for (int r = 0; r < rows; ++r) {
            double s = nrm[r];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if (lane == 0) red[r * 8 + warp] = s;
        }
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            double tot = 0.0;
            for (int qw = 0; qw < nw; ++qw) tot += red[r * 8 + qw];
            const double inv = tot > 0 ? 1.0 / sqrt(tot) : 0.0;
            for (int e = threadIdx.x; e < n; e += blockDim.x) Op[(size_t)r * n + e] = cscale(inv, Op[(size_t)r * n + e]);
        }
    }
}

// res[slot][r] = ||A x - lambda M x|| / (||A x|| + |lambda| ||M x||),  A x = S x + sigma M x (matrix-free)
__global__ void __launch_bounds__(256) band_residual_kernel(BandGeom g, const Z* __restrict__ phases, const int* __restrict__ active,
        const Z* __restrict__ X, int64_t x_stride, const Z* __restrict__ lam, int k, double* res) {
    const int slot = blockIdx.x / k, r = blockIdx.x % k, p = active[slot];
    const Z phx = phases[2 * p], phy = phases[2 * p + 1];
    const int N = g.Nx, Ny = g.Ny, n = N * Ny;
    const Z* x = X + (size_t)p * x_stride + (size_t)r * n;
    const Z l = lam[(size_t)slot * k + r];
    double a2 = 0.0, m2 = 0.0, d2 = 0.0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int j = e / N, i = e - j * N;
        const Entries en = band_entries(g, phx, phy, i, j);
        const int iw = i == 0 ? N - 1 : i - 1, ie = i == N - 1 ? 0 : i + 1;
        const int js = j == 0 ? Ny - 1 : j - 1, jn = j == Ny - 1 ? 0 : j + 1;
        Z sx = cmul(en.dg, x[e]);
        sx = cfma(en.kw, x[j * N + iw], sx);
        sx = cfma(en.ke, x[j * N + ie], sx);
        sx = cfma(en.ks, x[js * N + i], sx);
        sx = cfma(en.kn, x[jn * N + i], sx);
        const Z mx = cmul(g.m[e], x[e]);
        const Z ax = mk<Z>(sx.x + g.sigma * mx.x, sx.y + g.sigma * mx.y);
        const Z d = csub(ax, cmul(l, mx));
        a2 += ax.x * ax.x + ax.y * ax.y; m2 += mx.x * mx.x + mx.y * mx.y; d2 += d.x * d.x + d.y * d.y;
    }
    __shared__ double red[3][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a2 += __shfl_down_sync(0xffffffffu, a2, off);
        m2 += __shfl_down_sync(0xffffffffu, m2, off);
        d2 += __shfl_down_sync(0xffffffffu, d2, off);
    }
    if (lane == 0) { red[0][warp] = a2; red[1][warp] = m2; red[2][warp] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A2 = 0, M2 = 0, D2 = 0;
        for (int q = 0; q < nw; ++q) { A2 += red[0][q]; M2 += red[1][q]; D2 += red[2][q]; }
        const double den = sqrt(A2) + sqrt(l.x * l.x + l.y * l.y) * sqrt(M2);
        res[(size_t)slot * k + r] = den > 0 ? sqrt(D2) / den : sqrt(D2);
    }
}

}  // namespace

struct BandPlan {
    BandGeom g;
    int P = 0, ncv = 0, nactive = 0;
    int64_t n = 0;
    Z *phases = nullptr, *Dinv = nullptr, *W = nullptr, *G = nullptr, *V = nullptr, *H = nullptr, *X = nullptr;
    Z *scratch = nullptr;          // Q / lambda uploads
    size_t scratch_bytes = 0;
    double* res = nullptr;
    int* active = nullptr;
    int kmax = 0;
    ~BandPlan() {
        cudaFree(phases); cudaFree(Dinv); cudaFree(W); cudaFree(G); cudaFree(V); cudaFree(H); cudaFree(X);
        cudaFree(scratch); cudaFree(res); cudaFree(active);
    }
};

static size_t band_factor_smem(int N) { return ((size_t)5 * N * N + 3 * N) * sizeof(Z); }
static size_t band_cycle_smem(int N, int Ny) { return ((size_t)3 * N * Ny + 2 * N + kBandMaxNcv + 1) * sizeof(Z); }

}  // namespace fdfd

using namespace fdfd;

struct fdfd_band { BandPlan plan; };

extern "C" int fdfd_band_max_nx(void) {
    int N = 1;
    while (band_factor_smem(N + 1) <= (size_t)220 * 1024) ++N;
    return N;
}

extern "C" int fdfd_band_create(fdfd_band_t** out, int kind, int Nx, int Ny, int P, const void* ca, const void* cb,
                                const void* m, double sx, double sy, double sigma, const double* phases_host,
                                int ncv, int kmax, void* stream) {
    FDFD_CHECK_ARG(out && ca && cb && m && phases_host, "null pointer");
    FDFD_CHECK_ARG(kind == FDFD_POL_TE || kind == FDFD_POL_TM, "operator type must be TE or TM");
    FDFD_CHECK_ARG(Nx >= 2 && Ny >= 2 && P >= 1, "bad cell / batch size");
    FDFD_CHECK_ARG(ncv >= 2 && ncv <= kBandMaxNcv && kmax >= 1 && kmax <= ncv, "ncv / k out of range");
    FDFD_CHECK_ARG(band_factor_smem(Nx) <= (size_t)220 * 1024, "cell too wide for the shared-memory block factorisation");
    FDFD_CHECK_ARG(band_cycle_smem(Nx, Ny) <= (size_t)220 * 1024, "cell too large for the shared-memory Arnoldi vector");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: fdfd_b200 has no CPU path");
        return FDFD_ERR_CUDA;
    }
    cudaStream_t st = (cudaStream_t)stream;
    fdfd_band* h = new fdfd_band();
    BandPlan& pl = h->plan;
    pl.g = BandGeom{kind, Nx, Ny, sx, sy, sigma, (const Z*)ca, (const Z*)cb, (const Z*)m};
    pl.P = P; pl.ncv = ncv; pl.n = (int64_t)Nx * Ny; pl.kmax = kmax; pl.nactive = P;
    const size_t NN = (size_t)Nx * Nx, mm = (size_t)Ny - 1;
    auto fail = [&](cudaError_t e, const char* what) { delete h; return cuda_fail(e, what, __FILE__, __LINE__); };
    cudaError_t e;
    if ((e = cudaMalloc(&pl.phases, sizeof(Z) * 2 * P)) != cudaSuccess) return fail(e, "cudaMalloc(phases)");
    if ((e = cudaMalloc(&pl.Dinv, sizeof(Z) * NN * (mm + 1) * P)) != cudaSuccess) return fail(e, "cudaMalloc(Dinv)");
    if ((e = cudaMalloc(&pl.W, sizeof(Z) * NN * mm * P)) != cudaSuccess) return fail(e, "cudaMalloc(W)");
    if ((e = cudaMalloc(&pl.G, sizeof(Z) * NN * mm * P)) != cudaSuccess) return fail(e, "cudaMalloc(G)");
    if ((e = cudaMalloc(&pl.V, sizeof(Z) * (size_t)(ncv + 1) * pl.n * P)) != cudaSuccess) return fail(e, "cudaMalloc(V)");
    if ((e = cudaMalloc(&pl.H, sizeof(Z) * (size_t)(ncv + 1) * ncv * P)) != cudaSuccess) return fail(e, "cudaMalloc(H)");
    if ((e = cudaMalloc(&pl.X, sizeof(Z) * (size_t)kmax * pl.n * P)) != cudaSuccess) return fail(e, "cudaMalloc(X)");
    pl.scratch_bytes = sizeof(Z) * (size_t)P * ncv * ncv;
    if ((e = cudaMalloc(&pl.scratch, pl.scratch_bytes)) != cudaSuccess) return fail(e, "cudaMalloc(scratch)");
    if ((e = cudaMalloc(&pl.res, sizeof(double) * (size_t)P * kmax)) != cudaSuccess) return fail(e, "cudaMalloc(res)");
    if ((e = cudaMalloc(&pl.active, sizeof(int) * P)) != cudaSuccess) return fail(e, "cudaMalloc(active)");
    std::vector<int> idx(P);
    for (int i = 0; i < P; ++i) idx[i] = i;
    cudaMemcpyAsync(pl.active, idx.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(pl.phases, phases_host, sizeof(Z) * 2 * P, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(pl.H, 0, sizeof(Z) * (size_t)(ncv + 1) * ncv * P, st);
    cudaStreamSynchronize(st);            // idx / phases_host are host temporaries
    const size_t smem = band_factor_smem(Nx);
    if ((e = cudaFuncSetAttribute(band_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return fail(e, "cudaFuncSetAttribute(band_factor_kernel)");
    if ((e = cudaFuncSetAttribute(band_cycle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)band_cycle_smem(Nx, Ny))) != cudaSuccess)
        return fail(e, "cudaFuncSetAttribute(band_cycle_kernel)");
    band_factor_kernel<<<P, 512, smem, st>>>(pl.g, pl.phases, pl.Dinv, pl.W, pl.G);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "band_factor_kernel");
    *out = h;
    return FDFD_OK;
}

extern "C" void fdfd_band_destroy(fdfd_band_t* h) { delete h; }

extern "C" int fdfd_band_set_start(fdfd_band_t* h, const void* v0, void* stream) {
    FDFD_CHECK_ARG(h && v0, "null pointer");
    BandPlan& pl = h->plan;
    // v0: [P][n] unit vectors on the device -> V[:, 0]
    FDFD_CUDA(cudaMemcpy2DAsync(pl.V, sizeof(Z) * (size_t)(pl.ncv + 1) * pl.n, v0, sizeof(Z) * pl.n, sizeof(Z) * pl.n, pl.P,
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FDFD_OK;
}

extern "C" int fdfd_band_set_active(fdfd_band_t* h, const int* idx_host, int count, void* stream) {
    FDFD_CHECK_ARG(h && idx_host && count >= 0 && count <= h->plan.P, "bad active list");
    FDFD_CUDA(cudaMemcpyAsync(h->plan.active, idx_host, sizeof(int) * count, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    FDFD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    h->plan.nactive = count;
    return FDFD_OK;
}

extern "C" int fdfd_band_cycle(fdfd_band_t* h, int j0, int j1, void* H_host, void* stream) {
    FDFD_CHECK_ARG(h && j0 >= 0 && j0 < j1 && j1 <= h->plan.ncv, "bad step range");
    BandPlan& pl = h->plan;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl.nactive > 0) {
        band_cycle_kernel<<<pl.nactive, 256, band_cycle_smem(pl.g.Nx, pl.g.Ny), st>>>(pl.g, pl.phases, pl.active, pl.Dinv, pl.W, pl.G,
                                                                                   pl.V, pl.H, pl.ncv, j0, j1);
        FDFD_CUDA(cudaGetLastError());
    }
    if (H_host) {
        FDFD_CUDA(cudaMemcpyAsync(H_host, pl.H, sizeof(Z) * (size_t)(pl.ncv + 1) * pl.ncv * pl.P, cudaMemcpyDeviceToHost, st));
        FDFD_CUDA(cudaStreamSynchronize(st));
    }
    return FDFD_OK;
}

// Krylov-Schur restart of the active points: V[:keep] = Q V[:ncv], V[keep] = V[ncv]; H <- Hnew ([Pa][ncv+1][ncv])
extern "C" int fdfd_band_restart(fdfd_band_t* h, const void* Q_host, const void* Hnew_host, int keep, void* stream) {
    FDFD_CHECK_ARG(h && Q_host && Hnew_host && keep >= 1 && keep < h->plan.ncv, "bad restart");
    BandPlan& pl = h->plan;
    cudaStream_t st = (cudaStream_t)stream;
    const int Pa = pl.nactive;
    if (Pa == 0) return FDFD_OK;
    FDFD_CUDA(cudaMemcpyAsync(pl.scratch, Q_host, sizeof(Z) * (size_t)Pa * keep * pl.ncv, cudaMemcpyHostToDevice, st));
    band_combine_kernel<<<Pa, 256, sizeof(Z) * (size_t)keep * pl.ncv, st>>>(pl.active, pl.scratch, pl.V, nullptr, 0, (int)pl.n,
                                                                          pl.ncv, keep, 1, 0);
    FDFD_CUDA(cudaGetLastError());
    // projected matrices of the active points (host order = active order) scattered to their slots
    std::vector<int> idx(Pa);
    FDFD_CUDA(cudaMemcpyAsync(idx.data(), pl.active, sizeof(int) * Pa, cudaMemcpyDeviceToHost, st));
    FDFD_CUDA(cudaStreamSynchronize(st));
    const size_t hb = sizeof(Z) * (size_t)(pl.ncv + 1) * pl.ncv;
    for (int a = 0; a < Pa; ++a)
        FDFD_CUDA(cudaMemcpyAsync((char*)pl.H + hb * idx[a], (const char*)Hnew_host + hb * a, hb, cudaMemcpyHostToDevice, st));
    FDFD_CUDA(cudaStreamSynchronize(st));
    return FDFD_OK;
}

// Ritz vectors X[p][:k] = normalise(S V[:ncv]) of the points listed (S_host: [count][k][ncv]) and their true
// residuals for the eigenvalues lam_host ([count][k] complex) -> res_host [count][k]; X stays on the device
// (fdfd_band_vectors returns the pointer)
extern "C" int fdfd_band_ritz(fdfd_band_t* h, const int* idx_host, int count, const void* S_host, const void* lam_host,
                              int k, double* res_host, void* stream) {
    FDFD_CHECK_ARG(h && idx_host && S_host && lam_host && res_host, "null pointer");
    BandPlan& pl = h->plan;
    FDFD_CHECK_ARG(count >= 1 && count <= pl.P && k >= 1 && k <= pl.kmax, "bad Ritz request");
    cudaStream_t st = (cudaStream_t)stream;
    int* didx = nullptr;
    FDFD_CUDA(cudaMalloc(&didx, sizeof(int) * count));
    Z* dlam = nullptr;
    cudaError_t e = cudaMalloc(&dlam, sizeof(Z) * (size_t)count * k);
    if (e != cudaSuccess) { cudaFree(didx); return cuda_fail(e, "cudaMalloc(lam)", __FILE__, __LINE__); }
    cudaMemcpyAsync(didx, idx_host, sizeof(int) * count, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dlam, lam_host, sizeof(Z) * (size_t)count * k, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(pl.scratch, S_host, sizeof(Z) * (size_t)count * k * pl.ncv, cudaMemcpyHostToDevice, st);
    band_combine_kernel<<<count, 256, sizeof(Z) * (size_t)k * pl.ncv, st>>>(didx, pl.scratch, pl.V, pl.X, (int64_t)pl.kmax * pl.n,
                                                                          (int)pl.n, pl.ncv, k, 0, 1);
    band_residual_kernel<<<count * k, 256, 0, st>>>(pl.g, pl.phases, didx, pl.X, (int64_t)pl.kmax * pl.n, dlam, k, pl.res);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(res_host, pl.res, sizeof(double) * (size_t)count * k, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(didx); cudaFree(dlam);
    if (e != cudaSuccess) return cuda_fail(e, "fdfd_band_ritz", __FILE__, __LINE__);
    return FDFD_OK;
}

extern "C" const void* fdfd_band_vectors(fdfd_band_t* h) { return h ? h->plan.X : nullptr; }

// y = S x for one path point with the matrix-free entries the factorisation uses (validation hook: must equal
// the K2 operator with cd = -sigma M) and x <- S^-1 x through the block factors (one Arnoldi-free solve)
namespace fdfd { namespace {
__global__ void band_apply_kernel(BandGeom g, const Z* __restrict__ phases, int p, const Z* __restrict__ x, Z* y) {
    const Z phx = phases[2 * p], phy = phases[2 * p + 1];
    const int N = g.Nx, Ny = g.Ny, n = N * Ny;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int j = e / N, i = e - j * N;
        const Entries en = band_entries(g, phx, phy, i, j);
        const int iw = i == 0 ? N - 1 : i - 1, ie = i == N - 1 ? 0 : i + 1;
        const int js = j == 0 ? Ny - 1 : j - 1, jn = j == Ny - 1 ? 0 : j + 1;
        Z s = cmul(en.dg, x[e]);
        s = cfma(en.kw, x[j * N + iw], s);
        s = cfma(en.ke, x[j * N + ie], s);
        s = cfma(en.ks, x[js * N + i], s);
        s = cfma(en.kn, x[jn * N + i], s);
        y[e] = s;
    }
}
} }

extern "C" int fdfd_band_apply_shifted(fdfd_band_t* h, int p, const void* x, void* y, void* stream) {
    FDFD_CHECK_ARG(h && x && y && p >= 0 && p < h->plan.P, "bad argument");
    BandPlan& pl = h->plan;
    band_apply_kernel<<<(int)((pl.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pl.g, pl.phases, p, (const Z*)x, (Z*)y);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}
