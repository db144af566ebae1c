// K2 — matrix-free 5-point complex Helmholtz / curl-curl stencil (see helm2d.cuh).
//
// Operator (unit-spacing yeeder2d operators, reference Scattering/yee_derivative.py:44-72,
// composed as in Scattering_Solver_2D.py:176-188 / Band_Diagram_Solver.py:312,319):
//   TE-type  y = sx*DHX ca DEX x + sy*DHY cb DEY x + shift*cd x
//            flux on the face between i and i+1 carries ca_i ; far wall: x_N = 0 ;
//            near wall: no flux.
//   TM-type  y = sx*DEX ca DHX x + sy*DEY cb DHY x + shift*cd x
//            flux on the face between i-1 and i carries ca_i ; near wall: x_-1 = 0 ;
//            far wall: no flux.
//   Bloch axis: x_N = ph*x_0, x_-1 = conj(ph)*x_{N-1}, coefficients periodic, and the centre
//            weight of the wrapped face is |ph|^2 (exactly what -D^H c D gives).
//
// Kernel shape: bandwidth bound (80 B/pt in complex128: x, y, ca, cb, cd once each).  One
// thread owns one grid column of a (blockDim.x wide, rows_per_cta tall) tile and marches in y
// with a three-row register window, so every x row and every cb row is fetched from L2/HBM
// once per tile (+1 halo row); the x-neighbours and the second ca value come from L1 (the
// row was just loaded by the adjacent lanes).  All global accesses are 16-byte, unit-stride
// across the warp.
#include <cstdlib>

#include "helm2d.cuh"
#include "comm.cuh"

namespace fdfd {

// Register budget: the loop needs ~70 registers (c128) to keep all seven loads of a row in flight;
// capping it lower makes ptxas serialise the loads into 3-4 dependent round trips per row
// (measured: 56 regs / 9 CTAs per SM = 5.3 TB/s, 70 regs / 7 CTAs = 6.35 TB/s at 4096^2).
#ifndef FDFD_C64_SMOOTH_MINB
#define FDFD_C64_SMOOTH_MINB 8
#endif
#ifndef FDFD_C64_AX_MINB
#define FDFD_C64_AX_MINB 9
#endif
template <typename C> __host__ __device__ constexpr int default_min_blocks(int mode) {
    return sizeof(C) == 16 ? (mode >= APPLY_JACOBI ? 6 : 7) : (mode >= APPLY_JACOBI ? FDFD_C64_SMOOTH_MINB : FDFD_C64_AX_MINB);
}

// load an element of the input vector, stored as X (complex64 input of a complex128 operator: the
// flexible-GMRES corrections are kept in complex64), widened to the arithmetic type C
template <typename C, typename X> __device__ __forceinline__ C ldx(const X* p) {
    const X v = *p;
    return mk<C>((typename real_of<C>::type)v.x, (typename real_of<C>::type)v.y);
}

// ---- ghost rows over NVLink peer memory (y-slab sharding, P2P = true) -------------------------------
// The first row of CTAs of a launch PUSHES this rank's two edge rows of x straight into the neighbours'
// receive buffers (peer-mapped symmetric heap, comm.cu) and raises one flag per column segment; the two
// rows of tiles that consume ghost rows are issued LAST (the tile index is rotated), so by the time they
// run the neighbour's push — done by the first CTAs of its own launch — has normally landed and the
// wait costs nothing.  No NCCL call, no extra launch, no ghost copy.  Receive buffers and flags are double
// buffered by the parity of a device-resident epoch (a rank can be at most one exchange ahead of its
// neighbour), which the last participating CTA advances, so the launch is CUDA-graph capturable.
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <typename C, typename X> __device__ __forceinline__ C ldx_cg(const X* p) {
    const X v = __ldcg(p);
    return mk<C>((typename real_of<C>::type)v.x, (typename real_of<C>::type)v.y);
}

template <typename C, bool TM, int MODE, typename X = C, bool P2P = false, int MINB = default_min_blocks<C>(MODE)>
__global__ void __launch_bounds__(128, MINB) helm2d_march_kernel(HelmArgs<C> a) {
    using R = typename real_of<C>::type;
    const X* xall = reinterpret_cast<const X*>(a.x);
    const X* xsouth = reinterpret_cast<const X*>(a.x_south);
    const X* xnorth = reinterpret_cast<const X*>(a.x_north);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int Nx = a.Nx, Ny = a.Ny;
    int tile = blockIdx.y;
    if (P2P) {
        // Everything the exchange needs happens here, before the marching loop, so that the loop itself is
        // the same code as in the unsharded kernel (its register budget is what keeps all loads of a row in
        // flight); the epilogue re-reads the epoch instead of keeping it live.
        const int nt = gridDim.y;
        // pusher row first, ghost-consuming tiles (0 and nt-1) last
        if (nt >= 3) tile = (int)blockIdx.y == nt - 1 ? nt - 1 : ((int)blockIdx.y == nt - 2 ? 0 : (int)blockIdx.y + 1);
        if (blockIdx.y == 0 || tile == 0 || tile == nt - 1) {
            const unsigned int ep = ld_volatile_u32(a.halo.state), par = ep & 1u;
            if (blockIdx.y == 0) {
                if (i < Nx) {
                    if (a.halo.push_north) reinterpret_cast<X*>(a.halo.push_north)[(size_t)par * Nx + i] = xall[(int64_t)(Ny - 1) * Nx + i];
                    if (a.halo.push_south) reinterpret_cast<X*>(a.halo.push_south)[(size_t)par * Nx + i] = xall[i];
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    __threadfence_system();
                    if (a.halo.flag_push_north) st_release_sys_u32(a.halo.flag_push_north + par * a.halo.segs + blockIdx.x, ep + 1);
                    if (a.halo.flag_push_south) st_release_sys_u32(a.halo.flag_push_south + par * a.halo.segs + blockIdx.x, ep + 1);
                }
            }
            // the ghost rows this tile will read must have landed (the neighbour pushed them at the start of
            // its own launch; these tiles are the last ones issued here)
            if (tile == 0 && a.halo.flag_south)
                while (ld_acquire_sys_u32(a.halo.flag_south + par * a.halo.segs + blockIdx.x) != ep + 1) { }
            if (tile == nt - 1 && a.halo.flag_north)
                while (ld_acquire_sys_u32(a.halo.flag_north + par * a.halo.segs + blockIdx.x) != ep + 1) { }
        }
    }
    const int jb = a.row_begin + tile * a.rows_per_cta;
    const int je = min(jb + a.rows_per_cta, a.row_end);
    if (i < Nx && jb < je) {

    // x-direction neighbour bookkeeping (column constants)
    const bool west_edge = (i == 0), east_edge = (i == Nx - 1);
    const int iw = west_edge ? Nx - 1 : i - 1;
    const int ie = east_edge ? 0 : i + 1;
    // multipliers that realise wall / Bloch behaviour on the two x faces
    R w_on = 1, e_on = 1;      // face present?
    R w_c2 = 1, e_c2 = 1;      // centre weight on that face
    bool w_ph = false, e_ph = false;
    if (west_edge) {
        if (a.bcx) { w_ph = true; if (!TM) w_c2 = a.phx2; }
        else if (!TM) w_on = 0;          // TE: no flux through the near wall
    }
    if (east_edge) {
        if (a.bcx) { e_ph = true; if (TM) e_c2 = a.phx2; }
        else if (TM) e_on = 0;           // TM: no flux through the far wall
    }
    const bool west_wall_val = west_edge && !a.bcx;   // x_-1 = 0
    const bool east_wall_val = east_edge && !a.bcx;   // x_N  = 0

    const X* xcol = xall + i;
    const C* cbcol = a.cb + i;
    // APPLY_SMOOTH: is this column relaxed by y-lines?
    bool in_strip = false;
    int line_idx = 0, line_nl = 0;
    if (MODE == APPLY_SMOOTH) {
        line_nl = a.strip_lo + a.strip_hi;
        if (i < a.strip_lo) { in_strip = true; line_idx = i; }
        else if (i >= Nx - a.strip_hi) { in_strip = true; line_idx = a.strip_lo + (i - (Nx - a.strip_hi)); }
    }

    // ---- prime the register window --------------------------------------------------------
    C xs, xc, cb_lo;  // cb_lo: TE -> cb[j-1] ; TM -> cb[j]
    if (jb > 0) {
        xs = ldx<C>(xcol + (int64_t)(jb - 1) * Nx);
    } else if (P2P ? a.halo.flag_south != nullptr : xsouth != nullptr) {
        // (peer memory: the parity half of the receive buffer is looked up again here and at the north edge,
        // rather than kept in registers across the loop; the epoch cannot move before this CTA checks in)
        if (P2P) xs = ldx_cg<C>(reinterpret_cast<const X*>(a.halo.rx_south) + (size_t)(ld_volatile_u32(a.halo.state) & 1u) * Nx + i);
        else xs = ldx<C>(xsouth + i);
        if (a.south_wrap) xs = cmulc(a.phy, xs);      // conj(phy) * x_{Ny-1}
    } else {
        xs = czero<C>();
    }
    xc = ldx<C>(xcol + (int64_t)jb * Nx);
    if (!TM) {
        if (jb > 0) cb_lo = ld_stream(cbcol + (int64_t)(jb - 1) * Nx);
        else if (a.cb_ghost) cb_lo = a.cb_ghost[i];
        else cb_lo = czero<C>();                      // TE: no flux through the near wall
    } else {
        cb_lo = ld_stream(cbcol + (int64_t)jb * Nx);
    }

    for (int j = jb; j < je; ++j) {
        const int64_t row = (int64_t)j * Nx;
        // north value and the second cb of this row
        C xn, cb_hi;
        R n_on = 1, s_c2 = 1, n_c2 = 1;
        if (j + 1 < Ny) {
            xn = ldx<C>(xcol + row + Nx);
            cb_hi = TM ? ld_stream(cbcol + row + Nx) : ld_stream(cbcol + row);
        } else {
            if (P2P ? a.halo.flag_north != nullptr : xnorth != nullptr) {
                if (P2P) xn = ldx_cg<C>(reinterpret_cast<const X*>(a.halo.rx_north) + (size_t)(ld_volatile_u32(a.halo.state) & 1u) * Nx + i);
                else xn = ldx<C>(xnorth + i);
                if (a.north_wrap) { xn = cmul(a.phy, xn); if (TM) n_c2 = a.phy2; }
            } else {
                xn = czero<C>();
            }
            if (TM) {
                if (a.cb_ghost) cb_hi = a.cb_ghost[i];
                else { cb_hi = czero<C>(); n_on = 0; }   // TM: no flux through the far wall
            } else {
                cb_hi = ld_stream(cbcol + row);
            }
        }
        if (!TM && j == 0 && a.south_wrap) s_c2 = a.phy2;

        // x neighbours of the centre row (L1 hits) and the two ca values
        C xw = ldx<C>(xall + row + iw), xe = ldx<C>(xall + row + ie);
        if (w_ph) xw = cmulc(a.phx, xw);
        if (e_ph) xe = cmul(a.phx, xe);
        if (west_wall_val) xw = czero<C>();
        if (east_wall_val) xe = czero<C>();
        C ca_w, ca_e;
        if (TM) { ca_w = a.ca[row + i]; ca_e = a.ca[row + ie]; }
        else    { ca_w = a.ca[row + iw]; ca_e = a.ca[row + i]; }

        // fluxes
        C fe = cmul(ca_e, csub(xe, cscale(e_c2, xc)));
        C fw = cmul(ca_w, csub(cscale(w_c2, xc), xw));
        C tx = csub(cscale(e_on, fe), cscale(w_on, fw));
        C c_s = TM ? cb_lo : cb_lo;      // south face coefficient: TE cb[j-1], TM cb[j]
        C c_n = cb_hi;                   // north face coefficient: TE cb[j],   TM cb[j+1]
        C fn = cmul(c_n, csub(xn, cscale(n_c2, xc)));
        C fs = cmul(c_s, csub(cscale(s_c2, xc), xs));
        C ty = csub(cscale(n_on, fn), fs);
        C ax = cadd(cscale(a.sx, tx), cscale(a.sy, ty));
        C dterm = czero<C>();
        if (a.cd) {
            dterm = cmul(a.shift, ld_stream(a.cd + row + i));
            ax = cfma(dterm, xc, ax);
        }

        C out;
        if (MODE == APPLY_AX) {
            out = ax;
        } else if (MODE == APPLY_RESID) {
            out = csub(ld_stream(a.b + row + i), ax);
        } else {
            // diagonal of A at this point
            C dg = cadd(cscale(-a.sx, cadd(cscale(e_on * e_c2, ca_e), cscale(w_on * w_c2, ca_w))),
                        cscale(-a.sy, cadd(cscale(n_on * n_c2, c_n), cscale(s_c2, c_s))));
            dg = cadd(dg, dterm);
            if (MODE == APPLY_JACOBI) {
                C r = csub(ld_stream(a.b + row + i), ax);
                out = cadd(xc, cscale(a.omega, cdiv(r, dg)));
            } else if (MODE == APPLY_SMOOTH) {
                C r = csub(ld_stream(a.b + row + i), ax);
                if (in_strip) {
                    out = xc;
                    a.line_rhs[(int64_t)(a.row0 + j) * line_nl + line_idx] = r;
                } else {
                    out = cadd(xc, cscale(a.omega, cdiv(r, dg)));
                }
            } else if (MODE == APPLY_DINV) {
                out = cdiv(mk<C>(1, 0), dg);
            } else {
                out = cdiv(ax, dg);
            }
        }
        a.y[row + i] = out;

        // rotate the window
        xs = xc; xc = xn;
        cb_lo = TM ? cb_hi : cb_hi;
    }
    }   // active column / non-empty tile
    if (P2P) {
        const int nt = gridDim.y;
        if (blockIdx.y == 0 || (int)blockIdx.y >= nt - 2) {      // the CTAs that read the epoch
            // the last of them advances it (nobody can have advanced it yet: this CTA has not checked in)
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned int ep = ld_volatile_u32(a.halo.state);
                const unsigned int rows = nt >= 3 ? 3u : (unsigned int)nt;
                const unsigned int prev = atomicAdd(a.halo.state + 1, 1u);
                if (prev == rows * gridDim.x - 1) {
                    a.halo.state[1] = 0;
                    __threadfence();
                    a.halo.state[0] = ep + 1;
                }
            }
        }
    }
}

struct Variant { int bx, ry; };
static const Variant kVariants[] = {
    {128, 8}, {128, 16}, {128, 32}, {64, 4}, {64, 32}, {64, 8}, {128, 4}, {64, 16}, {128, 2},
};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <typename C>
HelmArgs<C> helm2d_args(const Helm2D& op, const C* x, C* y, const C* b, const C* ca, const C* cb,
                        const C* cd, double shift_re, double shift_im, double omega) {
    using R = typename real_of<C>::type;
    HelmArgs<C> a;
    a.x = x; a.y = y; a.b = b; a.ca = ca; a.cb = cb; a.cd = cd;
    a.Nx = op.Nx; a.Ny = op.Ny;
    a.bcx = (op.bcx == FDFD_BC_BLOCH) ? 1 : 0;
    a.sx = (R)op.sx; a.sy = (R)op.sy;
    a.phx = mk<C>((R)op.phase[0], (R)op.phase[1]);
    a.phy = mk<C>((R)op.phase[2], (R)op.phase[3]);
    a.phx2 = (R)(op.phase[0] * op.phase[0] + op.phase[1] * op.phase[1]);
    a.phy2 = (R)(op.phase[2] * op.phase[2] + op.phase[3] * op.phase[3]);
    a.shift = mk<C>((R)shift_re, (R)shift_im);
    a.omega = (R)omega;
    a.strip_lo = a.strip_hi = 0; a.line_rhs = nullptr; a.row0 = 0;
    a.row_begin = 0; a.row_end = op.Ny;
    a.halo = op.halo;
    // ghost rows
    const bool first = op.j0 == 0, last = op.j0 + op.Ny == op.Ny_global;
    const bool sharded = op.Ny != op.Ny_global;
    a.x_south = nullptr; a.x_north = nullptr; a.cb_ghost = nullptr;
    a.south_wrap = a.north_wrap = 0;
    const bool TM = op.pol == FDFD_POL_TM;
    if (!sharded) {
        if (op.bcy == FDFD_BC_BLOCH) {
            a.x_south = x + (int64_t)(op.Ny - 1) * op.Nx;
            a.x_north = x;
            a.south_wrap = a.north_wrap = 1;
            a.cb_ghost = TM ? cb : cb + (int64_t)(op.Ny - 1) * op.Nx;
        }
    } else {
        const bool bloch_y = op.bcy == FDFD_BC_BLOCH;
        if (!first || bloch_y) { a.x_south = (const C*)op.halo_south; a.south_wrap = first && bloch_y; }
        if (!last || bloch_y) { a.x_north = (const C*)op.halo_north; a.north_wrap = last && bloch_y; }
        const bool need = TM ? (!last || bloch_y) : (!first || bloch_y);
        if (need) a.cb_ghost = (const C*)op.cb_ghost_buf;
    }
    return a;
}

template <typename C>
int helm2d_launch_args(const Helm2D& op, int mode, HelmArgs<C>& a, cudaStream_t stream) {
    const bool TM = op.pol == FDFD_POL_TM;
    const int shape = op.variant;
    Variant v = kVariants[(shape >= 0 && shape < kNumVariants) ? shape : 0];
    if (shape == 0) {
        // automatic shape: small grids (coarse multigrid levels) are latency bound and need more,
        // shorter CTAs to fill 148 SMs; large grids amortise the halo row over 8 rows
        const int64_t n = (int64_t)op.Nx * op.Ny;
        if (n < ((int64_t)1 << 20)) v = Variant{64, 2};
        else if (n < ((int64_t)1 << 22)) v = Variant{128, 4};
    }
    int bx = v.bx;
    while (bx > 32 && bx / 2 >= op.Nx) bx /= 2;
    a.rows_per_cta = v.ry;
    if (a.row_end <= a.row_begin) return FDFD_OK;
    dim3 grid(ceil_div(op.Nx, bx), ceil_div(a.row_end - a.row_begin, v.ry));
    FDFD_CHECK_ARG(grid.y <= 65535, "too many row tiles for one launch");
    // peer-memory ghost rows: the launch itself exchanges the edge rows of a.x (whole-slab launches only;
    // APPLY_DINV does not read x)
    const bool p2p = a.halo.state != nullptr && mode != APPLY_DINV && a.row_begin == 0 && a.row_end == op.Ny;
    if (!p2p) a.halo = HaloP2P();
    // (the exchange prologue / epilogue keep nothing live across the marching loop, so the sharded variant
    // runs with the launch bounds of the unsharded one: measured 0.2124 ms per 4096^2 apply on 2 GPUs against
    // 0.2074 ms unsharded; one CTA less per SM cost 9 us more)
#define LAUNCH(TMv, MODEv) do { \
        if (p2p) helm2d_march_kernel<C, TMv, MODEv, C, true><<<grid, bx, 0, stream>>>(a); \
        else helm2d_march_kernel<C, TMv, MODEv><<<grid, bx, 0, stream>>>(a); } while (0)
#define LAUNCH_MODE(MODEv) do { if (TM) LAUNCH(true, MODEv); else LAUNCH(false, MODEv); } while (0)
    switch (mode) {
        case APPLY_AX: LAUNCH_MODE(APPLY_AX); break;
        case APPLY_RESID: LAUNCH_MODE(APPLY_RESID); break;
        case APPLY_JACOBI: LAUNCH_MODE(APPLY_JACOBI); break;
        case APPLY_DINV_AX: LAUNCH_MODE(APPLY_DINV_AX); break;
        case APPLY_DINV: LAUNCH_MODE(APPLY_DINV); break;
        case APPLY_SMOOTH: LAUNCH_MODE(APPLY_SMOOTH); break;
        default: set_error("bad apply mode %d", mode); return FDFD_ERR_ARG;
    }
#undef LAUNCH
#undef LAUNCH_MODE
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

template <typename C>
int helm2d_launch(const Helm2D& op, int mode, const C* x, C* y, const C* b, const C* ca,
                  const C* cb, const C* cd, double shift_re, double shift_im, double omega,
                  cudaStream_t stream) {
    HelmArgs<C> a = helm2d_args<C>(op, x, y, b, ca, cb, cd, shift_re, shift_im, omega);
    return helm2d_launch_args<C>(op, mode, a, stream);
}

template int helm2d_launch<double2>(const Helm2D&, int, const double2*, double2*, const double2*,
                                    const double2*, const double2*, const double2*, double, double,
                                    double, cudaStream_t);
template int helm2d_launch<float2>(const Helm2D&, int, const float2*, float2*, const float2*,
                                   const float2*, const float2*, const float2*, double, double,
                                   double, cudaStream_t);
template HelmArgs<double2> helm2d_args<double2>(const Helm2D&, const double2*, double2*, const double2*,
        const double2*, const double2*, const double2*, double, double, double);
template HelmArgs<float2> helm2d_args<float2>(const Helm2D&, const float2*, float2*, const float2*,
        const float2*, const float2*, const float2*, double, double, double);
template int helm2d_launch_args<double2>(const Helm2D&, int, HelmArgs<double2>&, cudaStream_t);
template int helm2d_launch_args<float2>(const Helm2D&, int, HelmArgs<float2>&, cudaStream_t);

// Ghost-row buffers of a y-slab operator and the one-off exchange of the static coefficient ghost:
// TE needs cb row j0-1 (the south rank's last row), TM needs cb row j0+Ny (the north rank's first).
int helm2d_setup_sharding(Helm2D& op, cudaStream_t st) {
    op.halo_south = op.halo_north = op.cb_ghost_buf = nullptr;
    op.halo = HaloP2P();
    if (op.Ny == op.Ny_global) return FDFD_OK;
    FDFD_CHECK_ARG(op.comm && op.comm->nranks > 1, "sharded operator needs a communicator");
    const size_t rb = (size_t)op.Nx * op.elem();
    FDFD_CUDA(cudaMalloc(&op.halo_south, rb));
    FDFD_CUDA(cudaMalloc(&op.halo_north, rb));
    FDFD_CUDA(cudaMalloc(&op.cb_ghost_buf, rb));
    FDFD_CUDA(cudaMemsetAsync(op.halo_south, 0, rb, st));
    FDFD_CUDA(cudaMemsetAsync(op.halo_north, 0, rb, st));
    FDFD_CUDA(cudaMemsetAsync(op.cb_ghost_buf, 0, rb, st));
    const int P = op.comm->nranks, r = op.comm->rank;
    const bool per = op.bcy == FDFD_BC_BLOCH;
    const int south = r > 0 ? r - 1 : (per ? P - 1 : -1);
    const int north = r < P - 1 ? r + 1 : (per ? 0 : -1);
    if (p2p_enabled(op.comm)) FDFD_TRY(p2p_halo_create(op.comm, op.Nx, south, north, &op.halo, op.halo_offs, st));
    if (op.pol == FDFD_POL_TE)
        return comm_sendrecv(op.comm, (const char*)op.cb + (size_t)(op.Ny - 1) * rb, north, op.cb_ghost_buf, south, rb, st);
    return comm_sendrecv(op.comm, op.cb, south, op.cb_ghost_buf, north, rb, st);
}

void helm2d_release_sharding(Helm2D& op) {
    cudaFree(op.halo_south); cudaFree(op.halo_north); cudaFree(op.cb_ghost_buf);
    if (op.halo.state) p2p_halo_destroy(op.comm, &op.halo, op.halo_offs);
    op.halo_south = op.halo_north = op.cb_ghost_buf = nullptr;
}

// Stand-alone ghost-row exchange into op.halo_south / op.halo_north (consumers other than the stencil
// kernel: the prolongation reads one coarse row beyond the slab).
int helm2d_halo_exchange(Helm2D& op, const void* x, cudaStream_t stream) {
    if (op.Ny == op.Ny_global) return FDFD_OK;
    if (op.halo.state)
        return p2p_halo_exchange(op.halo, x, op.halo_south, op.halo_north, op.Nx, op.Ny, op.elem(), stream);
    return comm_halo_rows(op.comm, x, op.halo_south, op.halo_north, op.Nx, op.Ny, op.elem(),
                          op.bcy == FDFD_BC_BLOCH, stream);
}

template <typename C>
static int helm2d_apply_t(Helm2D& op, int mode, const C* x, C* y, const C* b, double omega, cudaStream_t stream) {
    HelmArgs<C> a = helm2d_args<C>(op, x, y, b, (const C*)op.ca, (const C*)op.cb, (const C*)op.cd,
                                   op.shift_re, op.shift_im, omega);
    // sharded: with a peer-mapped heap the launch exchanges its own ghost rows, else NCCL does it first
    if (op.Ny != op.Ny_global && !op.halo.state) FDFD_TRY(helm2d_halo_exchange(op, x, stream));
    return helm2d_launch_args<C>(op, mode, a, stream);
}

// y = A x for a complex128 operator whose input vector is stored in complex64 (72 B/pt instead of 80)
int helm2d_apply_x32(Helm2D& op, const float2* x, double2* y, cudaStream_t stream) {
    FDFD_CHECK_ARG(op.dtype == FDFD_C128, "mixed-input apply needs a complex128 operator");
    using C = double2;
    HelmArgs<C> a = helm2d_args<C>(op, reinterpret_cast<const C*>(x), y, nullptr, (const C*)op.ca, (const C*)op.cb,
                                   (const C*)op.cd, op.shift_re, op.shift_im, 0.0);
    const bool sharded = op.Ny != op.Ny_global;
    if (!sharded && op.bcy == FDFD_BC_BLOCH) {       // the wrap rows are addressed in float2 elements
        a.x_south = reinterpret_cast<const C*>(x + (int64_t)(op.Ny - 1) * op.Nx);
        a.x_north = reinterpret_cast<const C*>(x);
    }
    const bool p2p = sharded && op.halo.state != nullptr;
    if (sharded && !p2p)
        FDFD_TRY(comm_halo_rows(op.comm, x, op.halo_south, op.halo_north, op.Nx, op.Ny, sizeof(float2),
                                op.bcy == FDFD_BC_BLOCH, stream));
    Variant v = kVariants[(op.variant >= 0 && op.variant < kNumVariants) ? op.variant : 0];
    if (op.variant == 0) {
        const int64_t n = (int64_t)op.Nx * op.Ny;
        if (n < ((int64_t)1 << 20)) v = Variant{64, 2};
        else if (n < ((int64_t)1 << 22)) v = Variant{128, 4};
    }
    int bx = v.bx;
    while (bx > 32 && bx / 2 >= op.Nx) bx /= 2;
    a.rows_per_cta = v.ry;
    dim3 grid(ceil_div(op.Nx, bx), ceil_div(op.Ny, v.ry));
    FDFD_CHECK_ARG(grid.y <= 65535, "too many row tiles for one launch");
    // (the widening loads cost this variant two registers more than the 72 that 7 CTAs per SM allow: 6 CTAs)
    if (p2p) {
        if (op.pol == FDFD_POL_TM) helm2d_march_kernel<C, true, APPLY_AX, float2, true, 6><<<grid, bx, 0, stream>>>(a);
        else helm2d_march_kernel<C, false, APPLY_AX, float2, true, 6><<<grid, bx, 0, stream>>>(a);
    } else {
        a.halo = HaloP2P();
        if (op.pol == FDFD_POL_TM) helm2d_march_kernel<C, true, APPLY_AX, float2, false, 6><<<grid, bx, 0, stream>>>(a);
        else helm2d_march_kernel<C, false, APPLY_AX, float2, false, 6><<<grid, bx, 0, stream>>>(a);
    }
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

int helm2d_apply(Helm2D& op, int mode, const void* x, void* y, const void* b, double omega,
                 cudaStream_t stream) {
    if (op.dtype == FDFD_C128)
        return helm2d_apply_t<double2>(op, mode, (const double2*)x, (double2*)y, (const double2*)b, omega, stream);
    return helm2d_apply_t<float2>(op, mode, (const float2*)x, (float2*)y, (const float2*)b, omega, stream);
}

}  // namespace fdfd
