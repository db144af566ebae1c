// K2 — matrix-free 5-point complex Helmholtz / curl-curl operator on a Yee grid (slab).
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace fdfd {

struct Comm;  // comm.cuh

// What a kernel launch computes from the stencil value  Ax  at each point.
enum ApplyMode {
    APPLY_AX = 0,        // y = A x
    APPLY_RESID = 1,     // y = b - A x
    APPLY_JACOBI = 2,    // y = x + omega * (b - A x) / diag(A)      (damped-Jacobi sweep)
    APPLY_DINV_AX = 3,   // y = (A x) / diag(A)
    APPLY_DINV = 4,      // y = 1 / diag(A)              (x ignored)
    APPLY_SMOOTH = 5,    // damped Jacobi outside the x-strips; inside them y = x and the residual
                         // b - A x is written to the line-solver workspace (see mg.cu)
};

template <typename C>
struct HelmArgs {
    using R = typename real_of<C>::type;
    const C* x;
    C* y;
    const C* b;          // APPLY_RESID / APPLY_JACOBI
    const C* ca;
    const C* cb;
    const C* cd;         // may be null (no diagonal term)
    const C* x_south;    // ghost row j = -1 (null = wall)
    const C* x_north;    // ghost row j = Ny (null = wall)
    const C* cb_ghost;   // TE: cb row j = -1 ; TM: cb row j = Ny ; null = wall
    int Nx, Ny;          // local extent
    int bcx;             // Bloch wrap along x
    int south_wrap, north_wrap;  // ghost rows belong to the Bloch wrap (phase applies)
    int rows_per_cta;
    R sx, sy;
    C phx, phy;          // exp(-1j k N d)
    R phx2, phy2;        // |ph|^2 (1 for real Bloch vectors)
    C shift;             // cd is multiplied by (shift) — (1,0) for A, (b1, b2) for the shifted preconditioner
    R omega;             // Jacobi damping
    // APPLY_SMOOTH: columns [0, strip_lo) and [Nx - strip_hi, Nx) belong to y-line relaxation
    int strip_lo, strip_hi;
    C* line_rhs;         // workspace [row][line], line = strip column index, row stride = strip_lo + strip_hi
    int row0;            // global row of local row 0 (sharded lines)
    int row_begin, row_end;  // rows this launch covers (the pipelined host-buffer apply works in row chunks)
    HaloP2P halo;            // peer-memory ghost rows (sharded operators with a mapped heap; see helm2d.cu)
};

struct Helm2D {
    int dtype, pol;
    int Nx, Ny_global, j0, Ny;  // Ny = local rows
    int bcx, bcy;
    double phase[4];
    double sx, sy;
    double shift_re, shift_im;   // multiplier of the diagonal term cd (1,0 for A itself)
    const void *ca, *cb, *cd;
    Comm* comm;
    int variant;
    // ghost rows for sharded operation (device, Nx elements each): landing buffers of the NCCL exchange and
    // of the stand-alone peer-memory exchange (the stencil kernel itself reads the parity buffers of `halo`)
    void *halo_south, *halo_north, *cb_ghost_buf;
    HaloP2P halo;            // peer-memory exchange state (halo.state == null: NCCL path)
    size_t halo_offs[4];     // its allocations in the symmetric heap
    // scratch for *_host entry points
    void *h_x, *h_y;
    int64_t n_local() const { return (int64_t)Nx * Ny; }
    size_t elem() const { return dtype == FDFD_C128 ? 16 : 8; }
};

// Fill the kernel argument block for operator `op` (ghost rows resolved from op's halo buffers).
template <typename C>
HelmArgs<C> helm2d_args(const Helm2D& op, const C* x, C* y, const C* b, const C* ca, const C* cb,
                        const C* cd, double shift_re, double shift_im, double omega);
template <typename C>
int helm2d_launch_args(const Helm2D& op, int mode, HelmArgs<C>& a, cudaStream_t stream);

// Launch one stencil pass.  x_south/x_north/cb_ghost resolved by the caller (halo exchange).
template <typename C>
int helm2d_launch(const Helm2D& op, int mode, const C* x, C* y, const C* b,
                  const C* ca, const C* cb, const C* cd, double shift_re, double shift_im,
                  double omega, cudaStream_t stream);
int helm2d_halo_exchange(Helm2D& op, const void* x, cudaStream_t stream);
int helm2d_setup_sharding(Helm2D& op, cudaStream_t stream);
void helm2d_release_sharding(Helm2D& op);

// y = A x for a complex128 operator with the input vector stored in complex64 (APPLY_AX only)
int helm2d_apply_x32(Helm2D& op, const float2* x, double2* y, cudaStream_t stream);

// y = A x including halo exchange when sharded (dispatches on dtype).
int helm2d_apply(Helm2D& op, int mode, const void* x, void* y, const void* b, double omega,
                 cudaStream_t stream);

}  // namespace fdfd
