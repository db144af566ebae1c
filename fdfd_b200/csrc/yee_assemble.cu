// K1 — Yee derivative-operator assembly on the GPU.
//
// Replaces the DIA->CSR / CSR+CSR scipy path of yeeder2d (reference
// Scattering/yee_derivative.py:44-72): one thread per matrix row writes indptr, the (sorted)
// column indices and the values of DEX or DEY in closed form, and optionally the values of
// DHX/DHY = -conj (the CSC index arrays of the H operators are the CSR arrays of the E ones,
// yee_derivative.py:71-72, so they are not written twice).
//
// Row layout (n = i + Nx*j, "pos" = i for the x axis, j for the y axis, N = axis length):
//   N == 1                : [n]                    value diag (= -1j*k)
//   pos <  N-1            : [n, n+stride]          values diag, fwd
//   pos == N-1, Dirichlet : [n]                    value diag         (explicit zero dropped)
//   pos == N-1, Bloch     : [n-span, n]            values wrap, diag  (wrap column sorts first)
// Row offsets are closed form, so no scan is needed and every thread writes independently.
#include "common.cuh"

namespace fdfd {

struct YeeFillArgs {
    int Nx, Ny, axis, bc;
    int is_complex;
    double diag_re, diag_im, fwd_re, fwd_im, wrap_re, wrap_im;
    int32_t* indptr;
    int32_t* indices;
    void* e_data;
    void* h_data;
};

template <bool COMPLEX>
__device__ __forceinline__ void put(void* e_data, void* h_data, int64_t at, double re, double im) {
    if (COMPLEX) {
        reinterpret_cast<double2*>(e_data)[at] = make_double2(re, im);
        // -conj(re + j im) = -re + j im
        if (h_data) reinterpret_cast<double2*>(h_data)[at] = make_double2(-re, im);
    } else {
        reinterpret_cast<double*>(e_data)[at] = re;
        if (h_data) reinterpret_cast<double*>(h_data)[at] = -re;
    }
}

template <bool COMPLEX>
__global__ void __launch_bounds__(256) yee2d_fill_kernel(YeeFillArgs a) {
    const int64_t M = (int64_t)a.Nx * a.Ny;
    const int N = a.axis == 0 ? a.Nx : a.Ny;
    const int64_t stride = a.axis == 0 ? 1 : a.Nx;
    const int64_t span = a.axis == 0 ? (a.Nx - 1) : (M - a.Nx);
    const bool periodic = a.bc == FDFD_BC_BLOCH;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < M;
         n += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(n % a.Nx);
        const int j = (int)(n / a.Nx);
        const int pos = a.axis == 0 ? i : j;
        int64_t off;
        if (N == 1) {
            off = n;
        } else if (periodic) {
            off = 2 * n;
        } else if (a.axis == 0) {
            off = (int64_t)j * (2 * a.Nx - 1) + 2 * i;
        } else {
            const int64_t full = M - a.Nx;  // rows with two entries
            off = n < full ? 2 * n : 2 * full + (n - full);
        }
        a.indptr[n] = (int32_t)off;
        if (N == 1) {
            a.indices[off] = (int32_t)n;
            put<COMPLEX>(a.e_data, a.h_data, off, a.diag_re, a.diag_im);
            if (n == M - 1) a.indptr[M] = (int32_t)(off + 1);
            continue;
        }
        if (pos < N - 1) {
            a.indices[off] = (int32_t)n;
            a.indices[off + 1] = (int32_t)(n + stride);
            put<COMPLEX>(a.e_data, a.h_data, off, a.diag_re, a.diag_im);
            put<COMPLEX>(a.e_data, a.h_data, off + 1, a.fwd_re, a.fwd_im);
            if (n == M - 1) a.indptr[M] = (int32_t)(off + 2);
        } else if (periodic) {
            a.indices[off] = (int32_t)(n - span);
            a.indices[off + 1] = (int32_t)n;
            put<COMPLEX>(a.e_data, a.h_data, off, a.wrap_re, a.wrap_im);
            put<COMPLEX>(a.e_data, a.h_data, off + 1, a.diag_re, a.diag_im);
            if (n == M - 1) a.indptr[M] = (int32_t)(off + 2);
        } else {
            a.indices[off] = (int32_t)n;
            put<COMPLEX>(a.e_data, a.h_data, off, a.diag_re, a.diag_im);
            if (n == M - 1) a.indptr[M] = (int32_t)(off + 1);
        }
    }
}

static int64_t axis_nnz(int Nx, int Ny, int axis, int bc) {
    const int64_t M = (int64_t)Nx * Ny;
    const int N = axis == 0 ? Nx : Ny;
    if (N == 1) return M;
    if (bc == FDFD_BC_BLOCH) return 2 * M;
    return axis == 0 ? 2 * M - Ny : 2 * M - Nx;
}
static int axis_complex(int Nx, int Ny, int axis, int bc) {
    const int N = axis == 0 ? Nx : Ny;
    return (N == 1 || bc == FDFD_BC_BLOCH) ? 1 : 0;
}

int yee2d_fill(int Nx, int Ny, int axis, int bc, const double vals[6], int32_t* indptr,
               int32_t* indices, void* e_data, void* h_data, cudaStream_t stream) {
    FDFD_CHECK_ARG(Nx >= 1 && Ny >= 1, "Nx, Ny must be >= 1");
    FDFD_CHECK_ARG(axis == 0 || axis == 1, "axis must be 0 (DEX) or 1 (DEY)");
    FDFD_CHECK_ARG(bc == 0 || bc == 1, "bc must be 0 (Dirichlet) or 1 (Bloch)");
    FDFD_CHECK_ARG(vals && indptr && indices && e_data, "null pointer");
    FDFD_CHECK_ARG(axis_nnz(Nx, Ny, axis, bc) < (int64_t)INT32_MAX, "nnz exceeds int32 index range");
    YeeFillArgs a;
    a.Nx = Nx; a.Ny = Ny; a.axis = axis; a.bc = bc;
    a.is_complex = axis_complex(Nx, Ny, axis, bc);
    a.diag_re = vals[0]; a.diag_im = vals[1];
    a.fwd_re = vals[2]; a.fwd_im = vals[3];
    a.wrap_re = vals[4]; a.wrap_im = vals[5];
    a.indptr = indptr; a.indices = indices; a.e_data = e_data; a.h_data = h_data;
    const int64_t M = (int64_t)Nx * Ny;
    int blocks = (int)((M + 255) / 256);
    const int cap = kNumSMs * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (a.is_complex)
        yee2d_fill_kernel<true><<<blocks, 256, 0, stream>>>(a);
    else
        yee2d_fill_kernel<false><<<blocks, 256, 0, stream>>>(a);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

// K1b — staggered-grid operators with two entries per row (rectangular forward differences between the
// staggered grids of the mode solvers, periodic z difference / forward average of the 3-D solver).
// Row (i, j, k) of the output grid: (self column, v_self) and (forward column, v_fwd), stored in ascending
// column order as scipy's coo -> csr leaves them (the periodic wrap puts the forward column first on the
// last plane).  indptr = 2 * row: closed form, no scan.
struct StaggerArgs {
    int inx, iny, inz, onx, ony, onz, axis;
    double v_self, v_fwd;
    int32_t* indptr; int32_t* indices; double* data; double* h_data;
};

__global__ void __launch_bounds__(256) stagger_fill_kernel(StaggerArgs a) {
    const int64_t rows = (int64_t)a.onx * a.ony * a.onz;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += stride) {
        a.indptr[r] = (int32_t)(2 * r);
        if (r == rows) break;
        const int i = (int)(r % a.onx);
        const int64_t q = r / a.onx;
        const int j = (int)(q % a.ony), k = (int)(q / a.ony);
        int fi = i, fj = j, fk = k;
        if (a.axis == 0) ++fi; else if (a.axis == 1) ++fj; else fk = (k + 1) % a.inz;
        const int64_t cs = i + (int64_t)a.inx * (j + (int64_t)a.iny * k);
        const int64_t cf = fi + (int64_t)a.inx * (fj + (int64_t)a.iny * fk);
        const bool self_first = cs < cf;
        a.indices[2 * r] = (int32_t)(self_first ? cs : cf);
        a.indices[2 * r + 1] = (int32_t)(self_first ? cf : cs);
        const double d0 = self_first ? a.v_self : a.v_fwd, d1 = self_first ? a.v_fwd : a.v_self;
        a.data[2 * r] = d0; a.data[2 * r + 1] = d1;
        if (a.h_data) { a.h_data[2 * r] = -d0; a.h_data[2 * r + 1] = -d1; }
    }
}

int stagger_fill(const int in_shape[3], const int out_shape[3], int axis, double v_self, double v_fwd,
                 int32_t* indptr, int32_t* indices, double* data, double* h_data, cudaStream_t stream) {
    FDFD_CHECK_ARG(in_shape && out_shape && indptr && indices && data, "null pointer");
    FDFD_CHECK_ARG(axis >= 0 && axis <= 2, "axis must be 0, 1 or 2");
    for (int d = 0; d < 3; ++d) FDFD_CHECK_ARG(in_shape[d] >= 1 && out_shape[d] >= 1, "grid extents must be >= 1");
    // every forward neighbour must exist: the input grid is one larger along a difference axis (x, y),
    // equal along the periodic axis (z) and at least as large elsewhere
    for (int d = 0; d < 3; ++d) {
        if (d == axis && axis < 2) FDFD_CHECK_ARG(in_shape[d] >= out_shape[d] + 1, "input grid must be one larger along the difference axis");
        else FDFD_CHECK_ARG(in_shape[d] >= out_shape[d], "input grid smaller than the output grid");
    }
    if (axis == 2) FDFD_CHECK_ARG(in_shape[2] == out_shape[2] && in_shape[2] >= 2, "periodic z operator needs equal grids with nz >= 2");
    const int64_t rows = (int64_t)out_shape[0] * out_shape[1] * out_shape[2];
    const int64_t cols = (int64_t)in_shape[0] * in_shape[1] * in_shape[2];
    FDFD_CHECK_ARG(2 * rows < (int64_t)INT32_MAX && cols < (int64_t)INT32_MAX, "operator exceeds the int32 index range");
    StaggerArgs a{in_shape[0], in_shape[1], in_shape[2], out_shape[0], out_shape[1], out_shape[2], axis, v_self, v_fwd,
                  indptr, indices, data, h_data};
    int blocks = (int)((rows + 1 + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    stagger_fill_kernel<<<blocks, 256, 0, stream>>>(a);
    FDFD_CUDA(cudaGetLastError());
    return FDFD_OK;
}

}  // namespace fdfd

using namespace fdfd;

// §8f-2 — coefficient arrays of the scattering operator from an object-id map: out[q] = table[ids[q]] for the
// three arrays at once (1 byte read, 48 bytes written per point).  No arithmetic on the values: the table is
// evaluated by the caller with the reference's numpy expressions, which is what keeps the device set-up
// bit-identical to Scattering_Solver_2D.py:72-153 + :176-188.
namespace fdfd { namespace {
__global__ void __launch_bounds__(256) scatter_fill_kernel(const uint8_t* __restrict__ ids, int64_t n, int ntab,
        const double2* __restrict__ tab, double2* __restrict__ ca, double2* __restrict__ cb, double2* __restrict__ cd) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += stride) {
        const int k = ids[q];
        ca[q] = tab[k]; cb[q] = tab[ntab + k]; cd[q] = tab[2 * ntab + k];
    }
}
} }

extern "C" int fdfd_scatter_fill(const uint8_t* ids, int64_t n, int ntab, const double* table_host, void* ca, void* cb,
                                 void* cd, void* stream) {
    FDFD_CHECK_ARG(ids && table_host && ca && cb && cd, "null pointer");
    FDFD_CHECK_ARG(n >= 1 && ntab >= 1 && ntab <= 256, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    double2* tab = nullptr;
    FDFD_CUDA(cudaMalloc(&tab, sizeof(double2) * 3 * ntab));
    cudaError_t e = cudaMemcpyAsync(tab, table_host, sizeof(double2) * 3 * ntab, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        int blocks = (int)((n + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        scatter_fill_kernel<<<blocks, 256, 0, st>>>(ids, n, ntab, tab, (double2*)ca, (double2*)cb, (double2*)cd);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tab);
    if (e != cudaSuccess) return cuda_fail(e, "fdfd_scatter_fill", __FILE__, __LINE__);
    return FDFD_OK;
}

extern "C" int fdfd_stagger_csr_fill(const int in_shape[3], const int out_shape[3], int axis, double v_self, double v_fwd,
                                     int32_t* indptr, int32_t* indices, double* data, double* h_data, void* stream) {
    return stagger_fill(in_shape, out_shape, axis, v_self, v_fwd, indptr, indices, data, h_data, (cudaStream_t)stream);
}

extern "C" int fdfd_stagger_csr_fill_host(const int in_shape[3], const int out_shape[3], int axis, double v_self,
                                          double v_fwd, int32_t* indptr, int32_t* indices, double* data, double* h_data) {
    FDFD_CHECK_ARG(in_shape && out_shape && indptr && indices && data, "null pointer");
    for (int d = 0; d < 3; ++d) FDFD_CHECK_ARG(out_shape[d] >= 1, "grid extents must be >= 1");
    const int64_t rows = (int64_t)out_shape[0] * out_shape[1] * out_shape[2];
    int32_t *d_ptr = nullptr, *d_idx = nullptr;
    double *d_e = nullptr, *d_h = nullptr;
    int st = FDFD_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_ptr, (rows + 1) * 4)) != cudaSuccess || (e = cudaMalloc(&d_idx, 2 * rows * 4)) != cudaSuccess ||
        (e = cudaMalloc(&d_e, 2 * rows * 8)) != cudaSuccess || (h_data && (e = cudaMalloc(&d_h, 2 * rows * 8)) != cudaSuccess))
        st = cuda_fail(e, "cudaMalloc (stagger host fill)", __FILE__, __LINE__);
    if (st == FDFD_OK) st = stagger_fill(in_shape, out_shape, axis, v_self, v_fwd, d_ptr, d_idx, d_e, d_h, 0);
    if (st == FDFD_OK) {
        if ((e = cudaMemcpy(indptr, d_ptr, (rows + 1) * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (e = cudaMemcpy(indices, d_idx, 2 * rows * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (e = cudaMemcpy(data, d_e, 2 * rows * 8, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (h_data && (e = cudaMemcpy(h_data, d_h, 2 * rows * 8, cudaMemcpyDeviceToHost)) != cudaSuccess))
            st = cuda_fail(e, "cudaMemcpy D2H (stagger host fill)", __FILE__, __LINE__);
    }
    cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_e); cudaFree(d_h);
    return st;
}

extern "C" int fdfd_yee2d_csr_nnz(int Nx, int Ny, int bcx, int bcy, int64_t* nnz_dex,
                                  int64_t* nnz_dey, int* dex_is_complex, int* dey_is_complex) {
    FDFD_CHECK_ARG(Nx >= 1 && Ny >= 1, "Nx, Ny must be >= 1");
    FDFD_CHECK_ARG((bcx == 0 || bcx == 1) && (bcy == 0 || bcy == 1), "BC must be 0 or 1");
    if (nnz_dex) *nnz_dex = axis_nnz(Nx, Ny, 0, bcx);
    if (nnz_dey) *nnz_dey = axis_nnz(Nx, Ny, 1, bcy);
    if (dex_is_complex) *dex_is_complex = axis_complex(Nx, Ny, 0, bcx);
    if (dey_is_complex) *dey_is_complex = axis_complex(Nx, Ny, 1, bcy);
    return FDFD_OK;
}

extern "C" int fdfd_yee2d_csr_fill(int Nx, int Ny, int axis, int bc, const double vals[6],
                                   int32_t* indptr, int32_t* indices, void* e_data,
                                   void* h_data, void* stream) {
    return yee2d_fill(Nx, Ny, axis, bc, vals, indptr, indices, e_data, h_data,
                      (cudaStream_t)stream);
}

extern "C" int fdfd_yee2d_csr_fill_host(int Nx, int Ny, int axis, int bc, const double vals[6],
                                        int32_t* indptr, int32_t* indices, void* e_data,
                                        void* h_data) {
    FDFD_CHECK_ARG(Nx >= 1 && Ny >= 1, "Nx, Ny must be >= 1");
    FDFD_CHECK_ARG(axis == 0 || axis == 1, "axis must be 0 (DEX) or 1 (DEY)");
    FDFD_CHECK_ARG(bc == 0 || bc == 1, "bc must be 0 (Dirichlet) or 1 (Bloch)");
    const int64_t M = (int64_t)Nx * Ny;
    const int64_t nnz = axis_nnz(Nx, Ny, axis, bc);
    const size_t esz = axis_complex(Nx, Ny, axis, bc) ? 16 : 8;
    int32_t *d_ptr = nullptr, *d_idx = nullptr;
    void *d_e = nullptr, *d_h = nullptr;
    int st = FDFD_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_ptr, (M + 1) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&d_idx, nnz * 4)) != cudaSuccess ||
        (e = cudaMalloc(&d_e, nnz * esz)) != cudaSuccess ||
        (h_data && (e = cudaMalloc(&d_h, nnz * esz)) != cudaSuccess)) {
        st = cuda_fail(e, "cudaMalloc (yee2d host fill)", __FILE__, __LINE__);
    }
    if (st == FDFD_OK) st = yee2d_fill(Nx, Ny, axis, bc, vals, d_ptr, d_idx, d_e, d_h, 0);
    if (st == FDFD_OK) {
        if ((e = cudaMemcpy(indptr, d_ptr, (M + 1) * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (e = cudaMemcpy(indices, d_idx, nnz * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (e = cudaMemcpy(e_data, d_e, nnz * esz, cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (h_data && (e = cudaMemcpy(h_data, d_h, nnz * esz, cudaMemcpyDeviceToHost)) != cudaSuccess))
            st = cuda_fail(e, "cudaMemcpy D2H (yee2d host fill)", __FILE__, __LINE__);
    }
    cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_e); cudaFree(d_h);
    return st;
}
