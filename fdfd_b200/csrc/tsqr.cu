// K7 — tall-skinny QR for the refined Ritz extraction (refined_shift_invert_arnoldi.py:85: the reference takes
// the SVD of the n x m matrix (A - theta B) V; its right singular vectors are those of the m x m factor R).
//
// W = AV - theta * BV is never formed: AV and BV are stored as m rows of length n (one row per basis vector =
// one COLUMN of W), and every CTA builds its panel of W while loading it.  One CTA factors one panel of `rows`
// rows by Householder reflections in shared memory (column-major, one warp per trailing column, lanes over the
// rows) and writes the m x m triangle; the triangles, stacked, are the next level's tall matrix.  The number
// of levels is log_{rows/m}(n / m): 4 for n = 119 040, m = 40.  Householder (not CholeskyQR) because the
// matrices of interest are singular to working precision by construction: sigma_min / sigma_max is the
// residual of the Ritz pair (1e-10 ... 1e-13), which a Gram matrix cannot represent.
#include "common.cuh"

namespace fdfd {

namespace {

constexpr int kTsqrThreads = 256;

// panel rows for a given width: as many as fit in 200 KB of shared memory, multiple of 32 (>= 2 m for m <= 64,
// which guarantees ceil(len / rows) * m < len on every level)
inline int tsqr_rows(int m) {
    int r = (int)(((size_t)200 << 10) / ((size_t)m * sizeof(double2)));
    r = (r / 32) * 32;
    if (r > 1024) r = 1024;
    return r;
}

// in0 / in1: m rows of length n with leading dimension ld (in1 may be null); panel p covers rows [p*rows, ...).
// out: m rows of length npanels * m with leading dimension ldo; out[c][p*m + i] = R_p[i][c].
__global__ void __launch_bounds__(kTsqrThreads, 1)
tsqr_panel_kernel(int m, int64_t n, int rows, const double2* __restrict__ in0, const double2* __restrict__ in1,
                  int64_t ld, double2 theta, double2* __restrict__ out, int64_t ldo) {
    extern __shared__ double2 A[];                 // [m][rows], column c at A + c * rows
    __shared__ double red[kTsqrThreads / 32];
    __shared__ double2 head;                       // v_k of the current reflector
    __shared__ double scale;                       // 2 / (v^H v)
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nwarp = kTsqrThreads / 32;
    const int64_t r0 = (int64_t)blockIdx.x * rows;
    for (int c = warp; c < m; c += nwarp) {
        const double2* a0 = in0 + (size_t)c * ld;
        const double2* a1 = in1 ? in1 + (size_t)c * ld : nullptr;
        for (int r = lane; r < rows; r += 32) {
            double2 v = make_double2(0, 0);
            if (r0 + r < n) {
                v = a0[r0 + r];
                if (a1) v = csub(v, cmul(theta, a1[r0 + r]));
            }
            A[(size_t)c * rows + r] = v;
        }
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        double2* x = A + (size_t)k * rows;
        double s = 0;
        for (int r = k + tid; r < rows; r += kTsqrThreads) s += x[r].x * x[r].x + x[r].y * x[r].y;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            double sig2 = 0;
            for (int w = 0; w < nwarp; ++w) sig2 += red[w];
            const double2 x0 = x[k];
            const double a0 = sqrt(x0.x * x0.x + x0.y * x0.y), nx = sqrt(sig2);
            const double2 ph = a0 > 0 ? make_double2(x0.x / a0, x0.y / a0) : make_double2(1, 0);
            const double2 alpha = make_double2(-ph.x * nx, -ph.y * nx);      // H x = alpha e_k
            head = csub(x0, alpha);
            const double vv = 2.0 * (sig2 + a0 * nx);                         // v^H v
            scale = vv > 0 ? 2.0 / vv : 0.0;
            x[k] = alpha;                                                     // R[k][k]; v_k lives in `head`
        }
        __syncthreads();
        const double2 vk = head;
        const double sc = scale;
        for (int j = k + 1 + warp; j < m; j += nwarp) {
            double2* a = A + (size_t)j * rows;
            double2 w = make_double2(0, 0);                                   // v^H a_j
            for (int r = k + lane; r < rows; r += 32) {
                const double2 vr = r == k ? vk : x[r];
                w = cfma(cconj(vr), a[r], w);
            }
            for (int o = 16; o; o >>= 1) {
                w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
                w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
            }
            w = make_double2(-sc * w.x, -sc * w.y);
            for (int r = k + lane; r < rows; r += 32) {
                const double2 vr = r == k ? vk : x[r];
                a[r] = cfma(w, vr, a[r]);
            }
        }
        __syncthreads();
    }
    for (int c = warp; c < m; c += nwarp)
        for (int i = lane; i < m; i += 32)
            out[(size_t)c * ldo + (size_t)blockIdx.x * m + i] = i <= c ? A[(size_t)c * rows + i] : make_double2(0, 0);
}

}  // namespace

// R (m x m, stored column by column, device) of the QR factorisation of W = AV^T - theta BV^T (n x m).
int tsqr_r(int m, int64_t n, const double2* AV, const double2* BV, int64_t ld, double2 theta, double2* R,
           cudaStream_t st) {
    const int rows = tsqr_rows(m);
    const size_t smem = (size_t)m * rows * sizeof(double2);
    static bool configured = false;
    if (!configured) {
        FDFD_CUDA(cudaFuncSetAttribute(tsqr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 205 << 10));
        configured = true;
    }
    int64_t np = (n + rows - 1) / rows;
    double2* buf[2] = {nullptr, nullptr};
    const size_t cap = (size_t)m * (size_t)np * m * sizeof(double2);
    FDFD_CUDA(dev_alloc(&buf[0], cap));
    cudaError_t e = dev_alloc(&buf[1], cap);
    if (e != cudaSuccess) { dev_free(buf[0]); return cuda_fail(e, "dev_alloc(tsqr)", __FILE__, __LINE__); }
    const double2* in0 = AV;
    const double2* in1 = BV;
    int64_t len = n, ldin = ld;
    int which = 0, status = FDFD_OK;
    for (;;) {
        np = (len + rows - 1) / rows;
        const int64_t ldo = np * m;
        tsqr_panel_kernel<<<(unsigned)np, kTsqrThreads, smem, st>>>(m, len, rows, in0, in1, ldin, theta, buf[which], ldo);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) { status = cuda_fail(le, "tsqr_panel_kernel", __FILE__, __LINE__); break; }
        if (np == 1) break;
        if (ldo >= len) { set_error("tsqr: no progress (m = %d, %d rows per panel)", m, rows); status = FDFD_ERR_ARG; break; }
        in0 = buf[which]; in1 = nullptr; len = ldo; ldin = ldo;
        which ^= 1;
    }
    // buf[which] holds out[c][i] = R[i][c] with leading dimension m, i.e. R column by column
    if (status == FDFD_OK) {
        cudaError_t ce = cudaMemcpyAsync(R, buf[which], (size_t)m * m * sizeof(double2), cudaMemcpyDeviceToDevice, st);
        if (ce != cudaSuccess) status = cuda_fail(ce, "cudaMemcpyAsync(tsqr R)", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(st);          // the work buffers go back to the cache: nothing may still read them
    dev_free(buf[0]); dev_free(buf[1]);
    return status;
}

}  // namespace fdfd

using namespace fdfd;

// R^T of the QR factorisation of W = (AV - theta * BV)^T: AV, BV device complex128, m rows of length n (leading
// dimension ld), BV may be NULL (theta ignored).  Rt_dev (device, m*m complex128) receives Rt[c][i] = R[i][c],
// i.e. R in column-major order.  Replaces the n x m SVD of refined_shift_invert_arnoldi.py:85 (the right singular
// vectors of W are those of R).
extern "C" int fdfd_tsqr_r(int m, int64_t n, const void* AV, const void* BV, int64_t ld, double theta_re,
                           double theta_im, void* Rt_dev, void* stream) {
    FDFD_CHECK_ARG(AV && Rt_dev, "null pointer");
    FDFD_CHECK_ARG(m >= 1 && m <= 64, "1 <= m <= 64 columns");   // panels of >= 2 m rows: every level shrinks the stack
    FDFD_CHECK_ARG(n >= 1 && ld >= n, "bad shape");
    if (fdfd_device_count() <= 0) { set_error("no CUDA device visible: fdfd_b200 has no CPU path"); return FDFD_ERR_CUDA; }
    return tsqr_r(m, n, (const double2*)AV, (const double2*)BV, ld, make_double2(theta_re, theta_im), (double2*)Rt_dev,
                  (cudaStream_t)stream);
}
