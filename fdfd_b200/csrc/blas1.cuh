// K4 — fused BLAS-1 for the Krylov solvers: linear combinations with device-resident scalar
// coefficients and in-kernel inner products (warp shuffles -> shared memory -> last-block
// finalisation, deterministic for a fixed launch shape).  All reductions accumulate in double.
#pragma once
#include "common.cuh"

namespace fdfd {

struct Comm;

// Device scalar slots (double2 each) + reduction scratch, one per solver instance.
struct Scalars {
    double2* slot = nullptr;     // kSlots complex scalars
    double* partials = nullptr;  // kMaxBlocks * kMaxDots * 2 doubles
    unsigned int* counter = nullptr;
    static constexpr int kSlots = 256;
    static constexpr int kMaxBlocks = kNumSMs * 8;
    static constexpr int kMaxDots = 8;
    int alloc();
    void release();
};

// out = c0*x0 + c1*x1 + c2*x2   (ck read from device slots; a null xk drops the term;
// neg_k negates ck).  Optionally also accumulates, in the same pass,
//   dot_u  -> slot[dot_u_slot]  = sum conj(u) * out
//   dot_oo -> slot[dot_oo_slot] = sum |out|^2          (stored as (value, 0))
// A negative slot index disables the respective reduction.  Results are LOCAL sums; callers
// owning a communicator follow up with allreduce_slots().
template <typename C>
int lincomb(int64_t n, C* out, const double2* c0, const C* x0, const double2* c1, const C* x1,
            const double2* c2, const C* x2, int neg_mask, const C* u, int dot_u_slot,
            int dot_oo_slot, Scalars& sc, cudaStream_t stream);

// slot[s0 + k] = sum conj(x_k) * y_k  for k < ndots (<= kMaxDots); pointers given on the host.
template <typename C>
int dots(int64_t n, int ndots, const C* const* xs, const C* const* ys, int s0, Scalars& sc,
         cudaStream_t stream);

// Multi-vector kernels for (F)GMRES: V holds vectors of length n at stride ldv.
//   multi_dot:  slot[s0 + i] = sum conj(V_i) * w , i < k
//   multi_axpy: w += sign * sum_i slot[s0 + i] * V_i  ; optional ||w||^2 -> slot[nrm_slot]
// The basis may be stored in lower precision than w (CV = float2, C = double2): arithmetic in C.
template <typename CV, typename C>
int multi_dot(int64_t n, int k, const CV* V, int64_t ldv, const C* w, int s0, Scalars& sc,
              cudaStream_t stream);
template <typename CV, typename C>
int multi_axpy(int64_t n, int k, const CV* V, int64_t ldv, C* w, int s0, double sign,
               int nrm_slot, Scalars& sc, cudaStream_t stream);

// out = x * (1 / sqrt(re(slot[nrm2_slot])))        (normalise by a squared norm on device)
template <typename C>
int scale_inv_sqrt(int64_t n, C* out, const C* x, int nrm2_slot, Scalars& sc, cudaStream_t stream);

// same, also writing a complex64 copy (compressed Krylov basis) and optionally a second complex64 copy (the
// input buffer of a complex64 preconditioner); `out` may be null when only the complex64 copies are needed
int scale_inv_sqrt_dual(int64_t n, double2* out, float2* out_single, float2* out_single2, const double2* x,
                        int nrm2_slot, Scalars& sc, cudaStream_t stream);

// out = x .* d (elementwise complex product; Jacobi preconditioner with d = 1/diag(A))
template <typename C>
int pointwise_mul(int64_t n, C* out, const C* x, const C* d, cudaStream_t stream);

// out = (Cout) in, elementwise precision conversion (mixed-precision preconditioning)
template <typename Cin, typename Cout>
int convert(int64_t n, const Cin* in, Cout* out, cudaStream_t stream);

// sum slots [s0, s0+count) across ranks (no-op without communicator)
int allreduce_slots(Comm* comm, Scalars& sc, int s0, int count, cudaStream_t stream);

// tiny device-side scalar programs (one thread) so the host never waits for a coefficient
enum ScalarOp {
    SOP_SET = 0,        // slot[a] = (p, q)
    SOP_DIV = 1,        // slot[a] = slot[b] / slot[c]
    SOP_BICG_BETA = 2,  // slot[a] = (slot[b]/slot[c]) * (slot[d]/slot[e]) ; slot[a+1] = -slot[a]*slot[e]
    SOP_COPY = 3,       // slot[a] = slot[b]
    SOP_NEG = 4,        // slot[a] = -slot[b]
};
int scalar_op(int op, int a, int b, int c, int d, int e, double p, double q, Scalars& sc,
              cudaStream_t stream);

}  // namespace fdfd
