// Shared device/host helpers for libfdfd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fdfd_b200.h"

namespace fdfd {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FDFD_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return ::fdfd::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define FDFD_CHECK_ARG(cond, msg)                                              \
    do {                                                                       \
        if (!(cond)) { ::fdfd::set_error("invalid argument: %s", msg); return FDFD_ERR_ARG; } \
    } while (0)

#define FDFD_TRY(call)                                                         \
    do { int _s = (call); if (_s != FDFD_OK) return _s; } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// cached device allocations for small work buffers (devpool.cu): same contract as cudaMalloc / cudaFree,
// the contents of a recycled block are undefined
cudaError_t dev_alloc(void** out, size_t bytes);
template <typename T> inline cudaError_t dev_alloc(T** out, size_t bytes) { return dev_alloc((void**)out, bytes); }
void dev_free(void* p);
size_t dev_trim();

// ---- complex arithmetic on (re, im) pairs -------------------------------------------------
template <typename T> struct cplx_of;
template <> struct cplx_of<double> { using type = double2; };
template <> struct cplx_of<float> { using type = float2; };
template <typename T> using cplx = typename cplx_of<T>::type;

template <typename C> struct real_of;
template <> struct real_of<double2> { using type = double; };
template <> struct real_of<float2> { using type = float; };

template <typename C> __host__ __device__ __forceinline__ C mk(typename real_of<C>::type a, typename real_of<C>::type b) {
    C r; r.x = a; r.y = b; return r;
}
template <typename C> __host__ __device__ __forceinline__ C czero() { return mk<C>(0, 0); }
template <typename C> __host__ __device__ __forceinline__ C cadd(C a, C b) { return mk<C>(a.x + b.x, a.y + b.y); }
template <typename C> __host__ __device__ __forceinline__ C csub(C a, C b) { return mk<C>(a.x - b.x, a.y - b.y); }
template <typename C> __host__ __device__ __forceinline__ C cmul(C a, C b) {
    return mk<C>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// conj(a) * b
template <typename C> __host__ __device__ __forceinline__ C cmulc(C a, C b) {
    return mk<C>(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
template <typename C> __host__ __device__ __forceinline__ C cconj(C a) { return mk<C>(a.x, -a.y); }
template <typename C> __host__ __device__ __forceinline__ C cscale(typename real_of<C>::type s, C a) { return mk<C>(s * a.x, s * a.y); }
template <typename C> __host__ __device__ __forceinline__ C cneg(C a) { return mk<C>(-a.x, -a.y); }
// a + b*c
template <typename C> __host__ __device__ __forceinline__ C cfma(C b, C c, C a) {
    return mk<C>(a.x + (b.x * c.x - b.y * c.y), a.y + (b.x * c.y + b.y * c.x));
}
template <typename C> __host__ __device__ __forceinline__ C cdiv(C a, C b) {
    // one reciprocal instead of two divisions (a double-precision divide is a ~30-instruction
    // Newton sequence; the smoother does one complex division per grid point)
    const typename real_of<C>::type inv = (typename real_of<C>::type)1 / (b.x * b.x + b.y * b.y);
    return mk<C>((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}

// 16-byte streaming loads/stores: vectors and coefficient arrays are touched once per apply,
// keep them out of L1's way where that helps (x rows are re-used and go through the normal path).
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(double2* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y));
}
__device__ __forceinline__ void st_stream(float2* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y));
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace fdfd
