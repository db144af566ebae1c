// Persistent coarsest-level solver of the multigrid cycle (mg_coarse_mode = 2).
//
// Included by mg.cu inside its anonymous namespace (uses PointCoefs / point_coefs / LineView).
//
// The coarsest level (512^2 for a 4096^2 grid) is solved by a fixed number of BiCGSTAB iterations
// preconditioned with one smoothing sweep from a zero guess (mode 1: krylov.cu bicgstab_t with
// fixed_its + MGPrecond::smooth).  That is ~30 dependent kernel launches per iteration on a problem
// whose vectors (2 MB in complex64) all sit in L2: pure launch latency, ~1.6 ms of the 3.6 ms an outer
// iteration took at 4096^2.  Here the same algorithm runs in ONE kernel:
//   * one CTA per SM (co-residency guaranteed by a cooperative launch), phases separated by a
//     hand-written generation barrier (one atomic per CTA, ~1.5 us for 148 CTAs);
//   * every relaxation line is owned by one CTA for the whole solve: its LU factors, spikes and the
//     inverse of its interface system are loaded into shared memory ONCE and reused by all 2 x its
//     sweeps, so a line phase is: fetch the right-hand side, 16-step chunk recurrences out of shared
//     memory, a 64 x 64 mat-vec out of shared memory, store;
//   * the ADI order (y-lines, then x-lines on the updated values) only needs a barrier among the CTAs
//     that own lines; the second half of the x-line update (x += omega e) is applied on the fly by the
//     stencil phase that consumes it, which removes one more grid-wide phase;
//   * inner products are reduced through a per-CTA partials array that every CTA sums in the same
//     order after the barrier (deterministic, no floating-point atomics); rho = (r^, r) is obtained
//     from (r^, s) - omega (r^, t), both accumulated with (t, s) and (t, t), so the update of x, r
//     and the next search direction share one phase.
// Per iteration: 6 grid barriers + 2 line-owner barriers (the first version of this file, which used
// 1024 CTAs, cooperative-groups grid.sync() and 11 phases, was slower than the launch chain).
//
//   XA  x += alpha p^ + omega s^ ; r = s - omega t ; p = r + beta (p - omega v) ; p^0 = omega_s D^-1 p
//       outside the strip columns, strip columns -> y-line right-hand sides               | pointwise
//   L1  y-lines of p^ (line owners)            -- owner barrier --
//   L2  x-lines of p^: row residual, solve; correction left in xl.g (line owners)
//   V   p^ final = p^0 + omega e_x on the fly ; v = M p^ ; (r^, v)                            | stencil
//   S   s = r - alpha v ; s^0 as in XA                                                      | pointwise
//   L1, L2 for s^
//   T   t = M s^ ; (t, s), (t, t), (r^, s), (r^, t)                                          | stencil
#pragma once

constexpr int kCoarseThreads = 512;
constexpr int kCoarseWarps = kCoarseThreads / 32;
constexpr int kCoarseVals = 8;          // doubles per CTA in the partials array

// y-slab sharded level: every rank runs this kernel on its slab; ghost rows of the swept vector, the inner
// products and the interface values of the y-lines (which run through all ranks) travel through peer-mapped
// buffers (p2p.cuh) with one flag per message.  Messages of one kind are numbered by a sequence that every CTA
// of every rank advances in lock-step (same control flow everywhere); buffers and flags are double buffered by
// its parity — between two messages of equal parity lies at least one all-rank reduction, so a rank cannot
// overwrite data a peer has not consumed.
template <typename C>
struct CoarseShard {
    int nranks = 1, rank = 0;
    int open_s = 0, open_n = 0;            // a neighbour slab exists below / above
    C* ghost_rx = nullptr;                 // [2 parity][2 sides: 0 = row below my slab, 1 = row above][Nx]
    C* ghost_tx_s = nullptr;               // the south neighbour's ghost_rx (I fill its side 1)
    C* ghost_tx_n = nullptr;               // the north neighbour's ghost_rx (I fill its side 0)
    unsigned int *gflag_rx = nullptr, *gflag_tx_s = nullptr, *gflag_tx_n = nullptr;   // [2][2][gridDim.x]
    double* red_rx = nullptr;              // [2][nranks][kCoarseVals]
    unsigned int* rflag_rx = nullptr;      // [2][nranks]
    C* ifg_rx = nullptr;                   // [2][nranks][nl][2 P]
    unsigned int* iflag_rx = nullptr;      // [2][nranks][nl]
    // the same buffers of every rank as mapped into this process: tables in device memory (indexed by rank at
    // run time; arrays inside the kernel parameter block would be copied to local memory for that)
    double* const* red_tx = nullptr;
    unsigned int* const* rflag_tx = nullptr;
    C* const* ifg_tx = nullptr;
    unsigned int* const* iflag_tx = nullptr;
    unsigned int* seq = nullptr;           // local, persistent: sequence numbers {ghost, reduction, line}
};

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
    C *phf, *shf;                        // p^, s^ with the x-line correction applied
    LineView<C> yl, xl;                  // line sets of the level (nl may be 0)
    const C* k5;                         // matrix entries of the level, [5][n]: west, east, south, north, diagonal
    const C* relax;                      // omega / diagonal, [n]
    double* partials;                    // [gridDim.x][kCoarseVals]
    unsigned int* bar;                   // {count, generation} of the grid barrier, {count, generation} of the owners'
    int owners;                          // CTAs that own at least one line
    int spikes_in_smem;                  // the spike arrays of the lines fit in shared memory as well
    CoarseShard<C> shard;                // shard.nranks == 1: unsharded level
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ double2 cs_div(double2 a, double2 b) {     // as blas1.cu safe_div
    const double d = b.x * b.x + b.y * b.y;
    if (!(d > 0.0) || !isfinite(d)) return make_double2(0.0, 0.0);
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ double2 cs_mul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// vectors are rewritten by other SMs between phases: read them through L2 only
__device__ __forceinline__ double2 ldv(const double2* p) { return __ldcg(p); }
__device__ __forceinline__ float2 ldv(const float2* p) { return __ldcg(p); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Generation barrier among `count` CTAs.  bar[0] = arrival counter, bar[1] = generation (monotonic over
// the life of the buffer; every CTA read it at kernel start, before anyone could have advanced it).
__device__ __forceinline__ void coarse_barrier(unsigned int* bar, unsigned int count, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gen;
        // release on arrival (cumulative over the CTA's writes ordered by the bar.sync above), acquire on
        // departure: no separate full fences
        unsigned int prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == count - 1) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
            st_release_u32(&bar[1], gen);
        } else {
            while (ld_acquire_u32(&bar[1]) != gen) { }
        }
    }
    __syncthreads();
}

// block-wide sum of K doubles -> this CTA's row of the partials array (before the barrier)
template <int K>
__device__ __forceinline__ void coarse_block_reduce(double (&v)[K], double* row, double* sm /* [K][kCoarseWarps] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
        if (lane == 0) sm[k * kCoarseWarps + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kCoarseWarps; ++w) s += sm[threadIdx.x * kCoarseWarps + w];
        row[threadIdx.x] = s;
    }
}

// after the barrier: every CTA sums the K partials of all CTAs in the same order
template <int K>
__device__ __forceinline__ void coarse_grid_total(const double* partials, double (&out)[K], double* sm) {
    if (threadIdx.x < 32) {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int bidx = threadIdx.x; bidx < (int)gridDim.x; bidx += 32) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)bidx * kCoarseVals + k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // fixed butterfly order -> identical bits in every CTA
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

// sharded level: the totals of all ranks, summed in rank order on every rank (identical bits everywhere).
// CTA 0 sends this rank's totals to every peer; every CTA waits for the peers' messages itself (no extra
// grid barrier).
template <int K, typename C>
__device__ __forceinline__ void coarse_rank_total(const CoarseShard<C>& sh, double (&tot)[K], double* sm, unsigned int seq_r) {
    if (sh.nranks <= 1) return;
    const unsigned int par = seq_r & 1u;
    const int me = sh.rank, P = sh.nranks, t = threadIdx.x;
    if (t < P && t != me) {
        if (blockIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) sh.red_tx[t][((size_t)par * P + me) * kCoarseVals + k] = tot[k];
            __threadfence_system();
            st_release_sys(sh.rflag_tx[t] + (size_t)par * P + me, seq_r);
        }
        while (ld_acquire_sys(sh.rflag_rx + (size_t)par * P + t) != seq_r) { }
    }
    __syncthreads();
    if (t < K) {
        double s = 0.0;
        for (int r = 0; r < P; ++r) s += (r == me) ? tot[t] : __ldcg(&sh.red_rx[((size_t)par * P + r) * kCoarseVals + t]);
        sm[t] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = sm[k];
    __syncthreads();
}

// shared-memory image of one relaxation line: factors (and, for partitioned lines, spikes + the inverse
// of the interface system), resident for the whole solve
template <typename C>
struct LineSmem { C *fl, *fc, *ip, *sv, *sw, *rinv; };

// Lines live in shared memory with one padding element per chunk (position t -> t + t / m): the chunk
// recurrences run one thread per chunk, and without the padding all chunk heads fall into the same bank
// (m = 16 elements of 8 bytes = 128 bytes: a 32-way conflict on every access, measured 4.9 us per solve).
__host__ __device__ inline int line_pad_len(int len, int m) { return len + (m > 0 ? (len + m - 1) / m : 0) + 1; }

// `spikes`: keep the two spike arrays in shared memory too (else they are read from L2, two loads per point:
// long lines / sharded levels, where shared memory is short)
template <typename C>
__host__ __device__ inline size_t line_smem_elems(int nl, int len, int m, int Ptot, bool spikes) {
    if (nl == 0) return 0;
    const size_t pl = (size_t)line_pad_len(len, m);
    return 3 * pl + (Ptot > 1 ? (spikes ? 2 * pl : 0) + (size_t)4 * Ptot * Ptot : 0);
}

template <typename C>
__device__ LineSmem<C> line_smem_carve(C*& cursor, const LineView<C>& v, bool spikes) {
    LineSmem<C> s{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (v.nl == 0) return s;
    const int pl = line_pad_len(v.len, v.m);
    s.fl = cursor; s.fc = s.fl + pl; s.ip = s.fc + pl; cursor = s.ip + pl;
    if (v.Ptot > 1) {
        if (spikes) { s.sv = cursor; s.sw = s.sv + pl; cursor = s.sw + pl; }
        s.rinv = cursor;
        cursor = s.rinv + 4 * v.Ptot * v.Ptot;
    }
    return s;
}

template <typename C>
__device__ void line_smem_load(const LineSmem<C>& s, const LineView<C>& v, int l) {
    const int64_t nl = v.nl;
    for (int t = threadIdx.x; t < v.len; t += kCoarseThreads) {
        const int pt = t + t / v.m;
        s.fl[pt] = v.fl[t * nl + l]; s.fc[pt] = v.fc[t * nl + l]; s.ip[pt] = v.fip[t * nl + l];
        if (s.sv) { s.sv[pt] = v.sv[t * nl + l]; s.sw[pt] = v.sw[t * nl + l]; }
    }
    if (v.Ptot > 1) {
        const int nn = 4 * v.Ptot * v.Ptot;
        const C* src = v.rinv + (int64_t)l * nn;
        for (int e = threadIdx.x; e < nn; e += kCoarseThreads) s.rinv[e] = src[e];
    }
}

// -DFDFD_COARSE_TRACE: CTA 0 prints the time (ns) between the phase boundaries of the second iteration
#ifdef FDFD_COARSE_TRACE
__device__ unsigned int g_coarse_trace_count = 0;
__device__ __forceinline__ unsigned long long coarse_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ unsigned long long g_line_tr[16];
#define COARSE_MARK(k) do { if (it == 1 && blockIdx.x == 0 && threadIdx.x == 0) tr[k] = coarse_now(); } while (0)
#define COARSE_LMARK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_line_tr[k] = coarse_now(); } while (0)
#else
#define COARSE_MARK(k) do { } while (0)
#define COARSE_LMARK(k) do { } while (0)
#endif

// tridiagonal solve of the line whose right-hand side sits in sl (padded positions): chunk-local recurrences,
// interface mat-vec, spike correction; the solution replaces sl.  Same algorithm as line_fused_kernel.
template <typename C, bool SH>
__device__ void coarse_line_solve(const LineView<C>& v, const LineSmem<C>& s, C* sl, C* sg, C* su, int l,
                                  const CoarseShard<C>& sh, unsigned int seq_l, int lm = 0) {
    using R = typename real_of<C>::type;
    const int len = v.len;
    COARSE_LMARK(lm);
    if ((int)threadIdx.x < v.P) {
        const int k = threadIdx.x;
        const int t0 = k * v.m, t1 = min(len, t0 + v.m);
        if (t0 < t1) {
            const int p0 = t0 + k, cnt = t1 - t0;           // padded position of the chunk head
            // operands of four steps are fetched ahead of the dependent chain
            C y = sl[p0];
            for (int q0 = 1; q0 < cnt; q0 += 4) {
                C rs[4], cf[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = min(q0 + u, cnt - 1);
                    rs[u] = sl[p0 + q]; cf[u] = s.fl[p0 + q];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (q0 + u < cnt) { y = csub(rs[u], cmul(cf[u], y)); sl[p0 + q0 + u] = y; }
                }
            }
            C e = cmul(y, s.ip[p0 + cnt - 1]);
            sl[p0 + cnt - 1] = e;
            const C e_last = e;
            for (int q0 = cnt - 2; q0 >= 0; q0 -= 4) {
                C rs[4], cf[4], ci[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = max(q0 - u, 0);
                    rs[u] = sl[p0 + q]; cf[u] = s.fc[p0 + q]; ci[u] = s.ip[p0 + q];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (q0 - u >= 0) { e = cmul(csub(rs[u], cmul(cf[u], e)), ci[u]); sl[p0 + q0 - u] = e; }
                }
            }
            sg[2 * (v.k0 + k)] = e; sg[2 * (v.k0 + k) + 1] = e_last;
        }
    }
    __syncthreads();
    if (SH && v.nranks > 1) {
        // the line runs through every rank: exchange the interface values of the local chunks (peer memory),
        // every rank then holds all 2 Ptot of them and applies the replicated interface inverse
        const unsigned int par = seq_l & 1u;
        const int nloc = 2 * v.P, me = sh.rank, P = sh.nranks;
        for (int e = threadIdx.x; e < nloc * P; e += kCoarseThreads) {
            const int r = e / nloc, q = e - r * nloc;
            if (r != me) sh.ifg_tx[r][(((size_t)par * P + me) * v.nl + l) * nloc + q] = sg[2 * v.k0 + q];
        }
        __syncthreads();
        if ((int)threadIdx.x < P && (int)threadIdx.x != me) {
            const int r = threadIdx.x;
            __threadfence_system();
            st_release_sys(sh.iflag_tx[r] + ((size_t)par * P + me) * v.nl + l, seq_l);
            while (ld_acquire_sys(sh.iflag_rx + ((size_t)par * P + r) * v.nl + l) != seq_l) { }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nloc * P; e += kCoarseThreads) {
            const int r = e / nloc, q = e - r * nloc;
            if (r != me) sg[2 * r * v.P + q] = __ldcg(&sh.ifg_rx[(((size_t)par * P + r) * v.nl + l) * nloc + q]);
        }
        __syncthreads();
    }
    COARSE_LMARK(lm + 1);
    if (v.Ptot > 1) {
        // interface mat-vec u = R^-1 g with every thread busy: `segs` adjacent lanes share a row
        const int n = 2 * v.Ptot;
        int segs = 32;
        while (segs > 1 && segs * n > kCoarseThreads) segs >>= 1;
        const int row = threadIdx.x / segs, seg = threadIdx.x % segs;
        R ax = 0, ay = 0;
        if (row < n) {
            const C* Ri = s.rinv + row * n;
            for (int q = seg; q < n; q += segs) {
                const C rv = Ri[q], g = sg[q];
                ax += rv.x * g.x - rv.y * g.y;
                ay += rv.x * g.y + rv.y * g.x;
            }
        }
        for (int off = segs >> 1; off > 0; off >>= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, off);
            ay += __shfl_xor_sync(0xffffffffu, ay, off);
        }
        if (row < n && seg == 0) su[row] = mk<C>(ax, ay);
        __syncthreads();
        for (int t = threadIdx.x; t < len; t += kCoarseThreads) {
            const int kl = t / v.m, pt = t + kl, kg = v.k0 + kl;
            C e = sl[pt];
            if (kg > 0) e = csub(e, cmul(s.sv ? s.sv[pt] : v.sv[(int64_t)t * v.nl + l], su[2 * (kg - 1) + 1]));
            if (kg < v.Ptot - 1) e = csub(e, cmul(s.sw ? s.sw[pt] : v.sw[(int64_t)t * v.nl + l], su[2 * (kg + 1)]));
            sl[pt] = e;
        }
        __syncthreads();
    }
    COARSE_LMARK(lm + 2);
}

// M x at one point from the five values (out-of-range neighbours have zero coefficients)
template <typename C>
__device__ __forceinline__ C coarse_stencil5(const PointCoefs<C>& pc, C xc, C xw, C xe, C xs, C xn) {
    C ax = cmul(pc.dg, xc);
    ax = cfma(pc.kw, xw, ax);
    ax = cfma(pc.ke, xe, ax);
    ax = cfma(pc.ks, xs, ax);
    ax = cfma(pc.kn, xn, ax);
    return ax;
}

template <typename C>
struct CoarseCtx {
    LineSmem<C> ys, xs;
    C *sl, *sg, *su;
    double* red;
    unsigned int gen_all, gen_own;
    unsigned int seq_g, seq_r, seq_l;     // message sequences of the sharded exchanges (ghost rows, sums, lines)
};

// pointwise part of the zero-guess smoothing sweep for the value `val` at point (i, j): returns the swept
// value; for a strip column (relaxed by a y-line) the value is 0 and `gidx` >= 0 is where `val` goes in yl.g
template <typename C, bool TM>
__device__ __forceinline__ C coarse_point_relax(const CoarseArgs<C>& a, int i, int j, C val, int64_t& gidx) {
    const LineView<C>& yl = a.yl;
    gidx = -1;
    if (yl.nl > 0 && (i < yl.lo || i >= a.Nx - yl.hi)) {
        const int l = i < yl.lo ? i : yl.lo + (i - (a.Nx - yl.hi));
        gidx = (int64_t)j * yl.nl + l;
        return czero<C>();
    }
    return cmul(val, a.relax[(int64_t)j * a.Nx + i]);
}

// the level's matrix entries are computed once per hierarchy (coarse_coefs_kernel) and read back here
template <typename C>
__device__ __forceinline__ PointCoefs<C> coarse_coefs(const CoarseArgs<C>& a, int64_t q) {
    const int64_t n = (int64_t)a.Nx * a.Ny;
    PointCoefs<C> p;
    p.kw = a.k5[q]; p.ke = a.k5[n + q]; p.ks = a.k5[2 * n + q]; p.kn = a.k5[3 * n + q]; p.dg = a.k5[4 * n + q];
    return p;
}

template <typename C, bool TM>
__global__ void coarse_coefs_kernel(int Nx, int Ny, const C* ca, const C* cb, const C* cd,
        typename real_of<C>::type sx, typename real_of<C>::type sy, C shift, typename real_of<C>::type omega,
        C* k5, C* relax, SlabEdge<C> edge) {
    const int64_t n = (int64_t)Nx * Ny;
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int j = (int)(q / Nx), i = (int)(q - (int64_t)j * Nx);
    const PointCoefs<C> p = point_coefs<C, TM>(ca, cb, cd, Nx, Ny, sx, sy, shift, i, j, edge);
    k5[q] = p.kw; k5[n + q] = p.ke; k5[2 * n + q] = p.ks; k5[3 * n + q] = p.kn; k5[4 * n + q] = p.dg;
    relax[q] = cscale(omega, cdiv(mk<C>(1, 0), p.dg));
}

// points a thread handles per batch: all loads of a batch are issued before its first store, so a phase costs
// about one L2 round trip instead of one per point
template <typename C> struct CoarseBatch {
    static constexpr int U = sizeof(C) == 8 ? 4 : 2;     // pointwise phases
    static constexpr int US = sizeof(C) == 8 ? 2 : 1;    // stencil phases (more live values per point)
};

// line phases of one smoothing sweep from a zero guess.  z holds omega D^-1 rhs outside the strip columns
// and 0 inside; yl.g holds the strip-column right-hand sides.  On return (after the caller's grid barrier)
// z is final except for the rows of the x-lines, whose correction e sits in xl.g (applied by coarse_zfin).
template <typename C, bool TM, bool SH>
__device__ void coarse_line_phases(const CoarseArgs<C>& a, CoarseCtx<C>& c, C* z, const C* rhs) {
    const int l = blockIdx.x;
    if (SH && a.yl.nl > 0 && a.yl.nranks > 1) ++c.seq_l;        // (every CTA keeps the count, owners or not)
    if (l >= a.owners) return;
    COARSE_LMARK(0);
    if (a.yl.nl > 0) {
        if (l < a.yl.nl) {
            const LineView<C>& v = a.yl;
            for (int t = threadIdx.x; t < v.len; t += kCoarseThreads) c.sl[t + t / v.m] = ldv(&v.g[(int64_t)t * v.nl + l]);
            __syncthreads();
            coarse_line_solve<C, SH>(v, c.ys, c.sl, c.sg, c.su, l, a.shard, c.seq_l, 1);
            const int col = l < v.lo ? l : v.Nx - v.hi + (l - v.lo);
            for (int t = threadIdx.x; t < v.len; t += kCoarseThreads)
                z[(int64_t)t * v.Nx + col] = cscale(a.omega, c.sl[t + t / v.m]);      // the column was zero
        }
        COARSE_LMARK(4);
        if (a.xl.nl > 0) coarse_barrier(a.bar + 2, (unsigned)a.owners, c.gen_own);
        COARSE_LMARK(5);
    }
    if (a.xl.nl > 0 && l < a.xl.nl) {
        const LineView<C>& v = a.xl;
        const int j = l < v.lo ? l : v.Ny - v.hi + (l - v.lo);
        for (int i = threadIdx.x; i < v.len; i += kCoarseThreads) {
            const int64_t pn = (int64_t)j * a.Nx + i;
            const PointCoefs<C> pc = coarse_coefs<C>(a, pn);
            const C zc = ldv(&z[pn]);
            const C zw = i > 0 ? ldv(&z[pn - 1]) : czero<C>();
            const C ze = i < a.Nx - 1 ? ldv(&z[pn + 1]) : czero<C>();
            const C zs = j > 0 ? ldv(&z[pn - a.Nx]) : czero<C>();
            const C zn = j < a.Ny - 1 ? ldv(&z[pn + a.Nx]) : czero<C>();
            c.sl[i + i / v.m] = csub(ldv(&rhs[pn]), coarse_stencil5<C>(pc, zc, zw, ze, zs, zn));
        }
        __syncthreads();
        coarse_line_solve<C, false>(v, c.xs, c.sl, c.sg, c.su, l, a.shard, 0u, 6);
        for (int i = threadIdx.x; i < v.len; i += kCoarseThreads) v.g[(int64_t)i * v.nl + l] = c.sl[i + i / v.m];
        COARSE_LMARK(9);
    }
}

// value of the swept vector at (i, j): z plus the pending x-line correction of its row
template <typename C>
__device__ __forceinline__ C coarse_zfin(const CoarseArgs<C>& a, const C* z, int i, int j, int64_t q) {
    C v = ldv(&z[q]);
    const LineView<C>& xl = a.xl;
    if (xl.nl > 0 && (j < xl.lo || j >= a.Ny - xl.hi)) {
        const int l = j < xl.lo ? j : xl.lo + (j - (a.Ny - xl.hi));
        v = cadd(v, cscale(a.omega, ldv(&xl.g[(int64_t)i * xl.nl + l])));
    }
    return v;
}

// out = M zfin at the thread's points, zfin stored to zf; MODE 0: acc = (rh, out);
// MODE 1: acc = (out, s), (out, out), (rh, s), (rh, out) with s = a.r
// ghost value of the swept vector in the row below (side 0) / above (side 1) the slab
template <typename C>
__device__ __forceinline__ C coarse_ghost(const CoarseShard<C>& sh, int Nx, int side, int i, unsigned int seq_g) {
    const unsigned int par = seq_g & 1u;
    const int chunk = (Nx + (int)gridDim.x - 1) / (int)gridDim.x;
    const unsigned int* f = sh.gflag_rx + ((size_t)par * 2 + side) * gridDim.x + i / chunk;
    while (ld_acquire_sys(f) != seq_g) { }
    return __ldcg(&sh.ghost_rx[((size_t)par * 2 + side) * Nx + i]);
}

template <typename C, bool TM, int MODE, bool SH>
__device__ __forceinline__ void coarse_apply_phase(const CoarseArgs<C>& a, const C* z, C* zf, C* out, double* acc,
                                                   int tid, int stride, int n, unsigned int seq_g) {
    constexpr int U = CoarseBatch<C>::US;
    const CoarseShard<C>& sh = a.shard;
    if (SH) {
        // push this rank's edge rows of the swept vector into the neighbours' ghost rows: CTA c owns the
        // columns [c chunk, (c+1) chunk) and raises flag c on the receiver
        const unsigned int par = seq_g & 1u;
        const int chunk = (a.Nx + (int)gridDim.x - 1) / (int)gridDim.x;
        const int i0 = blockIdx.x * chunk, i1 = min(a.Nx, i0 + chunk);
        for (int i = i0 + threadIdx.x; i < i1; i += kCoarseThreads) {
            if (sh.open_n) sh.ghost_tx_n[((size_t)par * 2 + 0) * a.Nx + i] = coarse_zfin<C>(a, z, i, a.Ny - 1, (int64_t)(a.Ny - 1) * a.Nx + i);
            if (sh.open_s) sh.ghost_tx_s[((size_t)par * 2 + 1) * a.Nx + i] = coarse_zfin<C>(a, z, i, 0, i);
        }
        __syncthreads();
        if (threadIdx.x == 0 && i0 < i1) {
            __threadfence_system();
            if (sh.open_n) st_release_sys(sh.gflag_tx_n + ((size_t)par * 2 + 0) * gridDim.x + blockIdx.x, seq_g);
            if (sh.open_s) st_release_sys(sh.gflag_tx_s + ((size_t)par * 2 + 1) * gridDim.x + blockIdx.x, seq_g);
        }
    }
    for (int base = tid; base < n; base += U * stride) {
        C o[U], zc[U], rhv[U], sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = base + u * stride;
            if (q < n) {
                const int j = q / a.Nx, i = q - j * a.Nx;
                const PointCoefs<C> pc = coarse_coefs<C>(a, q);
                zc[u] = coarse_zfin<C>(a, z, i, j, q);
                const C zw = i > 0 ? coarse_zfin<C>(a, z, i - 1, j, q - 1) : czero<C>();
                const C ze = i < a.Nx - 1 ? coarse_zfin<C>(a, z, i + 1, j, q + 1) : czero<C>();
                const C zs = j > 0 ? coarse_zfin<C>(a, z, i, j - 1, q - a.Nx)
                                   : (SH && sh.open_s ? coarse_ghost<C>(sh, a.Nx, 0, i, seq_g) : czero<C>());
                const C zn = j < a.Ny - 1 ? coarse_zfin<C>(a, z, i, j + 1, q + a.Nx)
                                          : (SH && sh.open_n ? coarse_ghost<C>(sh, a.Nx, 1, i, seq_g) : czero<C>());
                o[u] = coarse_stencil5<C>(pc, zc[u], zw, ze, zs, zn);
                rhv[u] = ldv(&a.rh[q]);
                if (MODE == 1) sv[u] = ldv(&a.r[q]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = base + u * stride;
            if (q < n) {
                zf[q] = zc[u];
                out[q] = o[u];
                if (MODE == 0) {
                    acc[0] += (double)rhv[u].x * o[u].x + (double)rhv[u].y * o[u].y;        // conj(r^) * v
                    acc[1] += (double)rhv[u].x * o[u].y - (double)rhv[u].y * o[u].x;
                } else {
                    acc[0] += (double)o[u].x * sv[u].x + (double)o[u].y * sv[u].y;          // conj(t) * s
                    acc[1] += (double)o[u].x * sv[u].y - (double)o[u].y * sv[u].x;
                    acc[2] += (double)o[u].x * o[u].x + (double)o[u].y * o[u].y;            // (t, t)
                    acc[3] += (double)rhv[u].x * sv[u].x + (double)rhv[u].y * sv[u].y;      // conj(r^) * s
                    acc[4] += (double)rhv[u].x * sv[u].y - (double)rhv[u].y * sv[u].x;
                    acc[5] += (double)rhv[u].x * o[u].x + (double)rhv[u].y * o[u].y;        // conj(r^) * t
                    acc[6] += (double)rhv[u].x * o[u].y - (double)rhv[u].y * o[u].x;
                }
            }
        }
    }
}

template <typename C, bool TM, bool SH>
__global__ void __launch_bounds__(kCoarseThreads, 1) coarse_bicgstab_kernel(CoarseArgs<C> a) {
#ifdef FDFD_COARSE_TRACE
    unsigned long long tr[16];
#endif
    using R = typename real_of<C>::type;
    extern __shared__ __align__(16) unsigned char coarse_smem[];
    __shared__ C sg[2 * kMaxChunks], su[2 * kMaxChunks];
    __shared__ double red[kCoarseVals * kCoarseWarps];
    CoarseCtx<C> c;
    C* cursor = reinterpret_cast<C*>(coarse_smem);
    c.ys = line_smem_carve<C>(cursor, a.yl, a.spikes_in_smem != 0);
    c.xs = line_smem_carve<C>(cursor, a.xl, a.spikes_in_smem != 0);
    c.sl = cursor;
    c.sg = sg; c.su = su; c.red = red;
    if (threadIdx.x == 0) { sg[0] = czero<C>(); }
    c.gen_all = ld_acquire_u32(&a.bar[1]);
    c.gen_own = ld_acquire_u32(&a.bar[3]);
    c.seq_g = c.seq_r = c.seq_l = 0;
    if (SH) { c.seq_g = a.shard.seq[0]; c.seq_r = a.shard.seq[1]; c.seq_l = a.shard.seq[2]; }
    const int n = a.Nx * a.Ny;                       // < 2^31 (checked on the host)
    const int tid = blockIdx.x * kCoarseThreads + threadIdx.x;
    const int stride = gridDim.x * kCoarseThreads;
    double* prow = a.partials + (size_t)blockIdx.x * kCoarseVals;
    const unsigned int nblk = gridDim.x;

    // the lines this CTA owns stay in shared memory for the whole solve
    if ((int)blockIdx.x < a.yl.nl) line_smem_load<C>(c.ys, a.yl, blockIdx.x);
    if ((int)blockIdx.x < a.xl.nl) line_smem_load<C>(c.xs, a.xl, blockIdx.x);

    // start: x = 0, r = r^ = p = b, rho = (b, b); first search direction swept
    constexpr int U = CoarseBatch<C>::U;
    double acc2[2] = {0.0, 0.0};
    for (int base = tid; base < n; base += U * stride) {
        C bv[U], zr[U];
        int64_t gi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = base + u * stride;
            if (q < n) {
                const int j = q / a.Nx, i = q - j * a.Nx;
                bv[u] = a.b[q];
                zr[u] = coarse_point_relax<C, TM>(a, i, j, bv[u], gi[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = base + u * stride;
            if (q < n) {
                a.r[q] = bv[u]; a.rh[q] = bv[u]; a.p[q] = bv[u]; a.x[q] = czero<C>();
                a.ph[q] = zr[u];
                if (gi[u] >= 0) a.yl.g[gi[u]] = bv[u];
                acc2[0] += (double)bv[u].x * bv[u].x + (double)bv[u].y * bv[u].y;
            }
        }
    }
    coarse_block_reduce<2>(acc2, prow, red);
    coarse_barrier(a.bar, nblk, c.gen_all);
    double tot2[2];
    coarse_grid_total<2>(a.partials, tot2, red);
    if (SH) coarse_rank_total<2, C>(a.shard, tot2, red, ++c.seq_r);
    double2 rho = make_double2(tot2[0], 0.0);
    double2 alpha = make_double2(1.0, 0.0), omega = make_double2(1.0, 0.0);

    for (int it = 0; it < a.its; ++it) {
        // (p^0 and the y-line right-hand sides of this iteration were written by the previous phase)
        COARSE_MARK(0);
        coarse_line_phases<C, TM, SH>(a, c, a.ph, a.p);
        COARSE_MARK(1);
        coarse_barrier(a.bar, nblk, c.gen_all);
        COARSE_MARK(2);
        // V: v = M p^, (r^, v)
        double accv[2] = {0.0, 0.0};
        coarse_apply_phase<C, TM, 0, SH>(a, a.ph, a.phf, a.v, accv, tid, stride, n, ++c.seq_g);
        COARSE_MARK(3);
        coarse_block_reduce<2>(accv, prow, red);
        coarse_barrier(a.bar, nblk, c.gen_all);
        COARSE_MARK(4);
        coarse_grid_total<2>(a.partials, tot2, red);
        if (SH) coarse_rank_total<2, C>(a.shard, tot2, red, ++c.seq_r);
        COARSE_MARK(5);
        alpha = cs_div(rho, make_double2(tot2[0], tot2[1]));
        // S: s = r - alpha v (in r), s^0 pointwise
        const C alphaC = mk<C>((R)alpha.x, (R)alpha.y);
        for (int base = tid; base < n; base += U * stride) {
            C sv[U], zr[U];
            int64_t gi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + u * stride;
                if (q < n) {
                    const int j = q / a.Nx, i = q - j * a.Nx;
                    sv[u] = csub(ldv(&a.r[q]), cmul(alphaC, ldv(&a.v[q])));
                    zr[u] = coarse_point_relax<C, TM>(a, i, j, sv[u], gi[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + u * stride;
                if (q < n) {
                    a.r[q] = sv[u];
                    a.sh[q] = zr[u];
                    if (gi[u] >= 0) a.yl.g[gi[u]] = sv[u];
                }
            }
        }
        COARSE_MARK(6);
        coarse_barrier(a.bar, nblk, c.gen_all);
        COARSE_MARK(7);
        coarse_line_phases<C, TM, SH>(a, c, a.sh, a.r);
        COARSE_MARK(8);
        coarse_barrier(a.bar, nblk, c.gen_all);
        COARSE_MARK(9);
        // T: t = M s^, (t, s), (t, t), (r^, s), (r^, t)
        double acct[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        coarse_apply_phase<C, TM, 1, SH>(a, a.sh, a.shf, a.t, acct, tid, stride, n, ++c.seq_g);
        COARSE_MARK(10);
        coarse_block_reduce<8>(acct, prow, red);
        coarse_barrier(a.bar, nblk, c.gen_all);
        COARSE_MARK(11);
        double tot8[8];
        coarse_grid_total<8>(a.partials, tot8, red);
        if (SH) coarse_rank_total<8, C>(a.shard, tot8, red, ++c.seq_r);
        COARSE_MARK(12);
        omega = cs_div(make_double2(tot8[0], tot8[1]), make_double2(tot8[2], 0.0));
        // rho_new = (r^, s - omega t)
        const double2 rt = cs_mul(omega, make_double2(tot8[5], tot8[6]));
        const double2 rho_new = make_double2(tot8[3] - rt.x, tot8[4] - rt.y);
        const double2 beta = cs_mul(cs_div(rho_new, rho), cs_div(alpha, omega));
        const double2 nbo = cs_mul(beta, omega);
        rho = rho_new;
        // XA: x += alpha p^ + omega s^ ; r = s - omega t ; p = r + beta (p - omega v) ; p^0
        const C omegaC = mk<C>((R)omega.x, (R)omega.y);
        const C betaC = mk<C>((R)beta.x, (R)beta.y);
        const C nboC = mk<C>((R)(-nbo.x), (R)(-nbo.y));
        const bool more = it + 1 < a.its;
        for (int base = tid; base < n; base += U * stride) {
            // every array read here was last written by this same thread (same q mapping)
            C xv[U], rv[U], pv[U], zr[U];
            int64_t gi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + u * stride;
                if (q < n) {
                    xv[u] = cfma(omegaC, ldv(&a.shf[q]), cfma(alphaC, ldv(&a.phf[q]), ldv(&a.x[q])));
                    if (more) {
                        const int j = q / a.Nx, i = q - j * a.Nx;
                        rv[u] = csub(ldv(&a.r[q]), cmul(omegaC, ldv(&a.t[q])));
                        pv[u] = cfma(nboC, ldv(&a.v[q]), cfma(betaC, ldv(&a.p[q]), rv[u]));
                        zr[u] = coarse_point_relax<C, TM>(a, i, j, pv[u], gi[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + u * stride;
                if (q < n) {
                    a.x[q] = xv[u];
                    if (more) {
                        a.r[q] = rv[u]; a.p[q] = pv[u]; a.ph[q] = zr[u];
                        if (gi[u] >= 0) a.yl.g[gi[u]] = pv[u];
                    }
                }
            }
        }
        COARSE_MARK(13);
        if (more) coarse_barrier(a.bar, nblk, c.gen_all);
        COARSE_MARK(14);
#ifdef FDFD_COARSE_TRACE
        if (it == 1 && blockIdx.x == 0 && threadIdx.x == 0 && atomicAdd(&g_coarse_trace_count, 1u) < 4u)
            printf("[coarse trace ns] lines %llu | bar %llu | V %llu | red+bar %llu | total %llu | S %llu | bar %llu | lines %llu | bar %llu | T %llu | red+bar %llu | total %llu | XA %llu | bar %llu || iteration %llu\n",
                   tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[7] - tr[6],
                   tr[8] - tr[7], tr[9] - tr[8], tr[10] - tr[9], tr[11] - tr[10], tr[12] - tr[11], tr[13] - tr[12],
                   tr[14] - tr[13], tr[14] - tr[0]);
        if (it == 1 && blockIdx.x == 0 && threadIdx.x == 0)
            printf("[coarse line ns] y: load %llu local %llu interface+spike %llu store %llu | owner bar %llu | x: resid %llu local %llu interface+spike %llu store %llu\n",
                   g_line_tr[1] - g_line_tr[0], g_line_tr[2] - g_line_tr[1], g_line_tr[3] - g_line_tr[2], g_line_tr[4] - g_line_tr[3],
                   g_line_tr[5] - g_line_tr[4], g_line_tr[6] - g_line_tr[5], g_line_tr[7] - g_line_tr[6], g_line_tr[8] - g_line_tr[7],
                   g_line_tr[9] - g_line_tr[8]);
#endif
    }
    // (every CTA read the sequence numbers before the first barrier of this launch)
    if (SH && blockIdx.x == 0 && threadIdx.x == 0) {
        a.shard.seq[0] = c.seq_g; a.shard.seq[1] = c.seq_r; a.shard.seq[2] = c.seq_l;
    }
}
