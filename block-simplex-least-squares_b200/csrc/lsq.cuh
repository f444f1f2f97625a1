// lsq.cuh -- the sparse least-squares half of the hot path on sm_100a: the SpMV pair
// r = A x - b, g = A^T r with the solver epilogues fused in, plus the vector and
// per-block kernels the solver drivers are made of.
//
// Replaces, in the reference:
//   sparse_least_squares_obj            python/algorithm_utils.py:88-94   (two scipy CSR mat-vecs + dot)
//   BATCH.solve_BB update / deltas      python/BATCH.py:87-102
//   BATCH.solve_MD / mirror_descent     python/BATCH.py:238-241, python/mirror_descent.py:39-47
//   normalization                       python/algorithm_utils.py:175-179
//   N z, N^T v (z-space)                python/main.py:53-54, python/bsls_utils.py:139-162
//   x2z_c / z2x_c                       python/c_extensions/c_extensions.pyx:195-248
//
// Everything here is HBM/L2-bound: A is streamed once per product (12 B per non-zero, or
// 4 B when the values are implicit ones), the gathered vector is served by L2.  No tensor
// cores.  Two SpMV shapes, picked per matrix side from its mean row length:
//   * STREAM (short rows: A^T has one row per route, a handful of links each): a CTA pulls
//     the contiguous non-zero range of 256 rows through shared memory with coalesced loads,
//     forming the products on the way; each thread then sums its own row LEFT TO RIGHT --
//     the order scipy's csr_matvec uses, so the result is bit-identical to the reference.
//   * VECTOR (long rows: A has one row per link, hundreds of routes each): LANES lanes per
//     row, strided, four gathers in flight per lane, xor-shuffle tree.
// Reductions (objective value, BB dot products) are deterministic: per-CTA partials, the
// last CTA to finish adds them in a fixed order.
#pragma once
#include "common.cuh"

namespace bsls {

constexpr int kRedMaxGrid = 148 * 8;  // upper bound on any reducing kernel's grid
constexpr int kRedSlots = 8;          // accumulators per CTA (sums first, maxima after)

struct RedCtx {
    double *partials;   // kRedMaxGrid * kRedSlots
    unsigned *ticket;   // zero between launches
    double *out;        // device scalars written by the finishing CTA
};

// ---- deterministic grid reduction ------------------------------------------------------
// acc[0..NSUM) are added, acc[NSUM..NSUM+NMAX) are maximised.  FIN maps slot k and the grid
// total to the value stored in red.out[FIN::slot(k)].
template <int NSUM, int NMAX, int THREADS, class FIN>
__device__ __forceinline__ void grid_reduce(double (&acc)[NSUM + NMAX], const RedCtx &red) {
    constexpr int N = NSUM + NMAX;
    static_assert(N <= kRedSlots, "too many accumulators");
    __shared__ double s_w[THREADS / 32][N];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = (k < NSUM) ? v + u : fmax(v, u);
        }
        if (lane == 0) s_w[wid][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double v = s_w[0][k];
            for (int w = 1; w < THREADS / 32; ++w) v = (k < NSUM) ? v + s_w[w][k] : fmax(v, s_w[w][k]);
            red.partials[(size_t)blockIdx.x * kRedSlots + k] = v;
        }
        __threadfence();
        const unsigned t = atomicAdd(red.ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the finishing CTA: fixed assignment of partials to threads, then the same tree as above
    double tot[N];
#pragma unroll
    for (int k = 0; k < N; ++k) tot[k] = 0.0;  // maxima are of absolute values (>= 0)
    for (int b = tid; b < (int)gridDim.x; b += THREADS) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double u = __ldcg(&red.partials[(size_t)b * kRedSlots + k]);
            tot[k] = (k < NSUM) ? tot[k] + u : fmax(tot[k], u);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double v = tot[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = (k < NSUM) ? v + u : fmax(v, u);
        }
        if (lane == 0) s_w[wid][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double v = s_w[0][k];
            for (int w = 1; w < THREADS / 32; ++w) v = (k < NSUM) ? v + s_w[w][k] : fmax(v, s_w[w][k]);
            FIN::store(red.out, k, v);
        }
        *red.ticket = 0u;
    }
}

// ---- device scalar slots (bsls_lsq::d_scal) -------------------------------------------------
enum Scal {
    kScalF = 0,     // 0.5 <r, r>
    kScalSxy = 1,   // <x_new - x, g_new - g>
    kScalSyy = 2,   // <g_new - g, g_new - g>
    kScalGd = 3,    // <g, x_new - x>
    kScalGnn = 4,   // <g_new, g_new>
    kScalStep = 5,  // max |x_new - x|
    kScalYgo = 6,   // <g_new - g, g>  (L-BFGS; shares the first generic slot, which no solver loop uses)
    kScalDot0 = 6,  // generic dot products: 6..9
    kScalMax0 = 10, // generic maximum
    kScalRR = 11,   // <r, r> (not halved)
    kScalRdr = 12,  // <r_old, r_new - r_old>      (line search along the step, see decide_kernel)
    kScalDrdr = 13, // <r_new - r_old, r_new - r_old>
    kScalCount = 16
};

// ---- solver state on the device (device-resident step logic) ------------------------------------
// The BATCH loops of the reference branch on scalars every iteration (python/BATCH.py:33-47,84-102,
// python/algorithm_utils.py:113-137,158-172).  Here those scalars never leave the GPU inside the loop: the reducing
// kernels leave them in the scalar block, decide_kernel (one thread) takes the line-search / step / stopping decisions
// and the kernels of the next iteration read what they need (the step `t`, the `done` flag) from this struct.
struct DevState {
    double f, f_old;       // objective at the current / previous iterate
    double t;              // step of the NEXT trial point (BB: <dx,dg>/<dg,dg>; else 1/(min_eig i + 1))
    double tau;            // line-search factor of the LAST iteration: x_new = x + tau (x_trial - x); 1 = trial accepted as is
    double sxy, syy;       // <dx, dg>, <dg, dg> of the last accepted step
    double stop_value;
    double change;         // mirror descent: max |x_new - x| of the last step
    int i;                 // the reference's iteration counter
    int done;              // 0 = running, else the stop code (1 max_iter, 2 f - f_min < opt_tol, 3 |f_old - f| < prog_tol, ...)
    int evals, backtracks;
    int parity;            // which buffer set holds the current iterate
    int hist;              // L-BFGS: stored curvature pairs
    double cg, cy, cs;     // L-BFGS: next trial point = proj(x + cg g + cy (g - g_prev) + cs (x - x_prev))
    double rho[64];        // L-BFGS: 1 / <dx, dg> of the stored pairs, oldest first
};
constexpr int kStepScalars = 6;  // slots 1..6 of every rank travel in the all-gather of a sharded solve
struct DevOpts {
    int method;            // BATCH: 0 projected gradient, 1 Barzilai-Borwein, 2 mirror descent, 5 L-BFGS; 3 BB.solve in z; 4 mirror_descent.least_squares
    int corrections;       // method 5: history length (at most 64)
    int search;            // run line_search_np
    int has_f_min, max_iter;
    double f_min, opt_tol, prog_tol, min_eig;
    double tolerance;      // method 4: stop when max |x - x_prev| < tolerance; method 3: opt_tol of solvers.stopping
    double Lf;             // method 4
    int nranks, progress_cap;
};

// ---- SpMV epilogues ----------------------------------------------------------------------------
// An epilogue sees (row, dot) once per row and accumulates into acc[].
// r_old (optional): the residual at the current iterate; then <r_old, dr> and <dr, dr> with dr = r - r_old come out too.
// They give the objective anywhere on the segment between the iterate and the trial point without another product:
// f(x + tau dx) = f(x) + tau <r_old, dr> + 0.5 tau^2 <dr, dr>  (the differences are formed element by element: no cancellation).
__device__ __forceinline__ void residual_sums(double v, double ro, bool have_old, double (&acc)[3]) {
    acc[0] += v * v;
    if (have_old) {
        const double d = v - ro;
        acc[1] += ro * d;
        acc[2] += d * d;
    }
}
struct FinResidual {
    static __device__ __forceinline__ void store(double *o, int k, double v) {
        if (k == 0) {
            o[kScalF] = 0.5 * v;
            o[kScalRR] = v;
        } else {
            o[k == 1 ? kScalRdr : kScalDrdr] = v;
        }
    }
};
struct EpiResidual {  // r = A x - b ; f = 0.5 <r, r>       (algorithm_utils.py:91,93)
    static constexpr int NSUM = 3, NMAX = 0;
    double *r;
    const double *b;      // may be null (partial product of one rank: b is subtracted after the all-reduce)
    const double *r_old;  // may be null
    __device__ __forceinline__ void apply(int64_t i, double dot, double (&acc)[3]) const {
        const double v = b ? dot - b[i] : dot;
        r[i] = v;
        residual_sums(v, r_old ? r_old[i] : 0.0, r_old != nullptr, acc);
    }
    static __device__ __forceinline__ void store(double *out, int k, double v) { FinResidual::store(out, k, v); }
};
struct EpiPlain {  // out = M v
    static constexpr int NSUM = 1, NMAX = 0;
    double *out;
    __device__ __forceinline__ void apply(int64_t i, double dot, double (&acc)[1]) const {
        out[i] = dot;
        acc[0] += dot * dot;
    }
    static __device__ __forceinline__ void store(double *o, int k, double v) { o[kScalGnn] = v; }
};
struct EpiGradBB {  // g_new = A^T r and the Barzilai-Borwein / line-search / L-BFGS dot products (BATCH.py:89,99-100,160-167; algorithm_utils.py:120)
    static constexpr int NSUM = 5, NMAX = 1;
    double *g_new;
    const double *g, *x, *x_new;
    __device__ __forceinline__ void apply(int64_t i, double dot, double (&acc)[6]) const {
        const double go = g[i];
        const double dx = x_new[i] - x[i];
        const double dg = dot - go;
        g_new[i] = dot;
        acc[0] += dx * dg;
        acc[1] += dg * dg;
        acc[2] += go * dx;
        acc[3] += dot * dot;
        acc[4] += dg * go;
        acc[5] = fmax(acc[5], fabs(dx));
    }
    static __device__ __forceinline__ void store(double *o, int k, double v) {
        constexpr int slot[6] = {kScalSxy, kScalSyy, kScalGd, kScalGnn, kScalYgo, kScalStep};
        o[slot[k]] = v;
    }
};

// gather of the dense operand.  BSLS_GATHER selects the cache path: 1 (default) ld.global.cg -- L2 only: the
// operand (8 MB of r, or a 48 MB slice of x) never fits L1 and every gather is its own sector, so skipping the
// L1 allocation is ~2 % faster (C5: 5.78 -> 5.65 ms per product); 0 plain load; 2 ld.global.nc.
#ifndef BSLS_GATHER
#define BSLS_GATHER 1
#endif
__device__ __forceinline__ double gather(const double *v, int32_t j) {
#if BSLS_GATHER == 1
    return __ldcg(v + j);
#elif BSLS_GATHER == 2
    return __ldg(v + j);
#else
    return v[j];
#endif
}

__device__ __forceinline__ int pad16(int k) { return k + (k >> 4); }  // one spare double per 128-byte row

// ---- STREAM SpMV: thread per row, non-zeros staged through shared memory ----------------------
template <class Epi, int THREADS, int CHUNK>
__global__ void __launch_bounds__(THREADS)
spmv_stream_kernel(int64_t rows, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                   const double *__restrict__ val, const double *__restrict__ v, Epi epi, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;  // the solver has stopped: iterations enqueued ahead do nothing
    __shared__ double prod[CHUNK + CHUNK / 16 + 2];
    __shared__ int64_t s_range[2];
    constexpr int NA = Epi::NSUM + Epi::NMAX;
    double acc[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[k] = 0.0;
    const int tid = threadIdx.x;
    const int64_t ntiles = (rows + THREADS - 1) / THREADS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = tile * THREADS + tid;
        const bool valid = row < rows;
        const int64_t p0 = valid ? ptr[row] : 0, p1 = valid ? ptr[row + 1] : 0;
        const int64_t last = min(rows, (tile + 1) * THREADS) - 1;
        if (tid == 0) s_range[0] = p0;
        if (row == last) s_range[1] = p1;
        __syncthreads();
        const int64_t base = s_range[0], end = s_range[1];
        __syncthreads();  // s_range may be rewritten for the next tile from here on
        double sum = 0.0;
        for (int64_t c = base; c < end; c += CHUNK) {
            const int cnt = (int)min((int64_t)CHUNK, end - c);
            const int32_t *ci = idx + c;
            int k = tid;
            for (; k + 3 * THREADS < cnt; k += 4 * THREADS) {  // four gathers in flight per thread
                int32_t j[4];
                double a[4], w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) j[u] = __ldcs(ci + k + u * THREADS);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = val ? __ldcs(val + c + k + u * THREADS) : 1.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) w[u] = gather(v, j[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) prod[pad16(k + u * THREADS)] = a[u] * w[u];
            }
            for (; k < cnt; k += THREADS) {
                const double a = val ? __ldcs(val + c + k) : 1.0;
                prod[pad16(k)] = a * gather(v, __ldcs(ci + k));
            }
            __syncthreads();
            const int64_t lo = max(p0, c), hi = min(p1, c + cnt);
            for (int64_t p = lo; p < hi; ++p) sum += prod[pad16((int)(p - c))];  // left to right, as csr_matvec
            __syncthreads();
        }
        if (valid) epi.apply(row, sum, acc);
    }
    grid_reduce<Epi::NSUM, Epi::NMAX, THREADS, Epi>(acc, red);
}

// ---- ELL SpMV: every row has exactly L entries (A^T of a network whose routes all traverse L links) -----------------
// No row pointers, no staging: a thread owns a row, fetches its L column ids with 16-byte loads (the warp reads one
// contiguous 32 * 4L-byte span), has all L gathers in flight at once and adds them LEFT TO RIGHT (scipy's order: the
// result is bit-identical to the stream kernel and to the reference).  The kernel is bound by the L1TEX wavefront rate
// of the scattered gathers (~1 per clock per SM, /opt/skills/guides/B300_MICROARCH.md "L1tex wavefront queue"), so
// everything else is kept off that path: indices and values stream with .cs / no L1 allocation.
// (Fetching the operand through a texture object instead of LDG measured the same: 0.624 against 0.621 ms for 1.6e8
// gathers -- same L1TEX pipe -- and was dropped.)
template <int L> struct EllVec {
    static constexpr int W = (L % 4 == 0) ? 4 : ((L % 2 == 0) ? 2 : 1);
};
template <class Epi, int THREADS, int L, bool HAS_VAL>
__global__ void __launch_bounds__(THREADS)
spmv_ell_kernel(int64_t rows, const int32_t *__restrict__ idx, const double *__restrict__ val, const double *__restrict__ v,
                Epi epi, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;
    constexpr int NA = Epi::NSUM + Epi::NMAX;
    constexpr int W = EllVec<L>::W;
    double acc[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * THREADS;
    for (int64_t row = (int64_t)blockIdx.x * THREADS + threadIdx.x; row < rows; row += stride) {
        int32_t j[L];
        const int32_t *ri = idx + row * L;
        if constexpr (W == 4) {
#pragma unroll
            for (int k = 0; k < L; k += 4) {
                const int4 q = __ldcs(reinterpret_cast<const int4 *>(ri + k));
                j[k] = q.x, j[k + 1] = q.y, j[k + 2] = q.z, j[k + 3] = q.w;
            }
        } else if constexpr (W == 2) {
#pragma unroll
            for (int k = 0; k < L; k += 2) {
                const int2 q = __ldcs(reinterpret_cast<const int2 *>(ri + k));
                j[k] = q.x, j[k + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) j[k] = __ldcs(ri + k);
        }
        double w[L];
#pragma unroll
        for (int k = 0; k < L; ++k) w[k] = gather(v, j[k]);
        double sum = 0.0;
        if constexpr (HAS_VAL) {
            const double *rv = val + row * L;
#pragma unroll
            for (int k = 0; k < L; ++k) sum += __ldcs(rv + k) * w[k];
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) sum += w[k];  // 1.0 * w is exact: same bits as the multiply
        }
        epi.apply(row, sum, acc);
    }
    grid_reduce<Epi::NSUM, Epi::NMAX, THREADS, Epi>(acc, red);
}

// ---- VECTOR SpMV: LANES lanes per row ------------------------------------------------------------
template <class Epi, int THREADS, int LANES>
__global__ void __launch_bounds__(THREADS)
spmv_vector_kernel(int64_t rows, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                   const double *__restrict__ val, const double *__restrict__ v, Epi epi, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;
    constexpr int NA = Epi::NSUM + Epi::NMAX;
    double acc[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[k] = 0.0;
    const int sub = threadIdx.x & (LANES - 1);
    const int64_t group = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / LANES;
    const int64_t ngroups = (int64_t)gridDim.x * THREADS / LANES;
    const int64_t rounds = (rows + ngroups - 1) / ngroups;  // uniform trip count: shuffles stay converged
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t row = it * ngroups + group;
        const bool valid = row < rows;
        const int64_t p0 = valid ? ptr[row] : 0, p1 = valid ? ptr[row + 1] : 0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int64_t p = p0 + sub;
        for (; p + 3 * LANES < p1; p += 4 * LANES) {
            const int32_t j0 = __ldcs(idx + p), j1 = __ldcs(idx + p + LANES), j2 = __ldcs(idx + p + 2 * LANES),
                          j3 = __ldcs(idx + p + 3 * LANES);
            double a0 = 1.0, a1 = 1.0, a2 = 1.0, a3 = 1.0;
            if (val) {
                a0 = __ldcs(val + p);
                a1 = __ldcs(val + p + LANES);
                a2 = __ldcs(val + p + 2 * LANES);
                a3 = __ldcs(val + p + 3 * LANES);
            }
            const double w0 = gather(v, j0), w1 = gather(v, j1), w2 = gather(v, j2), w3 = gather(v, j3);
            s0 += a0 * w0;
            s1 += a1 * w1;
            s2 += a2 * w2;
            s3 += a3 * w3;
        }
        for (; p < p1; p += LANES) s0 += (val ? __ldcs(val + p) : 1.0) * gather(v, __ldcs(idx + p));
        double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int o = LANES / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (valid && sub == 0) epi.apply(row, sum, acc);
    }
    grid_reduce<Epi::NSUM, Epi::NMAX, THREADS, Epi>(acc, red);
}

// ---- VECTOR SpMV for short rows: every lane takes up to 8 entries of its row in ONE pass ----------------------------
// The column panels of A cut a link row into pieces of ~40-50 entries.  The kernel above walks such a piece in steps of
// 4 * LANES plus a remainder loop: two dependent rounds of gathers.  Here a lane fetches its (up to) eight indices at
// once, predicated on the row end, and has all its gathers in flight together; longer rows take further passes.
template <class Epi, int THREADS, int LANES>
__global__ void __launch_bounds__(THREADS)
spmv_vector8_kernel(int64_t rows, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                    const double *__restrict__ val, const double *__restrict__ v, Epi epi, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;
    constexpr int NA = Epi::NSUM + Epi::NMAX;
    double acc[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[k] = 0.0;
    const int sub = threadIdx.x & (LANES - 1);
    const int64_t group = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / LANES;
    const int64_t ngroups = (int64_t)gridDim.x * THREADS / LANES;
    const int64_t rounds = (rows + ngroups - 1) / ngroups;  // uniform trip count: shuffles stay converged
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t row = it * ngroups + group;
        const bool valid = row < rows;
        const int64_t p0 = valid ? ptr[row] : 0, p1 = valid ? ptr[row + 1] : 0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int64_t p = p0 + sub; p < p1; p += 8 * LANES) {
            int32_t j[8];
            double w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) j[u] = (p + u * LANES < p1) ? __ldcs(idx + p + u * LANES) : -1;
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = (j[u] >= 0) ? gather(v, j[u]) : 0.0;
            if (val) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (j[u] >= 0) w[u] *= __ldcs(val + p + u * LANES);
            }
            s0 += w[0];
            s1 += w[1];
            s2 += w[2];
            s3 += w[3];
            s0 += w[4];
            s1 += w[5];
            s2 += w[6];
            s3 += w[7];
        }
        double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int o = LANES / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (valid && sub == 0) epi.apply(row, sum, acc);
    }
    grid_reduce<Epi::NSUM, Epi::NMAX, THREADS, Epi>(acc, red);
}

// ---- vector kernels -------------------------------------------------------------------------------
// out = a x + b y, each product rounded on its own (np.add(x, -t*g, x_new); (1-t)*x + t*x_new)
// `out` may alias x or y (element i is read before it is written).
__global__ void __launch_bounds__(256) axpby_kernel(double *out, double a, const double *x, double b, const double *y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double u = a * x[i];
        out[i] = u + b * y[i];
    }
}

// out = a x + b y with a, b read from device scalars: a = sa * *pa (or sa), b = sb * (*pb0 - *pb1)
struct DevCoef {
    const double *p0, *p1;  // value = s * ((p0 ? *p0 : 1) - (p1 ? *p1 : 0))
    double s;
    __device__ __forceinline__ double get() const { return s * ((p0 ? *p0 : 1.0) - (p1 ? *p1 : 0.0)); }
};

struct FinDots {
    static __device__ __forceinline__ void store(double *o, int k, double v) { o[k] = v; }
};

// up to four dot products <x_k, y_k> in one pass and max |x_0 - y_0| ; results to out[0..3], out[4]
struct DotArgs {
    const double *x[4], *y[4];
    int count;
    int want_max;  // 1: out[4] = max |x0 - y0|
};
__global__ void __launch_bounds__(256) dots_kernel(DotArgs a, int64_t n, RedCtx red) {
    double acc[5] = {0, 0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    // two elements per trip: twice the loads in flight (each accumulator still adds its terms in index order)
    for (; i + stride < n; i += 2 * stride) {
        double xa[4], ya[4], xb[4], yb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < a.count) {
                xa[k] = a.x[k][i];
                ya[k] = a.y[k][i];
                xb[k] = a.x[k][i + stride];
                yb[k] = a.y[k][i + stride];
            }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < a.count) {
                acc[k] += xa[k] * ya[k];
                acc[k] += xb[k] * yb[k];
            }
        if (a.want_max) {
            acc[4] = fmax(acc[4], fabs(xa[0] - ya[0]));
            acc[4] = fmax(acc[4], fabs(xb[0] - yb[0]));
        }
    }
    for (; i < n; i += stride) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < a.count) acc[k] += a.x[k][i] * a.y[k][i];
        if (a.want_max) acc[4] = fmax(acc[4], fabs(a.x[0][i] - a.y[0][i]));
    }
    grid_reduce<4, 1, 256, FinDots>(acc, red);
}

// route-flow error metrics of LS_postprocess (python/main.py:112-134) for one iterate, in one pass:
//   out[0] = sum |s (xt - xh)|, out[1] = sum s xt, out[2] = #{xt - xh > thresh}, out[3] = sum (xt - xh)^2, out[4] = max s (xt - xh)
__global__ void __launch_bounds__(256) flow_metrics_kernel(const double *__restrict__ s, const double *__restrict__ xt,
                                                            const double *__restrict__ xh, double thresh, int64_t n, RedCtx red) {
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double sc = s ? s[i] : 1.0, t = xt[i];
        const double d = t - xh[i];
        const double sd = sc * d;
        acc[0] += fabs(sd);
        acc[1] += sc * t;
        acc[2] += (d > thresh) ? 1.0 : 0.0;
        acc[3] += d * d;
        acc[4] = fmax(acc[4], sd);
    }
    grid_reduce<4, 1, 256, FinDots>(acc, red);
}

// d <- d + c * v  (c from device scalars) and, in the same pass, <w, d_new> -> out[0]  scaled by *scale
// The L-BFGS two-loop recursion chains these without the host (LBFGS.py:60-71, BATCH.py:196-214).
struct FinScaled {
    static __device__ __forceinline__ void store(double *o, int k, double v) { o[k] = v; }
};
__global__ void __launch_bounds__(256) axpy_dot_kernel(double *__restrict__ d, DevCoef c, const double *v, const double *w, int64_t n,
                                                        RedCtx red) {
    const double cc = c.get();
    double acc[1] = {0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double di = d[i];
        if (v) {
            di = di + cc * v[i];
            d[i] = di;
        } else if (cc != 1.0) {
            di = cc * di;
            d[i] = di;
        }
        if (w) acc[0] += w[i] * di;
    }
    grid_reduce<1, 0, 256, FinScaled>(acc, red);
}

// r <- r - b ; f = 0.5 <r, r>   (after the all-reduce of the per-rank partial products)
__global__ void __launch_bounds__(256) residual_finish_kernel(double *__restrict__ r, const double *__restrict__ b, const double *__restrict__ r_old,
                                                               int64_t m, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;
    double acc[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < m; i += (int64_t)gridDim.x * 256) {
        const double v = r[i] - b[i];
        r[i] = v;
        residual_sums(v, r_old ? r_old[i] : 0.0, r_old != nullptr, acc);
    }
    grid_reduce<3, 0, 256, FinResidual>(acc, red);
}

// r_i = sum_p partial[p m + i] (- b_i), panels added in ascending order; f = 0.5 <r, r>.
// Closes the column-panelled product A x (see bsls_lsq_set_panels).
__global__ void __launch_bounds__(256) panel_reduce_kernel(double *__restrict__ r, const double *__restrict__ partial,
                                                            const double *__restrict__ b, const double *__restrict__ r_old, int64_t m,
                                                            int panels, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;
    double acc[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < m; i += (int64_t)gridDim.x * 256) {
        double v = partial[i];
        for (int p = 1; p < panels; ++p) v += partial[(int64_t)p * m + i];
        if (b) v -= b[i];
        r[i] = v;
        residual_sums(v, (b && r_old) ? r_old[i] : 0.0, b && r_old, acc);
    }
    grid_reduce<3, 0, 256, FinResidual>(acc, red);
}

// ---- device-resident solver steps ----------------------------------------------------------------------------------
// out = x - t g with t read from the solver state (np.add(x, -t*g, x_new), python/BATCH.py:38,91)
__global__ void __launch_bounds__(256) step_axpy_kernel(double *__restrict__ out, const double *__restrict__ x, const double *__restrict__ g,
                                                         const DevState *st, int64_t n) {
    if (st->done) return;
    const double nt = -st->t;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double u = nt * g[i];
        out[i] = x[i] + u;
    }
}

// BB.solve in z (python/BB.py:17-37): the four sums of an iteration in one pass over the z-vectors --
// <z - z_prev, g - g_prev>, |g - g_prev|^2, sum(g - g_prev) (the reference's "no change in gradient" test, :22) and |g|^2
// (solvers.stopping, python/solvers.py:47) -- into scalar slots 1..4.
struct FinZbb {
    static __device__ __forceinline__ void store(double *o, int k, double v) { o[kScalSxy + k] = v; }
};
__global__ void __launch_bounds__(256) zbb_dots_kernel(const double *__restrict__ z, const double *__restrict__ zp, const double *__restrict__ g,
                                                        const double *__restrict__ gp, int64_t n, RedCtx red, const int *__restrict__ skip) {
    if (skip && *skip) return;
    double acc[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double gi = g[i];
        const double dg = gi - gp[i], dz = z[i] - zp[i];
        acc[0] += dz * dg;
        acc[1] += dg * dg;
        acc[2] += dg;
        acc[3] += gi * gi;
    }
    grid_reduce<4, 0, 256, FinZbb>(acc, red);
}
// x_next = x - t g with t = <dx,dg> / <dg,dg> read from the scalar block (BB.py:26,29)
__global__ void __launch_bounds__(256) zbb_step_kernel(double *__restrict__ out, const double *__restrict__ z, const double *__restrict__ g,
                                                        const double *__restrict__ scal, const int *__restrict__ skip, int64_t n) {
    if (skip && *skip) return;
    const double t = scal[kScalSxy] / scal[kScalSyy];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double u = t * g[i];
        out[i] = z[i] - u;
    }
}

// L-BFGS trial point before projection: out = x + d with d = cg g + cy (g - g_prev) + cs (x - x_prev).  In the reference's
// solve_LBFGS every stored pair is a reference to the SAME two difference buffers (python/BATCH.py:153-155: the deques
// hold delta_x / delta_g themselves, which the loop overwrites in place), so the two-loop recursion (:196-214) only ever
// combines g, the latest delta_g and the latest delta_x: its inner products obey scalar recurrences (decide_step) and d
// is this one combination.  `out` may be the buffer of x_prev (element i is read before it is written).
__global__ void __launch_bounds__(256) lbfgs_step_kernel(double *out, const double *__restrict__ x, const double *xp,
                                                          const double *__restrict__ g, const double *__restrict__ gp, const DevState *st,
                                                          int64_t n) {
    if (st->done) return;
    const double cg = st->cg, cy = st->cy, cs = st->cs;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double gi = g[i], xi = x[i];
        double d = cg * gi;
        if (cy != 0.0 || cs != 0.0) {
            const double u = cy * (gi - gp[i]);
            const double v = cs * (xi - xp[i]);
            d = (d + u) + v;
        }
        out[i] = xi + d;
    }
}

// After a back-tracked line search (tau < 1) the trial point, its gradient and its residual are pulled back along the
// segment: v_new <- (1 - tau) v + tau v_new for x (the reference's own update, algorithm_utils.py:133), and -- the
// objective being quadratic -- for g and r, which the reference recomputes with two more products (:134).  tau == 0 is
// the reference's "step too small" reset (:125-131).  Nothing to do when the trial point was accepted as is.
__global__ void __launch_bounds__(256) commit_kernel(const DevState *st, double *__restrict__ xn, const double *__restrict__ x,
                                                      double *__restrict__ gn, const double *__restrict__ g, int64_t n,
                                                      double *__restrict__ rn, const double *__restrict__ r, int64_t m) {
    const double tau = st->tau;
    if (tau == 1.0) return;
    const double a = 1.0 - tau;
    const int64_t stride = (int64_t)gridDim.x * 256, i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (tau == 0.0) {
        for (int64_t i = i0; i < n; i += stride) {
            xn[i] = x[i];
            gn[i] = g[i];
        }
        for (int64_t i = i0; i < m; i += stride) rn[i] = r[i];
        return;
    }
    for (int64_t i = i0; i < n; i += stride) {
        const double u = a * x[i];
        xn[i] = u + tau * xn[i];
        const double w = a * g[i];
        gn[i] = w + tau * gn[i];
    }
    for (int64_t i = i0; i < m; i += stride) {
        const double u = a * r[i];
        rn[i] = u + tau * rn[i];
    }
}

// The scalar half of one iteration of BATCH.solve / solve_BB / solve_MD (python/BATCH.py:33-47,84-102,230-247):
// line_search_np (python/algorithm_utils.py:113-137), the new step and `stopping` (:158-172), by one thread.
//   scal        the scalar block the reducing kernels of this evaluation wrote (f_trial, <r,dr>, <dr,dr> replicated on
//               every rank; slots 1..5 = this rank's share of <dx,dg>, <dg,dg>, <g,dx>, <g_new,g_new>, max|dx|)
//   gathered    nranks x 5: slots 1..5 of every rank (all-gather); null on one GPU
// The Armijo test of a back-tracked point needs f there: f + tau <r,dr> + 0.5 tau^2 <dr,dr> (exact for this objective),
// so a back-track costs no product; the compounding x_new <- (1-t) x + t x_new with t = .8, .64, ... shrinks dx by
// tau = prod t, and <g,dx>, max|dx| scale with it.
__device__ __forceinline__ void decide_step(DevState *st, const double *scal, const double *gathered, const DevOpts &o,
                                            double *progress_f, double *progress_t, int first) {
    if (st->done) {  // an iteration enqueued ahead of the stop: nothing to decide, nothing to commit
        st->tau = 1.0;
        return;
    }
    unsigned long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    if (o.method == 3) {  // BB.solve (python/BB.py:17-37) with solvers.stopping (python/solvers.py:40-63)
        if (first) {      // st->i was set by the host to the number of iterations already done (segmented runs)
            st->parity = 0;
            st->done = 0;
            return;
        }
        if (scal[kScalGd] == 0.0) {  // sum(delta_g) == 0: 'Exiting... no change in gradient', x stays where it is (BB.py:22-24)
            st->done = 5;
            return;
        }
        const double fx = scal[kScalF];
        st->t = scal[kScalSxy] / scal[kScalSyy];
        st->f = fx;
        st->i += 1;
        st->parity += 1;
        st->evals += 1;
        if (st->i >= o.max_iter)
            st->done = 1;
        else if (scal[kScalGnn] <= o.opt_tol * (1.0 + fabs(fx)))
            st->done = 6;   // 'Exiting... norm(grad) too small'
        else if (sqrt(scal[kScalSyy]) == 0.0)
            st->done = 5;
        return;
    }
    if (o.method == 4) {  // mirror_descent.least_squares (mirror_descent.py:37-52): no objective, stop on max |x - x_prev|
        if (first) {
            st->i = 1;
            st->parity = 0;
            st->done = o.max_iter < 1 ? 1 : 0;
            return;
        }
        st->change = scal[kScalMax0];
        st->parity ^= 1;
        st->i += 1;
        st->evals += 1;
        if (st->change < o.tolerance)
            st->done = 4;
        else if (st->i > o.max_iter)
            st->done = 1;
        return;
    }
    double gd_acc = 0.0, ygo_acc = 0.0, sxy_raw = 0.0, syy_raw = 0.0, tau_acc = 1.0;
    if (first) {  // f = obj(x, g) of the starting point (BATCH.py:29,77,228)
        st->hist = 0;
        st->cg = -1.0;  // first step: x - g (BATCH.py:149)
        st->cy = st->cs = 0.0;
        st->f = scal[kScalF];
        st->f_old = __longlong_as_double(0x7ff0000000000000LL);
        st->i = 1;
        st->evals = 1;
        st->backtracks = 0;
        st->tau = 1.0;
        st->sxy = st->syy = 0.0;
        st->parity = 0;
        st->change = 0.0;
        st->stop_value = 0.0;
        if (progress_f && o.progress_cap > 0) {
            progress_f[0] = st->f;
            progress_t[0] = (double)now;
        }
    } else {
        double sxy = scal[kScalSxy], syy = scal[kScalSyy], gd = scal[kScalGd], step = scal[kScalStep], ygo = scal[kScalYgo];
        if (gathered) {
            sxy = syy = gd = step = ygo = 0.0;
            for (int r = 0; r < o.nranks; ++r) {  // rank order: every rank forms the same sums
                sxy += gathered[r * kStepScalars + 0];
                syy += gathered[r * kStepScalars + 1];
                gd += gathered[r * kStepScalars + 2];
                step = fmax(step, gathered[r * kStepScalars + 4]);
                ygo += gathered[r * kStepScalars + 5];
            }
        }
        gd_acc = gd, ygo_acc = ygo, sxy_raw = sxy, syy_raw = syy;
        const double f = st->f;
        double f_new = scal[kScalF];
        double tau = 1.0;
        int bt = 0;
        if (o.search) {
            const double rdr = scal[kScalRdr], drdr = scal[kScalDrdr];
            double t = 1.0, gdt = gd, stept = step;
            while (f_new > f + 1e-4 * gdt) {
                t *= .8;
                if (stept < 1e-12) {  // step too small: stay where we are
                    tau = 0.0;
                    f_new = f;
                    break;
                }
                tau *= t;
                gdt = tau * gd;
                stept = tau * step;
                const double half = 0.5 * tau;
                f_new = f + (tau * rdr + (half * tau) * drdr);
                ++bt;
            }
        }
        tau_acc = tau;
        st->tau = tau;
        st->backtracks += bt;
        st->evals += 1;
        st->sxy = (tau * tau) * sxy;
        st->syy = (tau * tau) * syy;
        st->f_old = f;
        st->f = f_new;
        st->parity ^= 1;
        st->i += 1;
        if (o.method == 2) st->change = scal[kScalMax0];
        if (progress_f && st->i - 1 < o.progress_cap) {
            progress_f[st->i - 1] = f_new;
            progress_t[st->i - 1] = (double)now;
        }
    }
    // the step of the next trial point
    const int i = st->i;
    if (o.method == 1)
        st->t = (i == 1) ? 1.0 : st->sxy / st->syy;  // BATCH.py:87-91
    else
        st->t = 1.0 / (o.min_eig * i + 1.0);         // decreasing_step_size(i, 1.0, min_eig), BATCH.py:38,238
    if (o.method == 5 && !first) {
        // solve_LBFGS (BATCH.py:150-167) for iteration i: append the pair of the step just taken, keep the last
        // `corrections`, then BB direction while i <= 5, else the two-loop recursion -- on scalars, see lbfgs_step_kernel.
        // s = dx, y = dg (tau-scaled), g = gradient at the new iterate = g_old + tau dg:
        const double c = st->sxy, yy = st->syy;                             // <s,y>, <y,y>
        const double sg = tau_acc * gd_acc + (tau_acc * tau_acc) * sxy_raw;  // <s,g>
        const double yg = tau_acc * ygo_acc + (tau_acc * tau_acc) * syy_raw; // <y,g>
        int m = st->hist;
        if (m == o.corrections) {  // popleft (the append below brings the count back to `corrections`)
            for (int j = 1; j < m; ++j) st->rho[j - 1] = st->rho[j];
            --m;
        }
        st->rho[m++] = 1.0 / c;
        st->hist = m;
        const double t = c / yy;
        if (i <= 5) {
            st->cg = -t;
            st->cy = st->cs = 0.0;
        } else {
            double al[64];
            double sd = sg, A = 0.0;
            for (int j = m - 1; j >= 0; --j) {  // alpha_j = rho_j <s,d>; d -= alpha_j y
                al[j] = st->rho[j] * sd;
                sd -= al[j] * c;
                A += al[j];
            }
            double yd = t * (yg - A * yy), Bc = 0.0;  // d *= t; <y,d>
            for (int j = 0; j < m; ++j) {              // beta_j = rho_j <y,d>; d += s (alpha_j - beta_j)
                const double beta = st->rho[j] * yd;
                const double coef = al[j] - beta;
                yd += coef * c;
                Bc += coef;
            }
            st->cg = -t;       // d = -(t g - t A y + Bc s)
            st->cy = t * A;
            st->cs = -Bc;
        }
    }
    // algorithm_utils.stopping (:158-172): later tests overwrite the reason of earlier ones
    int code = 0;
    if (i == o.max_iter) code = 1;
    if (o.has_f_min && st->f - o.f_min < o.opt_tol) {
        code = 2;
        st->stop_value = st->f - o.f_min;
    }
    if (fabs(st->f_old - st->f) < o.prog_tol) {
        code = 3;
        st->stop_value = fabs(st->f_old - st->f);
    }
    st->done = code;
}

// wait_flags != null (peer-memory build, p2p.cuh): the step scalars of the other ranks arrive by NVLink stores; lane q waits
// until rank q has raised its flag for this evaluation before lane 0 reads them
__global__ void decide_kernel(DevState *st, const double *__restrict__ scal, const double *gathered, DevOpts o,
                              double *__restrict__ progress_f, double *__restrict__ progress_t, int first,
                              const unsigned long long *wait_flags = nullptr, unsigned long long epoch = 0) {
    if (blockIdx.x) return;
    if (wait_flags && !st->done && (int)threadIdx.x < o.nranks) {
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(wait_flags + threadIdx.x) : "memory");
        } while (v < epoch);
    }
    __syncwarp();
    if (threadIdx.x) return;
    decide_step(st, scal, gathered, o, progress_f, progress_t, first);
}

// ---- per-block kernels: G lanes per block ---------------------------------------------------------
struct BlockLayout {
    const int32_t *starts;  // nb + 1 entries (last = n); ignored when uniform > 0
    int nb;
    int uniform;            // common block size or 0
    int first;              // first index of block 0
    __device__ __forceinline__ void range(int b, int &s, int &e) const {
        if (uniform > 0) {
            s = first + b * uniform;
            e = s + uniform;
        } else {
            s = starts[b];
            e = starts[b + 1];
        }
    }
};

// mirror-descent update: x_new = x * exp(-t g), then every block divided by its sum.
//   per_block_log == 0: t = step                          (BATCH.py:239-241, algorithm_utils.py:175-179)
//   per_block_log == 1: t = sqrt(2 ln K_block) / step     (mirror_descent.py:26-28,39-47; step = sqrt(k) * Lf)
// out[kScalMax0] = max |x_new - x| (mirror_descent.py:50)
struct FinMax {
    static __device__ __forceinline__ void store(double *o, int k, double v) { o[kScalMax0] = v; }
};
// st != null: the step comes from the solver state on the device (step = st->t for the BATCH loop, sqrt(i) * step for
// mirror_descent.least_squares) and the kernel does nothing once the solver has stopped
__device__ __forceinline__ double md_step(double step, int per_block_log, const DevState *st) {
    if (!st) return step;
    return per_block_log ? sqrt((double)st->i) * step : st->t;
}
template <int G>
__global__ void __launch_bounds__(256) md_update_kernel(double *__restrict__ xn, const double *__restrict__ x, const double *__restrict__ g,
                                                         double step, int per_block_log, BlockLayout lay, RedCtx red, const DevState *st) {
    if (st && st->done) return;
    step = md_step(step, per_block_log, st);
    const int sub = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
    const int64_t ngroups = (int64_t)gridDim.x * 256 / G;
    const int64_t rounds = (lay.nb + ngroups - 1) / ngroups;
    double acc[1] = {0};
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t b = it * ngroups + group;
        const bool valid = b < lay.nb;
        int s = 0, e = 0;
        if (valid) lay.range((int)b, s, e);
        double t = step;
        if (per_block_log) t = sqrt(2.0 * log((double)(e - s > 0 ? e - s : 1))) / step;
        double sum = 0.0;
        for (int i = s + sub; i < e; i += G) {
            const double up = per_block_log ? g[i] * t : -t * g[i];
            const double w = x[i] * exp(per_block_log ? -up : up);
            xn[i] = w;
            sum += w;
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        for (int i = s + sub; i < e; i += G) {
            const double w = xn[i] / sum;
            xn[i] = w;
            acc[0] = fmax(acc[0], fabs(w - x[i]));
        }
    }
    grid_reduce<0, 1, 256, FinMax>(acc, red);
}

// The same update with every block held in the registers of its G lanes (blocks of at most R*G entries): x and g are
// read once, x_new is written once (24 n bytes; the two-pass kernel above re-reads what it wrote).  Same order of
// additions as md_update_kernel (per lane in steps of G, then the xor tree), so both give the same bits.
template <int G, int R>
__global__ void __launch_bounds__(256) md_update_reg_kernel(double *__restrict__ xn, const double *__restrict__ x, const double *__restrict__ g,
                                                             double step, int per_block_log, BlockLayout lay, RedCtx red, const DevState *st) {
    if (st && st->done) return;
    step = md_step(step, per_block_log, st);
    const int sub = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
    const int64_t ngroups = (int64_t)gridDim.x * 256 / G;
    const int64_t rounds = (lay.nb + ngroups - 1) / ngroups;
    double acc[1] = {0};
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t b = it * ngroups + group;
        const bool valid = b < lay.nb;
        int s = 0, e = 0;
        if (valid) lay.range((int)b, s, e);
        double t = step;
        if (per_block_log) t = sqrt(2.0 * log((double)(e - s > 0 ? e - s : 1))) / step;
        double xv[R], w[R];
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = s + sub + r * G;
            xv[r] = 0.0;
            w[r] = 0.0;
            if (i < e) {
                xv[r] = x[i];
                const double up = per_block_log ? g[i] * t : -t * g[i];
                w[r] = xv[r] * exp(per_block_log ? -up : up);
                sum += w[r];
            }
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = s + sub + r * G;
            if (i < e) {
                const double q = w[r] / sum;
                xn[i] = q;
                acc[0] = fmax(acc[0], fabs(q - xv[r]));
            }
        }
    }
    grid_reduce<0, 1, 256, FinMax>(acc, red);
}

// per-block scale: y[block k] <- y * f_k or y / f_k (get_solver_parts with f given, algorithm_utils.py:232-265)
template <int G>
__global__ void __launch_bounds__(256) block_scale_kernel(double *__restrict__ y, const double *__restrict__ f, int divide, BlockLayout lay) {
    const int sub = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
    const int64_t ngroups = (int64_t)gridDim.x * 256 / G;
    for (int64_t b = group; b < lay.nb; b += ngroups) {
        int s, e;
        lay.range((int)b, s, e);
        const double k = f[b];
        for (int i = s + sub; i < e; i += G) y[i] = divide ? y[i] / k : y[i] * k;
    }
}

// x = x0 + N z: x_l = z_l - z_{l-1} inside a block (z_{-1} = 0, the last entry is -z_last [+ 1 when add_x0]).
// Block b of x starts at s_b, the matching block of z at s_b - b (bsls_utils.py:139-162,327-328).
template <int G>
__global__ void __launch_bounds__(256) nz_kernel(double *__restrict__ x, const double *__restrict__ z, int add_x0, BlockLayout lay) {
    const int sub = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
    const int64_t ngroups = (int64_t)gridDim.x * 256 / G;
    for (int64_t b = group; b < lay.nb; b += ngroups) {
        int s, e;
        lay.range((int)b, s, e);
        const int64_t zoff = (int64_t)b + lay.first;  // z index = x index - zoff
        for (int i = s + sub; i < e; i += G) {
            const double hi = (i < e - 1) ? z[i - zoff] : 0.0;
            const double lo = (i > s) ? z[i - zoff - 1] : 0.0;
            double v = hi - lo;
            if (add_x0 && i == e - 1) v = 1.0 + v;
            x[i] = v;
        }
    }
}

// z-gradient = N^T v: (N^T v)_l = v_l - v_{l+1} for l < K - 1
template <int G>
__global__ void __launch_bounds__(256) ntv_kernel(double *__restrict__ zg, const double *__restrict__ v, BlockLayout lay) {
    const int sub = threadIdx.x & (G - 1);
    const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
    const int64_t ngroups = (int64_t)gridDim.x * 256 / G;
    for (int64_t b = group; b < lay.nb; b += ngroups) {
        int s, e;
        lay.range((int)b, s, e);
        const int64_t zoff = (int64_t)b + lay.first;
        for (int i = s + sub; i < e - 1; i += G) zg[i - zoff] = v[i] - v[i + 1];
    }
}

// x -> z: running sums of a block without its last entry, summed left to right by one thread so
// that the result is bit-identical to the reference (c_extensions.pyx:195-220, bsls_utils.py:267-287)
__global__ void __launch_bounds__(256) x2z_kernel(const double *__restrict__ x, double *__restrict__ z, BlockLayout lay) {
    for (int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x; b < lay.nb; b += (int64_t)gridDim.x * 256) {
        int s, e;
        lay.range((int)b, s, e);
        const int64_t zoff = b + lay.first;
        double run = 0.0;
        for (int i = s; i < e - 1; ++i) {
            run += x[i];
            z[i - zoff] = run;
        }
    }
}

// x -> z for uniform layouts with short blocks: a CTA stages THREADS whole blocks in shared memory (coalesced
// cp.async into rows of odd pitch), every thread sums ITS block left to right in place (the reference's order, so the
// result is bit-identical), and the K - 1 running sums of every block go out coalesced.  The thread-per-block kernel
// above reads and writes with a stride of one block per lane and stops near 30 % of the HBM peak.
constexpr int kX2zTileThreads = 128;
constexpr int kX2zTileMaxK = 64;
__global__ void __launch_bounds__(kX2zTileThreads)
x2z_tile_kernel(const double *__restrict__ x, double *__restrict__ z, int nb, int K, FastDiv kdiv, FastDiv kdiv1) {
    extern __shared__ __align__(16) double x2z_rows[];
    const int KS = K | 1, K1 = K - 1, tid = threadIdx.x;
    const int ntiles = (nb + kX2zTileThreads - 1) / kX2zTileThreads;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nblk = min(kX2zTileThreads, nb - tile * kX2zTileThreads);
        const double *gx = x + (size_t)tile * kX2zTileThreads * K;
        double *gz = z + (size_t)tile * kX2zTileThreads * K1;
        const int nel = nblk * K;
        for (int e = tid; e < nel; e += kX2zTileThreads) {
            const uint32_t r = fdiv((uint32_t)e, kdiv);
            cp_async_elem<8>(&x2z_rows[r * KS + (e - r * K)], gx + e);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (tid < nblk) {
            double *row = x2z_rows + tid * KS;
            double run = 0.0;
            for (int c = 0; c < K1; ++c) {
                run += row[c];
                row[c] = run;
            }
        }
        __syncthreads();
        const int nz = nblk * K1;
        for (int i = tid; i < nz; i += kX2zTileThreads) {
            const uint32_t r = fdiv((uint32_t)i, kdiv1);
            gz[i] = x2z_rows[r * KS + (i - r * K1)];
        }
        __syncthreads();
    }
}

// z -> x (adjacent differences, last entry 1 - z_last; c_extensions.pyx:223-248) is nz_kernel with add_x0 = 1.

}  // namespace bsls
