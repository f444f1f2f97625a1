// solver_tiny.cuh -- the whole BATCH.solve / solve_BB loop (python/BATCH.py:7-106) in ONE launch of ONE CTA, for problems
// whose vectors fit the shared memory of an SM (BASELINE config 1: 1,000 OD blocks x 5 routes, 2,000 links).
//
// Such problems are latency-bound: an iteration moves ~1.6 MB, and as a chain of kernels it costs the launch and drain of
// each (44 us per objective evaluation in round 1, 4x one CPU core).  Here x, g, r of the iterate and of the trial point
// and b live in shared memory for the whole solve; per iteration the CTA forms x_new = proj(x - t g) (one thread per OD
// block, the reference's own routine: copy, sort descending, running sum left to right, last passing index --
// python/c_extensions/proj_simplex.h:17-34,50-74), r_new = A x_new - b (one thread per link row, entries added left to
// right as scipy's csr_matvec does), g_new = A^T r_new with the step / line-search dot products, takes the decisions of
// decide_step (lsq.cuh) in thread 0 and pulls a back-tracked trial point back.  Only the index / value arrays of A and
// A^T stream from L2.  No host round trip, no kernel boundary, no global synchronisation inside the solve.
#pragma once
#include "lsq.cuh"

namespace bsls {

constexpr int kTinyThreads = 1024;
constexpr int kTinyMaxBlock = 64;  // longest OD block the per-thread projection takes (insertion sort: short blocks only)

struct TinyArgs {
    int n, m, nb;
    const int32_t *starts;   // nb + 1 block starts
    const int64_t *a_ptr, *t_ptr;
    const int32_t *a_idx, *t_idx;
    const double *a_val, *t_val;  // may be null: implicit ones
    const double *b;
    double *x;               // in: starting point; out: solution
    DevState *st;
    double *progress_f, *progress_t;
    int proj_mode;           // 0 simplex, 1 l1-ball
};

inline size_t tiny_smem_bytes(int n, int m) { return sizeof(double) * (4 * (size_t)n + 3 * (size_t)m + 64); }

// deterministic CTA reduction of NS sums and NM maxima (fixed tree); result valid in thread 0
template <int NS, int NM> __device__ __forceinline__ void tiny_reduce(double (&acc)[NS + NM], double *s_red /* 32 * (NS+NM) */, double *out) {
    constexpr int N = NS + NM;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = (k < NS) ? v + u : fmax(v, u);
        }
        if (lane == 0) s_red[wid * N + k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double v = s_red[k];
            for (int w = 1; w < kTinyThreads / 32; ++w) v = (k < NS) ? v + s_red[w * N + k] : fmax(v, s_red[w * N + k]);
            out[k] = v;
        }
    }
    __syncthreads();
}

// proj_simplex (proj_simplex.h:17-34) in place on w[0..K) (shared memory), u[0..K) = scratch for the sorted copy:
// the reference's own order of operations
__device__ __forceinline__ void tiny_proj_simplex(double *w, double *u, int K) {
    for (int i = 0; i < K; ++i) {  // insertion sort, descending (the sorted values do not depend on the algorithm)
        const double v = w[i];
        int j = i;
        while (j > 0 && u[j - 1] < v) {
            u[j] = u[j - 1];
            --j;
        }
        u[j] = v;
    }
    double sum = u[0];
    double lambda = 1. - sum;
    for (int i = 1; i < K; ++i) {
        sum += u[i];
        const double tmp = (1. - sum) / ((double)i + 1.);
        if (u[i] + tmp > 0) lambda = tmp;
    }
    for (int i = 0; i < K; ++i) {
        const double t = lambda + w[i];
        w[i] = t > 0. ? t : 0.;
    }
}

// One row of a CSR product, entries added left to right (scipy's csr_matvec), the index / value loads issued eight at
// a time so that their L2 latency overlaps (the adds stay in order).
__device__ __forceinline__ double tiny_row_dot(const int32_t *__restrict__ idx, const double *__restrict__ val, int64_t p0, int64_t p1,
                                               const double *v) {
    double sum = 0.0;
    int64_t p = p0;
    for (; p + 8 <= p1; p += 8) {
        int32_t j[8];
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) j[k] = idx[p + k];
        if (val) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = val[p + k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += (val ? a[k] : 1.0) * v[j[k]];
    }
    for (; p < p1; ++p) sum += (val ? val[p] : 1.0) * v[idx[p]];
    return sum;
}

__global__ void __launch_bounds__(kTinyThreads, 1) solver_tiny_kernel(TinyArgs a, DevOpts o) {
    extern __shared__ __align__(16) double tiny_sm[];
    const int n = a.n, m = a.m, tid = threadIdx.x;
    double *xs[2] = {tiny_sm, tiny_sm + n};
    double *gs[2] = {tiny_sm + 2 * n, tiny_sm + 3 * n};
    double *rs[2] = {tiny_sm + 4 * n, tiny_sm + 4 * n + m};
    double *bs = tiny_sm + 4 * n + 2 * m;
    double *scal = bs + m;  // kScalCount scalars + reduction scratch
    __shared__ double s_red[32 * 5];
    __shared__ double s_out[5];
    __shared__ DevState s_state;  // the solver state lives on chip for the whole solve; copied out at the end
    DevState *st = &s_state;
    if (tid == 0) {
        s_state = DevState{};
    }

    for (int i = tid; i < n; i += kTinyThreads) xs[0][i] = a.x[i];
    for (int i = tid; i < m; i += kTinyThreads) bs[i] = a.b[i];
    if (tid < kScalCount) scal[tid] = 0.0;
    __syncthreads();

    // r = A x - b and the sums of EpiResidual; g = A^T r and the sums of EpiGradBB (or only <g,g>)
    auto residual = [&](const double *x, double *r, const double *r_old) {
        double acc[3] = {0, 0, 0};
        for (int row = tid; row < m; row += kTinyThreads) {
            const double sum = tiny_row_dot(a.a_idx, a.a_val, a.a_ptr[row], a.a_ptr[row + 1], x);
            const double v = sum - bs[row];
            r[row] = v;
            residual_sums(v, r_old ? r_old[row] : 0.0, r_old != nullptr, acc);
        }
        tiny_reduce<3, 0>(acc, s_red, s_out);
        if (tid == 0) {
            scal[kScalF] = 0.5 * s_out[0];
            scal[kScalRR] = s_out[0];
            scal[kScalRdr] = s_out[1];
            scal[kScalDrdr] = s_out[2];
        }
    };
    auto gradient = [&](const double *r, double *g_new, const double *g, const double *x, const double *x_new) {
        double acc[5] = {0, 0, 0, 0, 0};
        for (int row = tid; row < n; row += kTinyThreads) {
            const double dot = tiny_row_dot(a.t_idx, a.t_val, a.t_ptr[row], a.t_ptr[row + 1], r);
            g_new[row] = dot;
            acc[3] += dot * dot;
            if (g) {
                const double go = g[row], dx = x_new[row] - x[row], dg = dot - go;
                acc[0] += dx * dg;
                acc[1] += dg * dg;
                acc[2] += go * dx;
                acc[4] = fmax(acc[4], fabs(dx));
            }
        }
        tiny_reduce<4, 1>(acc, s_red, s_out);
        if (tid == 0) {
            scal[kScalSxy] = s_out[0];
            scal[kScalSyy] = s_out[1];
            scal[kScalGd] = s_out[2];
            scal[kScalGnn] = s_out[3];
            scal[kScalStep] = s_out[4];
        }
    };

    // f = obj(x, g) at the starting point
    residual(xs[0], rs[0], nullptr);
    gradient(rs[0], gs[0], nullptr, nullptr, nullptr);
    if (tid == 0) decide_step(st, scal, nullptr, o, a.progress_f, a.progress_t, 1);
    __syncthreads();

    int cur = 0;
    while (!st->done) {
        const int nxt = cur ^ 1;
        const double nt = -st->t;
        // ---- x_new = proj(x - t g): one thread per OD block ---------------------------------------
        for (int blk = tid; blk < a.nb; blk += kTinyThreads) {
            const int s = a.starts[blk], K = a.starts[blk + 1] - s;
            double *w = xs[nxt] + s;   // the trial point is formed in place; the trial gradient's slot is the sort scratch
            double sum = 0.0;
            for (int i = 0; i < K; ++i) {
                const double u = nt * gs[cur][s + i];  // np.add(x, -t*g, x_new): product and sum rounded separately
                double v = xs[cur][s + i] + u;
                if (a.proj_mode == 1) {                // proj_multi_ball (proj_simplex.h:54-62)
                    if (v < 0.0)
                        v = 0.0;
                    else
                        sum += v;
                }
                w[i] = v;
            }
            if (a.proj_mode == 0 || sum > 1.0) tiny_proj_simplex(w, gs[nxt] + s, K);
        }
        __syncthreads();
        // ---- objective and gradient at the trial point, decision, pull-back ------------------------------
        residual(xs[nxt], rs[nxt], rs[cur]);
        gradient(rs[nxt], gs[nxt], gs[cur], xs[cur], xs[nxt]);
        if (tid == 0) decide_step(st, scal, nullptr, o, a.progress_f, a.progress_t, 0);
        __syncthreads();
        const double tau = st->tau;
        if (tau != 1.0) {  // commit_kernel's update on the shared-memory vectors
            const double c = 1.0 - tau;
            for (int i = tid; i < n; i += kTinyThreads) {
                if (tau == 0.0) {
                    xs[nxt][i] = xs[cur][i];
                    gs[nxt][i] = gs[cur][i];
                } else {
                    const double u = c * xs[cur][i];
                    xs[nxt][i] = u + tau * xs[nxt][i];
                    const double v = c * gs[cur][i];
                    gs[nxt][i] = v + tau * gs[nxt][i];
                }
            }
            for (int i = tid; i < m; i += kTinyThreads) {
                if (tau == 0.0) {
                    rs[nxt][i] = rs[cur][i];
                } else {
                    const double u = c * rs[cur][i];
                    rs[nxt][i] = u + tau * rs[nxt][i];
                }
            }
            __syncthreads();
        }
        cur = nxt;
    }
    for (int i = tid; i < n; i += kTinyThreads) a.x[i] = xs[cur][i];
    if (tid == 0) *a.st = s_state;
}

}  // namespace bsls
