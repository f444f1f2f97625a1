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
// A^T stream from L2, from sliced-ELL copies made once per problem (SellMatrix).  No host round trip, no kernel boundary, no global synchronisation inside the solve.
#pragma once
#include "lsq.cuh"
#include "simplex_core.cuh"

namespace bsls {

constexpr int kTinyThreads = 1024;
constexpr int kTinyMaxBlock = 64;  // longest OD block the per-thread projection takes (insertion sort: short blocks only)

// A sparse matrix in sliced-ELL form for the single-CTA solver: rows in groups of 32 (one warp), every group stored
// entry-major -- slab[k * 32 + lane] is the k-th entry of row 32 g + lane, padded with -1 to the longest row of the group.
// A warp reading entry k of its 32 rows touches ONE 128-byte line, where the CSR layout made every lane walk its own row
// (one L1TEX wavefront per lane and entry: measured 71 % L1TEX utilisation of the one SM, 40 us per iteration on config 1);
// each thread still adds the entries of its row left to right.
struct SellMatrix {
    const int32_t *idx;   // group slabs
    const double *val;    // same layout, or null: implicit ones
    const int32_t *goff;  // groups + 1 offsets into idx / val
};

struct TinyArgs {
    int n, m, nb;
    const int32_t *starts;   // nb + 1 block starts
    SellMatrix A, AT;
    const double *b;
    double *x;               // in: starting point; out: solution
    DevState *st;
    double *progress_f, *progress_t;
    int proj_mode;           // 0 simplex, 1 l1-ball
    long long *prof;         // development: 16 phase cycle counters of CTA 0 / thread 0 (BSLS_TINY_PROF=1), else null
};

inline size_t tiny_smem_bytes(int n, int m) { return sizeof(double) * (4 * (size_t)(n + 1) + 3 * (size_t)(m + 1) + 64); }

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

// One row of a product, entries added left to right (scipy's csr_matvec order).  Padding entries point at a slot that holds
// 0.0 (sum + 0.0 keeps the bits of sum), group widths are multiples of 4: the loop is unconditional, four index loads in
// flight per pass.  `vec` is an offset into the shared-memory arena: the compiler emits LDS, not generic loads.
// R rows of one thread (row0, row0 + kTinyThreads, ...) at once: their index loads are in flight together, so a pass costs
// ONE round trip to L2 whatever R is (a thread owns 2 rows of A and 5 of A^T on config 1; one row at a time made the
// solve latency-bound: 33 dependent round trips per iteration).  Rows beyond `total` give 0.
template <int R, bool HV>
__device__ __forceinline__ void tiny_rows_dot_impl(const SellMatrix &M, int row0, int total, const double *sm, int vec, double (&sum)[R]) {
    const char *base = reinterpret_cast<const char *>(sm + vec);  // the stored ids are BYTE offsets (8 * column): LDS [id + base]
    auto at = [&](int off) { return *reinterpret_cast<const double *>(base + off); };
    const int lane = row0 & 31;
    const int32_t *ip[R];
    const double *vp[R];
    int width[R], wmax = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * kTinyThreads;
        sum[r] = 0.0;
        width[r] = 0;
        ip[r] = M.idx;
        vp[r] = M.val;
        if (row < total) {  // uniform over the warp (total is a multiple of 32 for the padded groups or the whole warp is out)
            const int g = row >> 5;
            const int o = M.goff[g];
            ip[r] = M.idx + o + lane;
            vp[r] = HV ? M.val + o + lane : nullptr;
            width[r] = (M.goff[g + 1] - o) >> 5;
        }
        wmax = max(wmax, width[r]);
    }
    for (int k = 0; k < wmax; k += 4) {
        int j[R][4];
        double a[HV ? R : 1][4];
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (k < width[r]) {
#pragma unroll
                for (int u = 0; u < 4; ++u) j[r][u] = ip[r][32 * (k + u)];
                if (HV) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[r][u] = vp[r][32 * (k + u)];
                }
            }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (k < width[r]) {
#pragma unroll
                for (int u = 0; u < 4; ++u) sum[r] += HV ? a[r][u] * at(j[r][u]) : at(j[r][u]);
            }
    }
}
// index-only matrices (0/1 incidence) take R rows at once; with a value array the registers allow one row at a time
template <int R>
__device__ __forceinline__ void tiny_rows_dot(const SellMatrix &M, int row0, int total, const double *sm, int vec, double (&sum)[R]) {
    if (!M.val) {
        tiny_rows_dot_impl<R, false>(M, row0, total, sm, vec, sum);
    } else {
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            double one[1];
            tiny_rows_dot_impl<1, true>(M, row0 + r * kTinyThreads, total, sm, vec, one);
            sum[r] = one[0];
        }
    }
}

// x_new block = proj(w) for one OD block in shared memory: blocks of at most 8 values go through the library's register
// sorting network and its restatement of the reference's scan (simplex_core.cuh: bit-identical to proj_simplex.h:17-34),
// longer ones through the insertion sort above.
__device__ __forceinline__ void tiny_project_block(double *w, double *scratch, int K) {
    if (K <= 8) {
        double v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = e < K ? w[e] : Num<double>::neg_inf();
        sort_desc_regs<double, 8>(v);
        const double shift = simplex_shift_sorted<double, 8, 1>(v, K, 0);
        for (int e = 0; e < K; ++e) {
            const double t = shift + w[e];
            w[e] = (t < 0.0) ? 0.0 : t;
        }
    } else {
        tiny_proj_simplex(w, scratch, K);
    }
}

// ---- building the sliced-ELL copies (once per problem handle) ------------------------------------------------
__global__ void sell_widths_kernel(const int64_t *__restrict__ ptr, int rows, int groups, int32_t *__restrict__ goff) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        int w = 0;
        for (int r = 32 * g; r < 32 * g + 32 && r < rows; ++r) w = max(w, (int)(ptr[r + 1] - ptr[r]));
        goff[g + 1] = 32 * ((w + 3) & ~3);  // widths are multiples of 4 (tiny_row_dot)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) goff[0] = 0;
}
__global__ void sell_scan_kernel(int32_t *goff, int groups) {  // a few hundred groups: one thread
    if (blockIdx.x || threadIdx.x) return;
    int run = 0;
    for (int g = 1; g <= groups; ++g) {
        run += goff[g];
        goff[g] = run;
    }
}
// pad = the column id padding entries carry (the slot after the last entry of the gathered vector, which holds 0.0)
__global__ void sell_fill_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ val, int rows,
                                 int groups, const int32_t *__restrict__ goff, int32_t *__restrict__ sidx, double *__restrict__ sval, int pad) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < 32 * groups; r += gridDim.x * blockDim.x) {
        const int g = r >> 5, lane = r & 31;
        const int base = goff[g] + lane, width = (goff[g + 1] - goff[g]) >> 5;
        const int64_t p0 = r < rows ? ptr[r] : 0, len = r < rows ? ptr[r + 1] - p0 : 0;
        for (int k = 0; k < width; ++k) {
            sidx[base + 32 * k] = 8 * (k < len ? idx[p0 + k] : pad);  // byte offset into the gathered vector
            if (sval) sval[base + 32 * k] = k < len ? val[p0 + k] : 0.0;
        }
    }
}

__global__ void __launch_bounds__(kTinyThreads, 1) solver_tiny_kernel(TinyArgs a, DevOpts o) {
    extern __shared__ __align__(16) double tiny_sm[];
    const int n = a.n, m = a.m, tid = threadIdx.x;
    // arena offsets (vectors carry one extra slot that holds 0.0: the target of padding entries)
    const int n1 = n + 1, m1 = m + 1;
    const int XS = 0, GS = 2 * n1, RS = 4 * n1, BS = 4 * n1 + 2 * m1, SC = BS + m1;  // x[2], g[2], r[2], b, scalars
    double *scal = tiny_sm + SC;
    __shared__ double s_red[32 * 5];
    __shared__ double s_out[5];
    __shared__ DevState s_state;  // the solver state lives on chip for the whole solve; copied out at the end
    DevState *st = &s_state;
    if (tid == 0) s_state = DevState{};
    for (int i = tid; i < n; i += kTinyThreads) tiny_sm[XS + i] = a.x[i];
    for (int i = tid; i < m; i += kTinyThreads) tiny_sm[BS + i] = a.b[i];
    if (tid < 2) {
        tiny_sm[XS + tid * n1 + n] = 0.0;
        tiny_sm[RS + tid * m1 + m] = 0.0;
    }
    if (tid < kScalCount) scal[tid] = 0.0;
    __syncthreads();

    // r = A x - b and the sums of EpiResidual (x, r, r_old: arena offsets; r_old < 0: none)
    auto residual = [&](int x, int r, int r_old) {
        double acc[3] = {0, 0, 0};
        constexpr int R = 1;  // measured: 2 rows in flight per thread are slower (32 warps already hide the latency)
        for (int row0 = tid; row0 < m; row0 += R * kTinyThreads) {
            double dot[R];
            tiny_rows_dot<R>(a.A, row0, m, tiny_sm, x, dot);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int row = row0 + q * kTinyThreads;
                if (row < m) {
                    const double v = dot[q] - tiny_sm[BS + row];
                    tiny_sm[r + row] = v;
                    residual_sums(v, r_old >= 0 ? tiny_sm[r_old + row] : 0.0, r_old >= 0, acc);
                }
            }
        }
        tiny_reduce<3, 0>(acc, s_red, s_out);
        if (tid == 0) {
            scal[kScalF] = 0.5 * s_out[0];
            scal[kScalRR] = s_out[0];
            scal[kScalRdr] = s_out[1];
            scal[kScalDrdr] = s_out[2];
        }
    };
    // g_new = A^T r and the sums of EpiGradBB (g < 0: only <g_new, g_new>)
    auto gradient = [&](int r, int g_new, int g, int x, int x_new) {
        double acc[5] = {0, 0, 0, 0, 0};
        constexpr int R = 1;
        for (int row0 = tid; row0 < n; row0 += R * kTinyThreads) {
            double dots[R];
            tiny_rows_dot<R>(a.AT, row0, n, tiny_sm, r, dots);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int row = row0 + q * kTinyThreads;
                if (row < n) {
                    const double dot = dots[q];
                    tiny_sm[g_new + row] = dot;
                    acc[3] += dot * dot;
                    if (g >= 0) {
                        const double go = tiny_sm[g + row], dx = tiny_sm[x_new + row] - tiny_sm[x + row], dg = dot - go;
                        acc[0] += dx * dg;
                        acc[1] += dg * dg;
                        acc[2] += go * dx;
                        acc[4] = fmax(acc[4], fabs(dx));
                    }
                }
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
    residual(XS, RS, -1);
    gradient(RS, GS, -1, -1, -1);
    if (tid == 0) decide_step(st, scal, nullptr, o, a.progress_f, a.progress_t, 1);
    __syncthreads();

    int cur = 0;
    while (!st->done) {
        const int nxt = cur ^ 1;
        const int xc = XS + cur * n1, xn = XS + nxt * n1, gc = GS + cur * n1, gn = GS + nxt * n1, rc = RS + cur * m1, rn = RS + nxt * m1;
        const double nt = -st->t;
        // ---- x_new = proj(x - t g): one thread per OD block ---------------------------------------
        for (int blk = tid; blk < a.nb; blk += kTinyThreads) {
            const int s = a.starts[blk], K = a.starts[blk + 1] - s;
            double *w = tiny_sm + xn + s;   // the trial point is formed in place; the trial gradient's slot is the sort scratch
            double sum = 0.0;
            for (int i = 0; i < K; ++i) {
                const double u = nt * tiny_sm[gc + s + i];  // np.add(x, -t*g, x_new): product and sum rounded separately
                double v = tiny_sm[xc + s + i] + u;
                if (a.proj_mode == 1) {                     // proj_multi_ball (proj_simplex.h:54-62)
                    if (v < 0.0)
                        v = 0.0;
                    else
                        sum += v;
                }
                w[i] = v;
            }
            if (a.proj_mode == 0 || sum > 1.0) tiny_project_block(w, tiny_sm + gn + s, K);
        }
        __syncthreads();
        // ---- objective and gradient at the trial point, decision, pull-back ------------------------------
        residual(xn, rn, rc);
        gradient(rn, gn, gc, xc, xn);
        if (tid == 0) decide_step(st, scal, nullptr, o, a.progress_f, a.progress_t, 0);
        __syncthreads();
        const double tau = st->tau;
        if (tau != 1.0) {  // commit_kernel's update on the shared-memory vectors
            const double c = 1.0 - tau;
            for (int i = tid; i < n; i += kTinyThreads) {
                if (tau == 0.0) {
                    tiny_sm[xn + i] = tiny_sm[xc + i];
                    tiny_sm[gn + i] = tiny_sm[gc + i];
                } else {
                    const double u = c * tiny_sm[xc + i];
                    tiny_sm[xn + i] = u + tau * tiny_sm[xn + i];
                    const double v = c * tiny_sm[gc + i];
                    tiny_sm[gn + i] = v + tau * tiny_sm[gn + i];
                }
            }
            for (int i = tid; i < m; i += kTinyThreads) {
                if (tau == 0.0) {
                    tiny_sm[rn + i] = tiny_sm[rc + i];
                } else {
                    const double u = c * tiny_sm[rc + i];
                    tiny_sm[rn + i] = u + tau * tiny_sm[rn + i];
                }
            }
            __syncthreads();
        }
        cur = nxt;
    }
    for (int i = tid; i < n; i += kTinyThreads) a.x[i] = tiny_sm[XS + cur * n1 + i];
    if (tid == 0) *a.st = s_state;
}

}  // namespace bsls
