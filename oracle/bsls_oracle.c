/*
 * bsls_oracle.c -- CPU restatement of the reference's block-simplex hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline.
 *
 * Every function restates, in plain C and in its own words, the arithmetic of a
 * reference routine (paths relative to /root/reference):
 *
 *   orc_proj_simplex            python/c_extensions/proj_simplex.h:17-34
 *   orc_proj_multi_simplex      python/c_extensions/proj_simplex.h:37-47
 *   orc_proj_multi_ball         python/c_extensions/proj_simplex.h:50-74
 *   orc_pava                    python/c_extensions/isotonic_regression.h:13-58
 *   orc_pava2                   python/c_extensions/isotonic_regression.h:61-82
 *   orc_pava3                   python/c_extensions/isotonic_regression.h:105-155
 *   orc_pava_multi{,2,3}        python/c_extensions/isotonic_regression.h:85-102,157-164
 *   orc_x2z / orc_z2x           python/c_extensions/c_extensions.pyx:195-248
 *   orc_csr_matvec              scipy.sparse csr_matvec (called at
 *                               python/algorithm_utils.py:91-92)
 *   orc_lsq_obj                 python/algorithm_utils.py:88-94
 *
 * Parity pinning: tests/test_oracle.py checks every function here against
 *   (1) the golden vectors of the reference's own tests
 *       (tests/fast/test_proj_simplex.py:24-52,77-81, test_c_extensions.py:67-79,
 *        isotonic_regression.h:180,187,194 comments), committed in tests/golden/,
 *   (2) oracle/_ref/libbsls_ref.so, the reference's unmodified C++ headers built
 *       by oracle/Makefile, when that library is present, on seeded random input,
 *   (3) fixtures produced by running the reference's own Python/Cython path
 *       (tests/golden/make_golden.py).
 *
 * Floating point: the order of every sum, product and quotient follows the
 * reference statement by statement so that results are bit-identical to
 * g++ -O2 on x86-64 (no FMA contraction: build with -ffp-contract=off).
 */

#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* simplex projection                                                        */
/* ------------------------------------------------------------------------ */

static int cmp_desc(const void *pa, const void *pb)
{
    double a = *(const double *)pa, b = *(const double *)pb;
    return (a < b) - (a > b);
}

/* One block: y[lo:hi) -> Euclidean projection on {x >= 0, sum x = 1}.
 * The shift is the LAST candidate (1 - prefix_k)/k, k = 1..K, whose sorted
 * element stays positive after shifting; prefix sums run left to right over
 * the descending-sorted copy.  `scratch` must hold hi-lo doubles. */
static void project_block(double *y, int64_t lo, int64_t hi, double *scratch)
{
    int64_t K = hi - lo;
    if (K <= 0) return;
    memcpy(scratch, y + lo, (size_t)K * sizeof(double));
    qsort(scratch, (size_t)K, sizeof(double), cmp_desc);

    double running = scratch[0];
    double shift = 1. - running;
    for (int64_t k = 1; k < K; ++k) {
        running += scratch[k];
        double cand = (1. - running) / ((double)k + 1.);
        if (scratch[k] + cand > 0) shift = cand;
    }
    for (int64_t k = lo; k < hi; ++k) {
        double v = shift + y[k];
        y[k] = (v < 0.) ? 0. : v;      /* std::max(v, 0.) keeps v unless v < 0 */
    }
}

void orc_proj_simplex(double *y, int64_t lo, int64_t hi)
{
    if (hi <= lo) return;
    double *scratch = (double *)malloc((size_t)(hi - lo) * sizeof(double));
    project_block(y, lo, hi, scratch);
    free(scratch);
}

static int64_t widest_block(const int32_t *starts, int64_t nb, int64_t n)
{
    int64_t w = 0;
    for (int64_t b = 0; b < nb; ++b) {
        int64_t hi = (b + 1 < nb) ? starts[b + 1] : n;
        if (hi - starts[b] > w) w = hi - starts[b];
    }
    return w;
}

/* starts[b] = first index of block b; the last block ends at n; entries in
 * front of starts[0] are left alone. */
void orc_proj_multi_simplex(double *y, const int32_t *starts, int64_t nb, int64_t n)
{
    if (nb <= 0) return;
    double *scratch = (double *)malloc((size_t)(widest_block(starts, nb, n) + 1) * sizeof(double));
    for (int64_t b = 0; b < nb; ++b) {
        int64_t hi = (b + 1 < nb) ? starts[b + 1] : n;
        project_block(y, starts[b], hi, scratch);
    }
    free(scratch);
}

/* "lasso" variant: clip negatives, and project only if the clipped block sums
 * to more than one.  The sum skips the clipped entries and runs left to right. */
static void ball_block(double *y, int64_t lo, int64_t hi, double *scratch)
{
    double total = 0.0;
    for (int64_t k = lo; k < hi; ++k) {
        if (y[k] < 0.0) y[k] = 0.0;
        else total += y[k];
    }
    if (total > 1.0) project_block(y, lo, hi, scratch);
}

void orc_proj_multi_ball(double *y, const int32_t *starts, int64_t nb, int64_t n)
{
    if (nb <= 0) return;
    double *scratch = (double *)malloc((size_t)(widest_block(starts, nb, n) + 1) * sizeof(double));
    for (int64_t b = 0; b < nb; ++b) {
        int64_t hi = (b + 1 < nb) ? starts[b + 1] : n;
        ball_block(y, starts[b], hi, scratch);
    }
    free(scratch);
}

/* Same arithmetic, blocks spread over host threads (blocks are independent, so
 * the result is identical to the serial loop).  Used only as the "best-effort
 * CPU, N cores" timing column. */
void orc_proj_multi_simplex_mt(double *y, const int32_t *starts, int64_t nb, int64_t n, int threads)
{
    if (nb <= 0) return;
    int64_t w = widest_block(starts, nb, n) + 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
    {
        double *scratch = (double *)malloc((size_t)w * sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t b = 0; b < nb; ++b) {
            int64_t hi = (b + 1 < nb) ? starts[b + 1] : n;
            project_block(y, starts[b], hi, scratch);
        }
        free(scratch);
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* isotonic regression (pool adjacent violators), three variants             */
/* ------------------------------------------------------------------------ */

/* Variant 1 (canonical pool structure).  sz[h] is the size of the pool whose
 * head is h; callers pass all ones unless warm-starting.  Sweeps repeat until a
 * sweep pools nothing.  In one sweep, starting at a head, follow heads while the
 * next head's value does not exceed the current one; if the first and last value
 * of that run differ the run collapses into its size-weighted mean (summed left
 * to right).  With spread != 0 the head value is copied over its pool at the end. */
void orc_pava(double *y, int64_t lo, int64_t hi, int32_t *sz, int spread)
{
    for (;;) {
        int merged_any = 0;
        int64_t head = lo;
        while (head < hi) {
            int64_t last = head;
            int64_t nxt = head + sz[head];
            while (nxt < hi && y[nxt] <= y[last]) {
                last = nxt;
                nxt += sz[nxt];
            }
            if (y[head] != y[last]) {
                double acc = 0.0;
                int32_t cnt = 0;
                for (int64_t p = head; p < nxt; p += sz[p]) {
                    acc += y[p] * sz[p];
                    cnt += sz[p];
                }
                y[head] = acc / cnt;
                sz[head] = cnt;
                merged_any = 1;
            }
            head = nxt;
        }
        if (!merged_any) break;
    }
    if (spread) {
        for (int64_t head = lo; head < hi; head += sz[head]) {
            int64_t stop = head + sz[head];
            for (int64_t p = head + 1; p < stop; ++p) y[p] = y[head];
        }
    }
}

/* Variant 2: no size array; every sweep rewrites each non-increasing stretch
 * (first != last) with its plain mean. */
void orc_pava2(double *y, int64_t lo, int64_t hi)
{
    int64_t top = hi - 1;
    for (;;) {
        int merged_any = 0;
        int64_t a = lo;
        while (a < top) {
            int64_t z = a;
            while (z < top && y[z] >= y[z + 1]) ++z;
            if (y[a] != y[z]) {
                double acc = 0.0;
                for (int64_t p = a; p <= z; ++p) acc += y[p];
                double mean = acc / (z + 1 - a);
                for (int64_t p = a; p <= z; ++p) y[p] = mean;
                merged_any = 1;
            }
            a = z + 1;
        }
        if (!merged_any) break;
    }
}

/* Variant 3: single forward pass; after collapsing a run it walks back over the
 * pools on its left while they are >= the new pool, folding them in pairwise.
 * sz[tail of pool] mirrors the pool size so the walk can find the previous head. */
void orc_pava3(double *y, int64_t lo, int64_t hi, int32_t *sz, int spread)
{
    int64_t head = lo;
    while (head < hi) {
        int64_t last = head;
        int64_t nxt = head + sz[head];
        while (nxt < hi && y[nxt] <= y[last]) {
            last = nxt;
            nxt += sz[nxt];
        }
        if (y[head] != y[last]) {
            double acc = 0.0;
            int32_t cnt = 0;
            for (int64_t p = head; p < nxt; p += sz[p]) {
                acc += y[p] * sz[p];
                cnt += sz[p];
            }
            y[head] = acc / cnt;
            sz[head] = cnt;
            sz[nxt - 1] = cnt;
            if (head > lo) {
                int64_t prev = head - sz[head - 1];
                while (prev >= lo && y[prev] >= y[head]) {
                    y[prev] = (sz[head] * y[head] + sz[prev] * y[prev]) / (sz[head] + sz[prev]);
                    sz[prev] = sz[head] + sz[prev];
                    head = prev;
                    if (prev == lo) break;
                    prev -= sz[prev - 1];
                }
                sz[nxt - 1] = sz[head];
            }
        } else {
            head = nxt;
        }
    }
    if (spread) {
        for (int64_t h = lo; h < hi; h += sz[h]) {
            int64_t stop = h + sz[h];
            for (int64_t p = h + 1; p < stop; ++p) y[p] = y[h];
        }
    }
}

void orc_pava_multi(double *y, const int32_t *starts, int64_t nb, int64_t n, int32_t *sz, int spread)
{
    for (int64_t b = 0; b < nb; ++b)
        orc_pava(y, starts[b], (b + 1 < nb) ? starts[b + 1] : n, sz, spread);
}

void orc_pava_multi2(double *y, const int32_t *starts, int64_t nb, int64_t n)
{
    for (int64_t b = 0; b < nb; ++b)
        orc_pava2(y, starts[b], (b + 1 < nb) ? starts[b + 1] : n);
}

void orc_pava_multi3(double *y, const int32_t *starts, int64_t nb, int64_t n, int32_t *sz, int spread)
{
    for (int64_t b = 0; b < nb; ++b)
        orc_pava3(y, starts[b], (b + 1 < nb) ? starts[b + 1] : n, sz, spread);
}

void orc_pava_multi_mt(double *y, const int32_t *starts, int64_t nb, int64_t n, int32_t *sz, int spread, int threads)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t b = 0; b < nb; ++b)
        orc_pava(y, starts[b], (b + 1 < nb) ? starts[b + 1] : n, sz, spread);
}

/* clip to [0, 1] (python/main.py:65, python/algorithm_utils.py:223-224) */
void orc_clip01(double *y, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double v = y[i];
        v = v > 0. ? v : 0.;   /* np.maximum(0., x) */
        v = v < 1. ? v : 1.;   /* np.minimum(1., x) */
        y[i] = v;
    }
}

/* ------------------------------------------------------------------------ */
/* x <-> z change of variables                                               */
/* ------------------------------------------------------------------------ */

/* z = per-block running sums of x without each block's last entry;
 * z has n - nb entries.  starts[0] must be 0. */
void orc_x2z(const double *x, double *z, const int32_t *starts, int64_t nb, int64_t n)
{
    int64_t out = 0;
    for (int64_t b = 0; b < nb; ++b) {
        int64_t hi = (b + 1 < nb) ? starts[b + 1] : n;
        double run = 0.0;
        for (int64_t i = starts[b]; i < hi - 1; ++i) {
            run += x[i];
            z[out++] = run;
        }
    }
}

/* inverse: adjacent differences, and the last entry of a block is 1 - z_last. */
void orc_z2x(double *x, const double *z, const int32_t *starts, int64_t nb, int64_t n)
{
    int64_t in = 0;
    for (int64_t b = 0; b < nb; ++b) {
        int64_t hi = (b + 1 < nb) ? starts[b + 1] : n;
        double before = 0.0;
        for (int64_t i = starts[b]; i < hi - 1; ++i) {
            x[i] = z[in] - before;
            before = z[in];
            ++in;
        }
        x[hi - 1] = 1.0 - before;
    }
}

/* ------------------------------------------------------------------------ */
/* sparse least-squares objective and gradient                               */
/* ------------------------------------------------------------------------ */

/* out = M v for a CSR matrix, each row summed left to right (scipy csr_matvec). */
void orc_csr_matvec(int64_t rows, const int64_t *ptr, const int32_t *idx, const double *val,
                    const double *v, double *out)
{
    for (int64_t r = 0; r < rows; ++r) {
        double acc = 0.0;
        for (int64_t p = ptr[r]; p < ptr[r + 1]; ++p) acc += val[p] * v[idx[p]];
        out[r] = acc;
    }
}

/* res = A x - b ; g = A^T res ; returns 0.5 * <res, res>.
 * A is given as CSR (m rows) and A^T as a second CSR (n rows), exactly the two
 * matrices the reference keeps (python/algorithm_utils.py:199-200). */
double orc_lsq_obj(int64_t m, int64_t n,
                   const int64_t *a_ptr, const int32_t *a_idx, const double *a_val,
                   const int64_t *at_ptr, const int32_t *at_idx, const double *at_val,
                   const double *x, const double *b, double *res, double *g)
{
    orc_csr_matvec(m, a_ptr, a_idx, a_val, x, res);
    for (int64_t i = 0; i < m; ++i) res[i] -= b[i];
    orc_csr_matvec(n, at_ptr, at_idx, at_val, res, g);
    double f = 0.0;
    for (int64_t i = 0; i < m; ++i) f += res[i] * res[i];   /* order differs from BLAS ddot: compare to 1e-12 rel */
    return .5 * f;
}
