/*
 * bsls_b200.h -- C ABI of libbsls_b200.so, the B200 (sm_100a) implementation of the
 * block-simplex least-squares inner loop.
 *
 * Two families of entry points:
 *
 *   bsls_<name>(...)        HOST buffers, blocking.  Argument lists are exactly those of
 *                           the reference's native functions (the ones its Cython layer
 *                           binds), so a maintainer swaps the `cdef extern` block and
 *                           nothing else.  Return value: status
 *                           (the reference's functions return void).
 *   bsls_dev_<name>(...)    DEVICE buffers, asynchronous on the given CUDA stream, no
 *                           allocation and no host synchronisation per call.  The block
 *                           layout is analysed once into a bsls_plan.
 *
 * "file:line" citations point into /root/reference.
 *
 * Block convention (python/c_extensions/proj_simplex.h:37-47, c_extensions.pyx:31-39):
 * `blocks` holds the FIRST index of every block, strictly increasing, blocks[0] >= 0,
 * blocks[numblocks-1] < n; the last block ends at n; entries before blocks[0] are left
 * untouched.
 *
 * All routines work in place, as the reference does.
 */
#ifndef BSLS_B200_H
#define BSLS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define BSLS_API __attribute__((visibility("default")))
#else
#define BSLS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status ---------------------------------------------------------------------- */
#define BSLS_OK            0
#define BSLS_ERR_ARG       1   /* precondition the reference asserts on (c_extensions.pyx:24,33-34) */
#define BSLS_ERR_CUDA      2   /* CUDA runtime error; see bsls_last_error() */
#define BSLS_ERR_NO_DEVICE 3   /* no sm_100 device: there is NO CPU fallback */
#define BSLS_ERR_ALLOC     4

BSLS_API const char *bsls_last_error(void);         /* text of the last failure on this thread */
BSLS_API const char *bsls_version(void);
BSLS_API int bsls_device_ok(void);                  /* BSLS_OK when a CUDA device of CC >= 10.0 is current */

typedef void *bsls_stream_t;               /* a cudaStream_t (0 = legacy default stream) */

/* ================================================================================== */
/* HOST-buffer entry points: drop-in for the reference's native layer                  */
/* ================================================================================== */

/* replaces proj_simplex            (python/c_extensions/proj_simplex.h:17-34) */
BSLS_API int bsls_proj_simplex(double *y, int start, int end);
/* replaces proj_multi_simplex      (python/c_extensions/proj_simplex.h:37-47) */
BSLS_API int bsls_proj_multi_simplex(double *y, const int *blocks, int numblocks, int n);
/* replaces proj_multi_ball         (python/c_extensions/proj_simplex.h:50-74) */
BSLS_API int bsls_proj_multi_ball(double *y, const int *blocks, int numblocks, int n);

/* replaces isotonic_regression     (python/c_extensions/isotonic_regression.h:13-58);
 * weight = pool sizes at the pool heads, in/out (all ones unless warm-starting); NULL means
 * "all ones in, result dropped", which is what the reference's Python layer does for
 * weight=None (c_extensions.pyx:70-71). */
BSLS_API int bsls_isotonic_regression(double *y, int start, int end, int *weight, int update);
/* replaces isotonic_regression_multi (isotonic_regression.h:85-92) */
BSLS_API int bsls_isotonic_regression_multi(double *y, const int *blocks, int numblocks, int n, int *weight, int update);
/* replace isotonic_regression_2 / _multi_2 (isotonic_regression.h:61-82,95-102) and
 * isotonic_regression_3 / _multi_3 (isotonic_regression.h:105-164): same regression, served
 * by the variant-1 kernel (values equal to ~1e-15 relative; weights are variant 1's). */
BSLS_API int bsls_isotonic_regression_2(double *y, int start, int end);
BSLS_API int bsls_isotonic_regression_multi_2(double *y, const int *blocks, int numblocks, int n);
BSLS_API int bsls_isotonic_regression_3(double *y, int start, int end, int *weight, int update);
BSLS_API int bsls_isotonic_regression_multi_3(double *y, const int *blocks, int numblocks, int n, int *weight, int update);

/* Pinned host memory for callers that want the host entry points to run at PCIe speed
 * (pageable buffers work too, through a staging copy). */
BSLS_API int bsls_host_alloc(void **ptr, int64_t bytes);
BSLS_API int bsls_host_free(void *ptr);

/* ================================================================================== */
/* DEVICE-buffer entry points                                                          */
/* ================================================================================== */

typedef struct bsls_plan bsls_plan;

/* Analyse a block layout once: validates it (the reference's asserts), detects uniform
 * block size, bins ragged blocks into tiles / large blocks, allocates scratch.
 * `d_blocks` is a DEVICE array of numblocks int32 start offsets; it is copied.
 * Synchronises `stream` once (layout statistics come back to the host). */
BSLS_API int bsls_plan_create(const int32_t *d_blocks, int numblocks, int n, bsls_stream_t stream, bsls_plan **out);
BSLS_API int bsls_plan_destroy(bsls_plan *plan);
/* layout facts: [0]=numblocks [1]=n [2]=first index [3]=uniform size or 0 [4]=min size
 * [5]=max size [6]=tiles [7]=large blocks */
BSLS_API int bsls_plan_info(const bsls_plan *plan, int64_t info[8]);

/* projections (a1-a3 of SURVEY section 8); fp32 twins are an extension */
BSLS_API int bsls_dev_proj_multi_simplex_f64(const bsls_plan *plan, double *y, bsls_stream_t stream);
BSLS_API int bsls_dev_proj_multi_ball_f64(const bsls_plan *plan, double *y, bsls_stream_t stream);
BSLS_API int bsls_dev_proj_multi_simplex_f32(const bsls_plan *plan, float *y, bsls_stream_t stream);
BSLS_API int bsls_dev_proj_multi_ball_f32(const bsls_plan *plan, float *y, bsls_stream_t stream);


/* segmented isotonic regression (a4-a6 of SURVEY section 8): weight may be NULL; clip01 != 0
 * fuses the [0,1] clamp the z-space projection applies afterwards (python/main.py:65,
 * python/algorithm_utils.py:223-224). */
BSLS_API int bsls_dev_isotonic_regression_multi_f64(const bsls_plan *plan, double *y, int32_t *weight, int update, int clip01, bsls_stream_t stream);
BSLS_API int bsls_dev_isotonic_regression_multi_f32(const bsls_plan *plan, float *y, int32_t *weight, int update, int clip01, bsls_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BSLS_B200_H */
