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
#define BSLS_ERR_UNSUPPORTED 5 /* valid input this revision does not serve (fp32 projection of a block > 8192 entries) */

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
 * isotonic_regression_3 / _multi_3 (isotonic_regression.h:105-164).  Both form their pool means in
 * another order than variant 1, and variant 3 leaves another weight array (tail markers w[k-1],
 * pools fused on ties): they run the reference's routine as written, one GPU thread per block, so
 * values and weights are the reference's bits (csrc/pava_seq.cuh).  Not the hot path. */
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
 * block size, bins ragged blocks into tiles / large blocks, allocates ALL scratch and auxiliary
 * streams (the plan is immutable afterwards).  `d_blocks` is a DEVICE array of numblocks int32
 * start offsets; it is copied.  Synchronises `stream` (layout statistics come back to the host).
 * A plan belongs to the device that was current at creation.  Calls on one plan share its scratch
 * (queue of dense blocks, fork/join streams): ONE STREAM AT A TIME PER PLAN -- give every stream
 * its own plan (the Python layer keys its plan cache on the stream). */
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
/* variants 2 and 3 on device buffers (see the host entry points above); variant 3 needs a weight array (ones for a cold call) */
BSLS_API int bsls_dev_isotonic_regression_multi_2_f64(const bsls_plan *plan, double *y, bsls_stream_t stream);
BSLS_API int bsls_dev_isotonic_regression_multi_3_f64(const bsls_plan *plan, double *y, int32_t *weight, int update, bsls_stream_t stream);


/* x <-> z change of variables (a7): z = running sums of every block without its last entry,
 * summed left to right; inverse = adjacent differences, last entry 1 - z_last.
 * replaces x2z_c / z2x_c (python/c_extensions/c_extensions.pyx:195-248).  plan->first must be 0. */
BSLS_API int bsls_dev_x2z_f64(const bsls_plan *plan, const double *x, double *z, bsls_stream_t stream);
BSLS_API int bsls_dev_z2x_f64(const bsls_plan *plan, double *x, const double *z, bsls_stream_t stream);
/* x = N z (+ x0 when add_x0 != 0) and zg = N^T v for the bidiagonal N of
 * python/bsls_utils.py:139-162 (x0 = particular_x0, :327-328); the z-space closures of
 * python/main.py:53-54 are built from these two and the SpMV pair below. */
BSLS_API int bsls_dev_nz_f64(const bsls_plan *plan, double *x, const double *z, int add_x0, bsls_stream_t stream);
BSLS_API int bsls_dev_ntv_f64(const bsls_plan *plan, double *zg, const double *v, bsls_stream_t stream);
/* y[block k] *= f[k]  (divide == 0)  or  /= f[k]  (divide != 0): the f-scaled projections of
 * get_solver_parts (python/algorithm_utils.py:232-265) */
BSLS_API int bsls_dev_block_scale_f64(const bsls_plan *plan, double *y, const double *f, int divide, bsls_stream_t stream);

/* ================================================================================== */
/* sparse least-squares objective, solver steps (a9-a17 of SURVEY section 8)           */
/* ================================================================================== */

/* Multi-GPU: OD blocks (columns of A) are sharded over ranks; the partial link vector A_p x_p
 * is summed with one ncclAllReduce per objective evaluation.  The communicator is NCCL's,
 * loaded at run time from `nccl_path` (NULL: "libnccl.so.2"). */
typedef struct bsls_comm bsls_comm;
BSLS_API int bsls_comm_unique_id(const char *nccl_path, char id[128]);
BSLS_API int bsls_comm_create(const char *nccl_path, int nranks, int rank, const char id[128], bsls_comm **out);
BSLS_API int bsls_comm_destroy(bsls_comm *comm);
BSLS_API int bsls_comm_allreduce_sum_f64(bsls_comm *comm, double *d_buf, int64_t count, bsls_stream_t stream);
/* Optional peer-memory exchange for the sharded solver loop (2..8 ranks of one NVSwitch domain): every rank allocates an
 * exchange region for link vectors of m entries (bsls_comm_p2p_alloc returns its 64-byte CUDA IPC handle), the caller
 * gathers the handles of all ranks (in rank order) and every rank maps them (bsls_comm_p2p_open).  The solver loop then
 * sums the link vector, subtracts b, forms the norms and distributes the result with ONE kernel over NVLink loads /
 * stores instead of ncclAllReduce + kernel, and exchanges the step scalars by NVLink stores instead of ncclAllGather.
 * Results are bit-identical on all ranks (sums in rank order).  Everything else keeps using NCCL. */
BSLS_API int bsls_comm_p2p_alloc(bsls_comm *comm, int64_t m, char handle[64]);
BSLS_API int bsls_comm_p2p_open(bsls_comm *comm, const char *handles);
BSLS_API int bsls_comm_p2p_ready(const bsls_comm *comm);
BSLS_API int bsls_comm_p2p_disable(bsls_comm *comm);   /* all ranks together, when one of them could not map its peers */

/* The problem 0.5 |A x - b|^2.  A is given as CSR with m rows and A^T as CSR with n rows -- the
 * two matrices the reference keeps (python/algorithm_utils.py:199-200).  All arrays are DEVICE
 * arrays owned by the caller and must outlive the handle.  a_val / at_val may be NULL: every
 * stored entry is 1 (route-link incidence, python/bsls_utils.py:494-507).  The handle owns an
 * m-vector for the residual, three n-vectors of solver workspace (allocated on first use) and
 * the reduction scratch.  One stream at a time per handle. */
typedef struct bsls_lsq bsls_lsq;
/* Reduction workspace: scratch of the deterministic grid reductions, a block of 16 device
 * scalars with a pinned host mirror, and (optionally) the communicator over which dot
 * products and maxima are summed.  Every bsls_lsq owns one (bsls_lsq_ws); solver drivers that
 * have no matrix create their own.  One stream at a time per workspace. */
typedef struct bsls_ws bsls_ws;
BSLS_API int bsls_ws_create(bsls_ws **out);
BSLS_API int bsls_ws_destroy(bsls_ws *ws);
BSLS_API int bsls_ws_set_comm(bsls_ws *ws, bsls_comm *comm);
BSLS_API double *bsls_ws_scalar_ptr(const bsls_ws *ws);           /* device pointer, 16 entries */
/* device scalars -> host (synchronises the stream): [0]=f [1]=<dx,dg> [2]=<dg,dg> [3]=<g,dx>
 * [4]=<g_new,g_new> [5]=max|dx| [6..9]=generic dots [10]=generic max [11]=<r,r> */
BSLS_API int bsls_ws_scalars(bsls_ws *ws, double out[16], bsls_stream_t stream);

BSLS_API int bsls_lsq_create(int64_t m, int64_t n, int64_t nnz,
                             const int64_t *a_ptr, const int32_t *a_idx, const double *a_val,
                             const int64_t *at_ptr, const int32_t *at_idx, const double *at_val,
                             const double *b, bsls_lsq **out);
BSLS_API int bsls_lsq_destroy(bsls_lsq *lsq);
BSLS_API int bsls_lsq_set_comm(bsls_lsq *lsq, bsls_comm *comm);   /* NULL: single GPU */
BSLS_API bsls_ws *bsls_lsq_ws(bsls_lsq *lsq);
BSLS_API int bsls_lsq_set_b(bsls_lsq *lsq, const double *b);
/* Optional column-panelled copy of A for problems whose x does not fit the 126 MB L2: the columns
 * are cut into `panels` contiguous slices, slice p is stored as a CSR matrix with m rows (column
 * ids global), and the `panels` matrices are stacked row-wise: ptr has panels*m + 1 entries.
 * A x is then formed panel after panel (the gathered slice of x stays L2-resident) and the
 * per-panel partial sums are added in ascending panel order.  panels <= 1 removes the copy. */
BSLS_API int bsls_lsq_set_panels(bsls_lsq *lsq, int panels, const int64_t *ptr, const int32_t *idx, const double *val);
/* kernel choice per side: 0 = from the mean row length, 1 = stream, 4/8/16/32 = lanes per row */
BSLS_API int bsls_lsq_set_modes(bsls_lsq *lsq, int a_mode, int at_mode);

/* replaces sparse_least_squares_obj (python/algorithm_utils.py:88-94): g <- A^T (A x - b),
 * *f_host = 0.5 |A x - b|^2.  Blocks until f is on the host. */
BSLS_API int bsls_lsq_obj_f64(bsls_lsq *lsq, const double *x, double *g, double *f_host, bsls_stream_t stream);
/* the same in two asynchronous halves; the residual stays inside the handle */
BSLS_API int bsls_dev_lsq_residual_f64(bsls_lsq *lsq, const double *x, bsls_stream_t stream);
BSLS_API int bsls_dev_lsq_gradient_f64(bsls_lsq *lsq, double *g, bsls_stream_t stream);
/* the gradient half fused with the step scalars of BATCH.solve_BB / line_search_np (python/BATCH.py:89,99-100,
 * python/algorithm_utils.py:120,126): g_new = A^T r, and scalar slots [1] <x_new - x, g_new - g>, [2] |g_new - g|^2,
 * [3] <g, x_new - x>, [4] |g_new|^2, [5] max |x_new - x| (this rank's share when sharded) */
BSLS_API int bsls_dev_lsq_gradient_bb_f64(bsls_lsq *lsq, double *g_new, const double *g, const double *x, const double *x_new,
                                          bsls_stream_t stream);
/* plain products: out = A v (m entries, summed over ranks) and out = A^T w (n entries) -- the
 * `linop` / `linop_T` of DORE.solve (python/DORE.py:6, python/gradient_descent.py:62-63) */
BSLS_API int bsls_dev_lsq_matvec_f64(bsls_lsq *lsq, const double *v, double *out, bsls_stream_t stream);
BSLS_API int bsls_dev_lsq_rmatvec_f64(bsls_lsq *lsq, const double *w, double *out, bsls_stream_t stream);
BSLS_API int bsls_lsq_scalars(bsls_lsq *lsq, double out[16], bsls_stream_t stream);
BSLS_API const double *bsls_lsq_residual_ptr(const bsls_lsq *lsq);   /* device pointer, m entries */
BSLS_API double *bsls_lsq_scalar_ptr(const bsls_lsq *lsq);           /* device pointer, 16 entries */

/* vector kernels of the solver drivers (all on device buffers) */
/* out = a x + b y, every product rounded separately (np.add(x, -t*g, x_new), python/BATCH.py:91) */
BSLS_API int bsls_dev_axpby_f64(double *out, double a, const double *x, double b, const double *y, int64_t n, bsls_stream_t stream);
/* count <= 4 dot products <x_k, y_k> in one pass (summed over ranks), and max |x_0 - y_0| when
 * want_max; blocking, results in out[0..3] and out[4] */
BSLS_API int bsls_ws_dots_f64(bsls_ws *ws, int count, const double *const *x, const double *const *y, int64_t n,
                               int want_max, double out[5], bsls_stream_t stream);
/* route-flow error metrics of LS_postprocess (python/main.py:112-134) for one iterate x_hat against x_true, one pass:
 * out[0] = sum |s (xt - xh)|, [1] = sum s xt, [2] = #{xt - xh > thresh}, [3] = |xt - xh|^2, [4] = max s (xt - xh)
 * (scaling NULL: ones).  Blocking, single GPU. */
BSLS_API int bsls_ws_flow_metrics_f64(bsls_ws *ws, const double *scaling, const double *x_true, const double *x_hat, int64_t n,
                                      double thresh, double out[5], bsls_stream_t stream);
/* d <- d + c v with c = scale * ((c0 ? *c0 : 1) - (c1 ? *c1 : 0)) read from DEVICE scalars (v NULL:
 * d <- c d), and *out (a DEVICE double, summed over ranks) <- <w, d> in the same pass (w NULL: no
 * dot).  Chains the L-BFGS two-loop recursion without the host (python/LBFGS.py:60-71,
 * python/BATCH.py:196-214). */
BSLS_API int bsls_dev_axpy_dot_f64(bsls_ws *ws, double *d, double scale, const double *c0, const double *c1, const double *v,
                                   const double *w, double *out, int64_t n, bsls_stream_t stream);
/* mirror-descent step: x_new = x * exp(-t g), every block divided by its sum; scalar slot 10 =
 * max |x_new - x|.  per_block_log == 0: t = step (python/BATCH.py:238-241);
 * per_block_log != 0: t = sqrt(2 ln K_block) / step (python/mirror_descent.py:26-28,39-47). */
BSLS_API int bsls_dev_md_update_f64(bsls_ws *ws, const bsls_plan *plan, double *x_new, const double *x, const double *g,
                                    double step, int per_block_log, bsls_stream_t stream);

/* The x-space BATCH solvers of the reference, run entirely by the library (no Python in the
 * loop): method 0 = solve (projected gradient, step 1/(min_eig*i+1)), 1 = solve_BB, 2 = solve_MD
 * (python/BATCH.py:7-106,217-250) with get_solver_parts' sparse objective, proj_multi_simplex_c
 * (proj_mode 0) or proj_multi_ball_c (1) and line_search_np (python/algorithm_utils.py:113-137). */
typedef struct {
    int method;          /* 0 projected gradient, 1 Barzilai-Borwein, 2 mirror descent, 5 L-BFGS (solve_LBFGS, python/BATCH.py:110-214) */
    int proj_mode;       /* 0 simplex, 1 l1-ball, 2 isotonic regression + clip to [0,1] (problem posed in z) */
    int use_line_search; /* BB always searches; solve() only when given one */
    int has_f_min;
    double f_min, opt_tol, prog_tol, min_eig;
    int max_iter;
    int corrections;     /* method 5 (L-BFGS): pairs kept, at most 64 (0: the reference's default 50) */
} bsls_batch_opts;
typedef struct {
    double f;
    int iterations;      /* the reference's counter `i` at exit */
    int stop_code;       /* 1 max_iter, 2 f - f_min < opt_tol, 3 |f_old - f| < prog_tol */
    double stop_value;   /* the number the reference formats into its stop string */
    int obj_evals, backtracks, kernel_launches;
    double device_ms;    /* CUDA-event time of the whole loop */
} bsls_batch_result;
BSLS_API int bsls_batch_solve_f64(bsls_lsq *lsq, const bsls_plan *plan, double *x, const bsls_batch_opts *opts,
                                  bsls_batch_result *res, double *progress_f, double *progress_t, int progress_cap,
                                  bsls_stream_t stream);

/* replaces mirror_descent.least_squares (python/mirror_descent.py:7-53): exponentiated gradient on the block simplices with
 * step sqrt(2 ln K_block) / (sqrt(k) Lf), at most `iters` updates, stops when max |x - x_prev| < tolerance.  x: in = the
 * starting point (1 / K_block), out = the result.  res->iterations = updates taken, stop_code 4 = tolerance reached,
 * stop_value = the last max |x - x_prev|.  The loop is device-resident (no host decision inside). */
BSLS_API int bsls_md_least_squares_f64(bsls_lsq *lsq, const bsls_plan *plan, double *x, int iters, double tolerance, double Lf,
                                       bsls_batch_result *res, bsls_stream_t stream);

/* replaces BB.solve (python/BB.py:7-45) for the z-space problem main.solve_in_z builds (python/main.py:47-65):
 * f(z) = 0.5 |A N z + target|^2 (the handle's b must hold -target), projection = isotonic regression of every z-block
 * clipped to [0, 1], stopping = solvers.stopping (python/solvers.py:40-63).  Runs iterations i_start + 1 .. i_end as a
 * device-resident loop (segments: the caller records a state every `record_every` iterations like the reference's log
 * callback).  xplan = blocks of x (defines N), zplan = blocks of z.  z, z_prev, g_prev (n - numblocks entries each) are
 * updated in place.  res->iterations = the reference's `i` at exit; stop_code 0 = segment finished, 1 = max_iter,
 * 5 = no change in gradient, 6 = norm(grad) too small; res->f = f at the returned z; stop_value = the last BB step. */
BSLS_API int bsls_zbb_run_f64(bsls_lsq *lsq, const bsls_plan *xplan, const bsls_plan *zplan, double *z, double *z_prev, double *g_prev,
                              int i_start, int i_end, int max_iter, double opt_tol, bsls_batch_result *res, bsls_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BSLS_B200_H */
