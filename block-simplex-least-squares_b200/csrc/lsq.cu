// lsq.cu -- host side of the sparse least-squares path: the bsls_lsq handle, kernel
// selection, the NCCL all-reduce of the link vector, and the x-space BATCH solver loop
// (python/BATCH.py:7-106,217-250 of the reference) run without the interpreter.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#include "kernels.h"
#include "lsq.cuh"
#include "p2p.cuh"
#include "solver_cluster.cuh"
#include "solver_tiny.cuh"

using namespace bsls;

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (torch ships libnccl.so.2; no link-time dependency)
// ---------------------------------------------------------------------------------------------
namespace {
struct Id128 {  // ncclUniqueId, passed by value
    char b[128];
};
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

int nccl_load(const char *path) {
    if (g_nccl.handle) return BSLS_OK;
    const char *names[] = {path, "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cannot load NCCL (%s)", dlerror());
        return BSLS_ERR_ARG;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        set_error("NCCL library lacks the expected symbols");
        return BSLS_ERR_ARG;
    }
    g_nccl.handle = h;
    return BSLS_OK;
}
#define BSLS_NCCL_TRY(expr)                                                                          \
    do {                                                                                             \
        int _r = (expr);                                                                             \
        if (_r != 0) {                                                                               \
            set_error("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, _r,               \
                      g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                      \
            return BSLS_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)
}  // namespace

struct bsls_comm {
    void *comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory exchange (p2p.cuh): this rank's region, the mapped regions of the others, the epoch of the last evaluation
    void *region = nullptr;
    unsigned long long *d_prof = nullptr;  // BSLS_P2P_PROF=1: wait / transfer times of the exchange kernels (development)
    void *peer_base[kP2pMaxRanks] = {};
    int64_t p2p_m = 0;
    bool p2p_ready = false;
    P2pView view{};
    unsigned long long epoch = 0;
    unsigned *ticket = nullptr;
    double *cta_partials = nullptr;
    int reduce_grid = 0;
};

// layout of an exchange region, in 8-byte words: partial[m] | rfull0[m] | rfull1[m] | slots[P*8] | gath[P*8] | flags[3*P]
static void p2p_fill_view(bsls_comm *c) {
    const int P = c->nranks;
    const int64_t m = (c->p2p_m + 1) & ~int64_t(1);
    for (int q = 0; q < P; ++q) {
        double *base = reinterpret_cast<double *>(q == c->rank ? c->region : c->peer_base[q]);
        c->view.partial[q] = base;
        c->view.rfull[q][0] = base + m;
        c->view.rfull[q][1] = base + 2 * m;
        c->view.slots[q] = base + 3 * m;
        c->view.gath[q] = base + 3 * m + P * 8;
        c->view.flags[q] = reinterpret_cast<unsigned long long *>(base + 3 * m + 2 * P * 8);
    }
    c->view.nranks = P;
    c->view.rank = c->rank;
    c->view.prof = c->d_prof;
}
static size_t p2p_region_bytes(int P, int64_t m) { return sizeof(double) * (size_t)(3 * ((m + 1) & ~int64_t(1)) + 2 * P * 8 + 3 * P + 8); }

// reduction scratch + device/host scalar block + communicator: everything a reducing kernel needs
struct bsls_ws {
    RedCtx red{};
    double *d_scal = nullptr, *h_scal = nullptr;  // kScalCount each (h_scal pinned)
    bsls_comm *comm = nullptr;
    int launches = 0;
    // device-resident solver loop
    DevState *d_state = nullptr, *h_state = nullptr;  // h_state: 2 pinned copies (one per iteration in flight)
    cudaEvent_t ev_state[2] = {nullptr, nullptr};
    double *d_gather = nullptr;                        // nranks * kStepScalars doubles (all-gather of the per-rank step scalars)
    int gather_cap = 0;
    double *d_prog = nullptr;                          // 2 * prog_cap doubles: objective and device time stamp per iteration
    int prog_cap = 0;
};

struct bsls_lsq {
    int64_t m = 0, n = 0, nnz = 0;
    const int64_t *a_ptr = nullptr, *t_ptr = nullptr;
    const int32_t *a_idx = nullptr, *t_idx = nullptr;
    const double *a_val = nullptr, *t_val = nullptr;
    const double *b = nullptr;
    int a_mode = 0, t_mode = 0;
    // column-panelled copy of A (optional): `panels` CSR matrices of m rows stacked into one of
    // panels*m rows; panel p holds the columns of its slice, so the gathered part of x stays in L2
    int panels = 0;
    const int64_t *p_ptr = nullptr;
    const int32_t *p_idx = nullptr;
    const double *p_val = nullptr;
    double *partial = nullptr;                    // panels * m (owned)
    int p_mode = 0;
    double *r = nullptr;                          // m
    double *r2 = nullptr;                         // m: residual of the trial point (solver loop; allocated with the workspace)
    int t_ell = 0, a_ell = 0;                     // common row length of A^T / A when every row has the same (1..16 supported), else 0
    double *wg = nullptr, *wxn = nullptr, *wgn = nullptr;  // n each, solver workspace
    // sliced-ELL copies of A and A^T for the single-CTA solver (built on first use; owned)
    int32_t *sell_idx[2] = {nullptr, nullptr}, *sell_goff[2] = {nullptr, nullptr};
    double *sell_val[2] = {nullptr, nullptr};
    bool sell_ready = false;
    std::vector<int32_t> sell_hgoff[2];  // host copies of the group offsets (the cluster solver sizes its shares from them)
    double *wz = nullptr;                         // n: one more vector for the z-space BB loop (allocated on first use)
    bsls_ws *ws = nullptr;                        // owned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

// ptr[i] == i * L for all rows?  (one reduction over the pointer array, at handle creation)
__global__ void uniform_rows_kernel(const int64_t *__restrict__ ptr, int64_t rows, int64_t L, int *bad) {
    int b = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= rows; i += (int64_t)gridDim.x * blockDim.x)
        if (ptr[i] != i * L) b = 1;
    if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}
int uniform_row_length(const int64_t *d_ptr, int64_t rows, int64_t nnz) {
    if (rows <= 0 || nnz <= 0 || nnz % rows) return 0;
    const int64_t L = nnz / rows;
    if (L < 1 || L > 64) return 0;
    int *d_bad = nullptr, h_bad = 1;
    if (cudaMalloc(&d_bad, sizeof(int)) != cudaSuccess) return 0;
    cudaMemset(d_bad, 0, sizeof(int));
    int64_t want = (rows + 1 + 255) / 256;
    uniform_rows_kernel<<<(int)(want < 1184 ? want : 1184), 256>>>(d_ptr, rows, L, d_bad);
    if (cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) h_bad = 1;
    cudaFree(d_bad);
    return h_bad ? 0 : (int)L;
}

// every stored value exactly 1.0?  (route-link incidence, python/bsls_utils.py:494-507): the products then skip the value
// arrays altogether -- 1.0 * w is exact, so the results keep their bits and two thirds of the matrix traffic go away
__global__ void all_ones_kernel(const double *__restrict__ val, int64_t nnz, int *bad) {
    int b = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        if (val[i] != 1.0) b = 1;
    if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}
bool all_ones(const double *d_val, int64_t nnz) {
    if (!d_val || nnz <= 0) return false;
    int *d_bad = nullptr, h_bad = 1;
    if (cudaMalloc(&d_bad, sizeof(int)) != cudaSuccess) return false;
    cudaMemset(d_bad, 0, sizeof(int));
    int64_t want = (nnz + 255) / 256;
    all_ones_kernel<<<(int)(want < 1184 ? want : 1184), 256>>>(d_val, nnz, d_bad);
    if (cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) h_bad = 1;
    cudaFree(d_bad);
    return !h_bad;
}

// BSLS_SPMV_V8=0 keeps the round-1 inner loop (4 gathers per lane and pass + remainder loop) for A/B measurements
inline bool vector8_on() {
    const char *e = getenv("BSLS_SPMV_V8");
    return !(e && atoi(e) == 0);
}

int pick_mode(int64_t nnz, int64_t rows) {
    const double avg = rows > 0 ? (double)nnz / (double)rows : 0.0;
    if (avg <= 32.0) return 1;
    if (avg < 64.0) return vector8_on() ? 4 : 8;  // spmv_vector8_kernel: 4 lanes x 8 entries cover a ~40-50-entry row piece in two passes (measured: 0.726 against 0.754 ms for 1.6e8 entries)
    if (avg < 128.0) return 16;
    return 32;
}

int grid_elems(int64_t n) {
    int64_t want = (n + 256 * 4 - 1) / (256 * 4);
    if (want < 1) want = 1;
    return (int)(want < kRedMaxGrid ? want : kRedMaxGrid);
}

// Grid of a persistent kernel: every CTA resident at once (one wave), never more than the
// reduction scratch holds.
template <class K> int resident_grid(K kern, int threads) {
    int dev = 0, sms = num_sms(), per = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, threads, 0) != cudaSuccess || per < 1) per = 1;
    const int g = sms * per;
    return g < kRedMaxGrid ? g : kRedMaxGrid;
}

template <class Epi, int LANES>
int launch_vector(bsls_ws *w, int64_t rows, const int64_t *ptr, const int32_t *idx, const double *val, const double *v,
                  const Epi &epi, cudaStream_t st, const int *skip) {
    constexpr int T = 256;
    int64_t want = (rows * LANES + T - 1) / T;
    if (vector8_on()) {
        static thread_local PerDevice<int> full8_pd;
        int &full8 = full8_pd.get(0);
        if (!full8) full8 = resident_grid(spmv_vector8_kernel<Epi, T, LANES>, T);
        const int grid = (int)(want < full8 ? (want < 1 ? 1 : want) : full8);
        spmv_vector8_kernel<Epi, T, LANES><<<grid, T, 0, st>>>(rows, ptr, idx, val, v, epi, w->red, skip);
        return 0;
    }
    static thread_local PerDevice<int> full_pd;
    int &full = full_pd.get(0);
    if (!full) full = resident_grid(spmv_vector_kernel<Epi, T, LANES>, T);
    const int grid = (int)(want < full ? (want < 1 ? 1 : want) : full);
    spmv_vector_kernel<Epi, T, LANES><<<grid, T, 0, st>>>(rows, ptr, idx, val, v, epi, w->red, skip);
    return 0;
}

// rows of exactly L entries: no pointers, no staging (spmv_ell_kernel)
template <class Epi, int L>
int launch_ell(bsls_ws *w, int64_t rows, const int32_t *idx, const double *val, const double *v, const Epi &epi, cudaStream_t st,
               const int *skip) {
    constexpr int T = 256;
#define ELL_GO(HV)                                                                                                 \
    {                                                                                                              \
        static thread_local PerDevice<int> full_pd;                                                                \
        int &full = full_pd.get(0);                                                                                \
        if (!full) full = resident_grid(spmv_ell_kernel<Epi, T, L, HV>, T);                                        \
        const int64_t want = (rows + T - 1) / T;                                                                   \
        spmv_ell_kernel<Epi, T, L, HV><<<(int)(want < full ? want : full), T, 0, st>>>(rows, idx, val, v, epi, w->red, skip); \
    }
    if (val) ELL_GO(true)
    else ELL_GO(false)
#undef ELL_GO
    return 0;
}

inline bool ell_supported(int L) { return L == 4 || L == 6 || L == 8 || L == 10 || L == 12 || L == 16; }

// mode: 1 = stream, 2 = ELL (ell = common row length), 4/8/16/32 = lanes per row
template <class Epi>
int launch_spmv(bsls_ws *w, int mode, int64_t rows, const int64_t *ptr, const int32_t *idx, const double *val, const double *v,
                const Epi &epi, cudaStream_t st, const int *skip = nullptr, int ell = 0) {
    constexpr int T = 256;
    if (rows <= 0) return BSLS_OK;
    if (mode == 2 && ell_supported(ell)) {
        switch (ell) {
            case 4: launch_ell<Epi, 4>(w, rows, idx, val, v, epi, st, skip); break;
            case 6: launch_ell<Epi, 6>(w, rows, idx, val, v, epi, st, skip); break;
            case 8: launch_ell<Epi, 8>(w, rows, idx, val, v, epi, st, skip); break;
            case 10: launch_ell<Epi, 10>(w, rows, idx, val, v, epi, st, skip); break;
            case 12: launch_ell<Epi, 12>(w, rows, idx, val, v, epi, st, skip); break;
            default: launch_ell<Epi, 16>(w, rows, idx, val, v, epi, st, skip); break;
        }
    } else if (mode == 1 || mode == 2) {
        static thread_local PerDevice<int> full_pd;
        int &full = full_pd.get(0);
        if (!full) full = resident_grid(spmv_stream_kernel<Epi, T, 2048>, T);
        const int64_t tiles = (rows + T - 1) / T;
        const int grid = (int)(tiles < full ? tiles : full);
        spmv_stream_kernel<Epi, T, 2048><<<grid, T, 0, st>>>(rows, ptr, idx, val, v, epi, w->red, skip);
    } else {
        switch (mode) {
            case 4: launch_vector<Epi, 4>(w, rows, ptr, idx, val, v, epi, st, skip); break;
            case 8: launch_vector<Epi, 8>(w, rows, ptr, idx, val, v, epi, st, skip); break;
            case 16: launch_vector<Epi, 16>(w, rows, ptr, idx, val, v, epi, st, skip); break;
            default: launch_vector<Epi, 32>(w, rows, ptr, idx, val, v, epi, st, skip); break;
        }
    }
    BSLS_LAUNCH_CHECK();
    w->launches++;
    return BSLS_OK;
}

int allreduce(bsls_ws *w, double *buf, int64_t count, int op, cudaStream_t st) {
    if (!w->comm || w->comm->nranks <= 1) return BSLS_OK;
    BSLS_NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat64, op, w->comm->comm, st));
    return BSLS_OK;
}

// A^T w with an epilogue, on whichever kernel the side was given
template <class Epi> int launch_at(bsls_lsq *q, const double *w, const Epi &epi, cudaStream_t st, const int *skip = nullptr) {
    return launch_spmv(q->ws, q->t_mode, q->n, q->t_ptr, q->t_idx, q->t_val, w, epi, st, skip, q->t_ell);
}

// the sharded solver loop exchanges over peer memory when the communicator has its regions mapped for this m
bool use_p2p(const bsls_lsq *q) {
    const bsls_comm *c = q->ws->comm;
    return c && c->nranks > 1 && c->p2p_ready && c->p2p_m == q->m;
}

// r = A x - b (summed over ranks), scalar F = 0.5 <r, r>; with r_old also <r_old, r - r_old> and |r - r_old|^2.
// p2p_nxt >= 0 (solver loop, peer-memory build): r is residual buffer `p2p_nxt` of the exchange region, r_old the other
// one; the sum over ranks, - b, the three sums and the distribution to all ranks are ONE kernel over NVLink (p2p.cuh).
int residual(bsls_lsq *q, const double *x, double *r, const double *b, cudaStream_t st, const double *r_old = nullptr,
             const int *skip = nullptr, int p2p_nxt = -1) {
    bsls_ws *w = q->ws;
    const bool dist = w->comm && w->comm->nranks > 1;
    const bool p2p = dist && p2p_nxt >= 0 && b && use_p2p(q);
    double *local = p2p ? w->comm->view.partial[w->comm->rank] : r;
    if (q->panels > 1) {
        // one launch per panel: the kernel boundary keeps all CTAs inside the same slice of x, which
        // therefore stays L2-resident (a single launch lets CTAs drift several panels apart)
        for (int p = 0; p < q->panels; ++p) {
            EpiResidual epi{q->partial + (size_t)p * q->m, nullptr, nullptr};
            if (int rc = launch_spmv(w, q->p_mode, q->m, q->p_ptr + (size_t)p * q->m, q->p_idx, q->p_val, x, epi, st, skip)) return rc;
        }
        panel_reduce_kernel<<<grid_elems(q->m), 256, 0, st>>>(local, q->partial, dist ? nullptr : b, r_old, q->m, q->panels, w->red, skip);
        BSLS_LAUNCH_CHECK();
        w->launches++;
    } else {
        const bool fin = !dist && b;
        EpiResidual epi{local, fin ? b : nullptr, fin ? r_old : nullptr};
        if (int rc = launch_spmv(w, q->a_mode, q->m, q->a_ptr, q->a_idx, q->a_val, x, epi, st, skip, q->a_ell)) return rc;
    }
    if (p2p) {
        bsls_comm *c = w->comm;
        ++c->epoch;
        p2p_reduce_kernel<<<c->reduce_grid, kP2pThreads, 0, st>>>(c->view, b, q->m, p2p_nxt, c->epoch, c->ticket, c->cta_partials, skip);
        BSLS_LAUNCH_CHECK();
        p2p_wait_kernel<<<1, 32, 0, st>>>(c->view, c->epoch, w->d_scal, skip);
        BSLS_LAUNCH_CHECK();
        w->launches += 2;
    } else if (dist) {
        if (int rc = allreduce(w, r, q->m, kNcclSum, st)) return rc;
        if (b) {
            residual_finish_kernel<<<grid_elems(q->m), 256, 0, st>>>(r, b, r_old, q->m, w->red, skip);
            BSLS_LAUNCH_CHECK();
            w->launches++;
        }
    }
    return BSLS_OK;
}

int ensure_workspace(bsls_lsq *q) {
    if (q->wg) return BSLS_OK;
    BSLS_CUDA_TRY(cudaMalloc(&q->r2, sizeof(double) * (size_t)q->m));
    BSLS_CUDA_TRY(cudaMalloc(&q->wg, sizeof(double) * (size_t)q->n));
    BSLS_CUDA_TRY(cudaMalloc(&q->wxn, sizeof(double) * (size_t)q->n));
    BSLS_CUDA_TRY(cudaMalloc(&q->wgn, sizeof(double) * (size_t)q->n));
    return BSLS_OK;
}

int fetch_scalars(bsls_ws *q, cudaStream_t st) {
    BSLS_CUDA_TRY(cudaMemcpyAsync(q->h_scal, q->d_scal, sizeof(double) * kScalCount, cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    return BSLS_OK;
}

BlockLayout layout_of(const bsls_plan *p) {
    BlockLayout l;
    l.starts = p->d_starts;
    l.nb = p->nb;
    l.uniform = p->uniform;
    l.first = p->first;
    return l;
}

int lanes_for(const bsls_plan *p) {
    const int k = p->uniform > 0 ? p->uniform : (int)((p->n - p->first) / (p->nb > 0 ? p->nb : 1));
    if (k <= 4) return 4;
    if (k <= 8) return 8;
    if (k <= 16) return 16;
    return 32;
}

int grid_groups(int nb, int lanes) {
    int64_t want = ((int64_t)nb * lanes + 255) / 256;
    if (want < 1) want = 1;
    return (int)(want < kRedMaxGrid ? want : kRedMaxGrid);
}

#define DISPATCH_LANES(lanes, KERNEL, grid, st, ...)                         \
    switch (lanes) {                                                         \
        case 4: KERNEL<4><<<grid, 256, 0, st>>>(__VA_ARGS__); break;         \
        case 8: KERNEL<8><<<grid, 256, 0, st>>>(__VA_ARGS__); break;         \
        case 16: KERNEL<16><<<grid, 256, 0, st>>>(__VA_ARGS__); break;       \
        default: KERNEL<32><<<grid, 256, 0, st>>>(__VA_ARGS__); break;       \
    }


// x_new = x * exp(-t g), blocks normalised (lsq.cuh); dst != null: step and stop flag from the solver state
int md_update(bsls_ws *q, const bsls_plan *plan, double *x_new, const double *x, const double *g, double step, int per_block_log,
              cudaStream_t st, const DevState *dst) {
    const int lanes = lanes_for(plan);
    const int grid = grid_groups(plan->nb, lanes);
    if (plan->max_size <= 2 * lanes) {  // every block fits the registers of its lanes: one pass, one resident wave
#define MD_REG(G)                                                                                                     \
    {                                                                                                                 \
        static thread_local PerDevice<int> full_pd;                                                                   \
        int &full = full_pd.get(0);                                                                                   \
        if (!full) full = resident_grid(md_update_reg_kernel<G, 2>, 256);                                             \
        md_update_reg_kernel<G, 2><<<grid < full ? grid : full, 256, 0, st>>>(x_new, x, g, step, per_block_log, layout_of(plan), q->red, dst); \
    }
        switch (lanes) {
            case 4: MD_REG(4) break;
            case 8: MD_REG(8) break;
            case 16: MD_REG(16) break;
            default: MD_REG(32) break;
        }
#undef MD_REG
    } else {
        DISPATCH_LANES(lanes, md_update_kernel, grid, st, x_new, x, g, step, per_block_log, layout_of(plan), q->red, dst);
    }
    BSLS_LAUNCH_CHECK();
    q->launches++;
    return BSLS_OK;
}

// shared-memory layout of the cluster solver: sized for the largest share (group offsets known on the host)
bool cluster_layout(const bsls_lsq *q, int max_k, ClusterLayout *L, size_t *smem) {
    const int n = (int)q->n, m = (int)q->m;
    const std::vector<int32_t> &ga = q->sell_hgoff[0], &gt = q->sell_hgoff[1];
    L->groups_a = (m + 31) / 32;
    L->groups_t = (n + 31) / 32;
    L->max_k = max_k;
    if ((int)ga.size() != L->groups_a + 1 || (int)gt.size() != L->groups_t + 1) return false;
    int na = 0, nt = 0, wa = 0, wt = 0;
    for (int c = 0; c < kClusterCtas; ++c) {
        const int a0 = cluster_share(L->groups_a, c), a1 = cluster_share(L->groups_a, c + 1);
        const int t0 = cluster_share(L->groups_t, c), t1 = cluster_share_end(L->groups_t, c, max_k);
        na = std::max(na, ga[a1] - ga[a0]);
        nt = std::max(nt, gt[t1] - gt[t0]);
        wa = std::max(wa, a1 - a0);
        wt = std::max(wt, t1 - t0);
    }
    int d = 0;  // doubles
    L->x_off = d, d += 2 * (n + 1);
    L->r_off = d, d += 2 * (m + 1);
    L->g_cap = 32 * wt;
    L->g_off = d, d += 2 * L->g_cap;
    L->b_off = d, d += 32 * wa;
    L->sc_off = d, d += 64;
    L->part_off = d, d += 2 * kClusterCtas * 8;
    int w = 2 * d;  // 32-bit words
    L->ia_off = w, w += na;
    L->it_off = w, w += nt;
    L->ga_off = w, w += wa + 1;
    L->gt_off = w, w += wt + 1;
    L->st_off = w, w += 32 * wt + 2;  // at most one block per column of the share
    *smem = sizeof(int32_t) * (size_t)w + 16;
    return *smem <= 220 * 1024;
}

template <bool HV, int E, int G>
int launch_cluster_one(const TinyArgs &a, const DevOpts &d, const ClusterLayout &lay, size_t smem, cudaStream_t st) {
    static thread_local PerDevice<bool> attr_pd;
    bool &attr = attr_pd.get(false);
    if (!attr) {
        BSLS_CUDA_TRY(cudaFuncSetAttribute(solver_cluster_kernel<HV, E, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    solver_cluster_kernel<HV, E, G><<<kClusterCtas, kTinyThreads, smem, st>>>(a, d, lay);
    return BSLS_OK;
}
template <bool HV>
int launch_cluster_hv(const TinyArgs &a, const DevOpts &d, const ClusterLayout &lay, size_t smem, int max_k, cudaStream_t st) {
    if (max_k <= 8) return launch_cluster_one<HV, 2, 4>(a, d, lay, smem, st);
    if (max_k <= 16) return launch_cluster_one<HV, 2, 8>(a, d, lay, smem, st);
    if (max_k <= 32) return launch_cluster_one<HV, 4, 8>(a, d, lay, smem, st);
    return launch_cluster_one<HV, 8, 8>(a, d, lay, smem, st);
}
int launch_cluster(const TinyArgs &a, const DevOpts &d, const ClusterLayout &lay, size_t smem, bool has_values, int max_k, cudaStream_t st) {
    return has_values ? launch_cluster_hv<true>(a, d, lay, smem, max_k, st) : launch_cluster_hv<false>(a, d, lay, smem, max_k, st);
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// communicator
// ---------------------------------------------------------------------------------------------
int bsls_comm_unique_id(const char *nccl_path, char id[128]) {
    if (!id) return BSLS_ERR_ARG;
    if (int rc = nccl_load(nccl_path)) return rc;
    BSLS_NCCL_TRY(g_nccl.GetUniqueId(id));
    return BSLS_OK;
}

int bsls_comm_create(const char *nccl_path, int nranks, int rank, const char id[128], bsls_comm **out) {
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return BSLS_ERR_ARG;
    *out = nullptr;
    if (int rc = device_ok()) return rc;
    if (int rc = nccl_load(nccl_path)) return rc;
    Id128 uid;
    memcpy(uid.b, id, 128);
    bsls_comm *c = new bsls_comm();
    c->nranks = nranks;
    c->rank = rank;
    int r = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
    if (r != 0) {
        set_error("ncclCommInitRank -> %d (%s)", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        delete c;
        return BSLS_ERR_CUDA;
    }
    *out = c;
    return BSLS_OK;
}

// Peer-memory exchange: allocate this rank's region for link vectors of m entries and return its IPC handle (64 bytes) ...
int bsls_comm_p2p_alloc(bsls_comm *c, int64_t m, char handle[64]) {
    if (!c || !handle || m <= 0 || c->nranks < 2 || c->nranks > kP2pMaxRanks) {
        set_error("comm_p2p_alloc: needs 2..%d ranks of one node", kP2pMaxRanks);
        return BSLS_ERR_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (c->region) return BSLS_ERR_ARG;
    c->p2p_m = m;
    const size_t bytes = p2p_region_bytes(c->nranks, m);
    BSLS_CUDA_TRY(cudaMalloc(&c->region, bytes));
    BSLS_CUDA_TRY(cudaMemset(c->region, 0, bytes));
    BSLS_CUDA_TRY(cudaMalloc(&c->ticket, sizeof(unsigned)));
    BSLS_CUDA_TRY(cudaMemset(c->ticket, 0, sizeof(unsigned)));
    const int64_t rows = (m + c->nranks - 1) / c->nranks;
    int64_t want = (rows + kP2pThreads - 1) / kP2pThreads;
    c->reduce_grid = (int)(want < 2 * num_sms() ? (want < 1 ? 1 : want) : 2 * num_sms());
    BSLS_CUDA_TRY(cudaMalloc(&c->cta_partials, sizeof(double) * 3 * (size_t)c->reduce_grid));
    if (const char *e = getenv("BSLS_P2P_PROF"))
        if (atoi(e)) {
            BSLS_CUDA_TRY(cudaMalloc(&c->d_prof, 8 * sizeof(unsigned long long)));
            BSLS_CUDA_TRY(cudaMemset(c->d_prof, 0, 8 * sizeof(unsigned long long)));
        }
    cudaIpcMemHandle_t h;
    BSLS_CUDA_TRY(cudaIpcGetMemHandle(&h, c->region));
    memcpy(handle, &h, 64);
    BSLS_CUDA_TRY(cudaDeviceSynchronize());
    return BSLS_OK;
}

// ... and map the regions of the other ranks from their handles (nranks x 64 bytes, in rank order).
int bsls_comm_p2p_open(bsls_comm *c, const char *handles) {
    if (!c || !handles || !c->region) return BSLS_ERR_ARG;
    for (int q = 0; q < c->nranks; ++q) {
        if (q == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * q, 64);
        BSLS_CUDA_TRY(cudaIpcOpenMemHandle(&c->peer_base[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    p2p_fill_view(c);
    c->p2p_ready = true;
    return BSLS_OK;
}

int bsls_comm_p2p_ready(const bsls_comm *c) { return c && c->p2p_ready ? 1 : 0; }
int bsls_comm_p2p_disable(bsls_comm *c) {  // a rank failed to map its peers: every rank goes back to NCCL
    if (c) c->p2p_ready = false;
    return BSLS_OK;
}

int bsls_comm_destroy(bsls_comm *c) {
    if (!c) return BSLS_OK;
    for (int q = 0; q < kP2pMaxRanks; ++q)
        if (c->peer_base[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
    if (c->d_prof) {
        unsigned long long h[8] = {0};
        cudaMemcpy(h, c->d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        const double k = h[3] ? 1e-3 / (double)h[3] : 0.0;
        fprintf(stderr, "[p2p prof] rank %d: %llu exchanges; per exchange: wait for every rank's partial vector %.1f us, reduce + push %.1f us, "
                        "wait for every rank's rows %.1f us\n", c->rank, h[3], k * h[0], k * h[1], k * h[2]);
        cudaFree(c->d_prof);
    }
    if (c->region) cudaFree(c->region);
    if (c->ticket) cudaFree(c->ticket);
    if (c->cta_partials) cudaFree(c->cta_partials);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return BSLS_OK;
}

int bsls_comm_allreduce_sum_f64(bsls_comm *c, double *buf, int64_t count, bsls_stream_t s) {
    if (!c || !buf || count < 0) return BSLS_ERR_ARG;
    if (c->nranks <= 1) return BSLS_OK;
    BSLS_NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, c->comm, (cudaStream_t)s));
    return BSLS_OK;
}

// ---------------------------------------------------------------------------------------------
// reduction workspace
// ---------------------------------------------------------------------------------------------
int bsls_ws_create(bsls_ws **out) {
    if (!out) return BSLS_ERR_ARG;
    *out = nullptr;
    if (int rc = device_ok()) return rc;
    bsls_ws *w = new bsls_ws();
    auto fail = [&](int rc) {
        bsls_ws_destroy(w);
        return rc;
    };
#define TRY_OR_FAIL(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return fail(BSLS_ERR_CUDA);                                                      \
        }                                                                                    \
    } while (0)
    TRY_OR_FAIL(cudaMalloc(&w->red.partials, sizeof(double) * (size_t)kRedMaxGrid * kRedSlots));
    TRY_OR_FAIL(cudaMalloc(&w->red.ticket, sizeof(unsigned)));
    TRY_OR_FAIL(cudaMalloc(&w->d_scal, sizeof(double) * kScalCount));
    TRY_OR_FAIL(cudaMemset(w->red.ticket, 0, sizeof(unsigned)));
    TRY_OR_FAIL(cudaMemset(w->d_scal, 0, sizeof(double) * kScalCount));
    TRY_OR_FAIL(cudaHostAlloc(&w->h_scal, sizeof(double) * kScalCount, cudaHostAllocDefault));
    TRY_OR_FAIL(cudaMalloc(&w->d_state, sizeof(DevState)));
    TRY_OR_FAIL(cudaMemset(w->d_state, 0, sizeof(DevState)));
    TRY_OR_FAIL(cudaHostAlloc(&w->h_state, 2 * sizeof(DevState), cudaHostAllocDefault));
    for (int k = 0; k < 2; ++k) TRY_OR_FAIL(cudaEventCreateWithFlags(&w->ev_state[k], cudaEventDisableTiming));
#undef TRY_OR_FAIL
    w->red.out = w->d_scal;
    *out = w;
    return BSLS_OK;
}

int bsls_ws_destroy(bsls_ws *w) {
    if (!w) return BSLS_OK;
    if (w->red.partials) cudaFree(w->red.partials);
    if (w->red.ticket) cudaFree(w->red.ticket);
    if (w->d_scal) cudaFree(w->d_scal);
    if (w->h_scal) cudaFreeHost(w->h_scal);
    if (w->d_state) cudaFree(w->d_state);
    if (w->h_state) cudaFreeHost(w->h_state);
    for (int k = 0; k < 2; ++k)
        if (w->ev_state[k]) cudaEventDestroy(w->ev_state[k]);
    if (w->d_gather) cudaFree(w->d_gather);
    if (w->d_prog) cudaFree(w->d_prog);
    delete w;
    return BSLS_OK;
}

int bsls_ws_set_comm(bsls_ws *w, bsls_comm *c) {
    if (!w) return BSLS_ERR_ARG;
    w->comm = c;
    return BSLS_OK;
}

double *bsls_ws_scalar_ptr(const bsls_ws *w) { return w ? w->d_scal : nullptr; }

int bsls_ws_scalars(bsls_ws *w, double out[16], bsls_stream_t s) {
    if (!w || !out) return BSLS_ERR_ARG;
    if (int rc = fetch_scalars(w, (cudaStream_t)s)) return rc;
    memcpy(out, w->h_scal, sizeof(double) * kScalCount);
    return BSLS_OK;
}

// ---------------------------------------------------------------------------------------------
// problem handle
// ---------------------------------------------------------------------------------------------
int bsls_lsq_create(int64_t m, int64_t n, int64_t nnz, const int64_t *a_ptr, const int32_t *a_idx, const double *a_val,
                    const int64_t *at_ptr, const int32_t *at_idx, const double *at_val, const double *b, bsls_lsq **out) {
    if (!out) return BSLS_ERR_ARG;
    *out = nullptr;
    if (int rc = device_ok()) return rc;
    if (m <= 0 || n <= 0 || nnz < 0 || !a_ptr || !at_ptr || (nnz > 0 && (!a_idx || !at_idx))) {
        set_error("lsq_create: need m, n > 0 and both CSR structures (A and A^T)");
        return BSLS_ERR_ARG;
    }
    if (m >= (1LL << 31) || n >= (1LL << 31)) {
        set_error("lsq_create: int32 indices as in the reference (m, n < 2^31)");
        return BSLS_ERR_ARG;
    }
    bsls_lsq *q = new bsls_lsq();
    q->m = m;
    q->n = n;
    q->nnz = nnz;
    q->a_ptr = a_ptr;
    q->a_idx = a_idx;
    q->a_val = a_val;
    q->t_ptr = at_ptr;
    q->t_idx = at_idx;
    q->t_val = at_val;
    if (a_val && at_val && all_ones(a_val, nnz) && all_ones(at_val, nnz)) q->a_val = q->t_val = nullptr;  // 0/1 incidence
    q->b = b;
    q->a_mode = pick_mode(nnz, m);
    q->t_mode = pick_mode(nnz, n);
    // rows of one common length (every route traverses L links): the pointer-free ELL kernel
    q->t_ell = uniform_row_length(at_ptr, n, nnz);
    q->a_ell = uniform_row_length(a_ptr, m, nnz);
    if (ell_supported(q->t_ell)) q->t_mode = 2;
    if (ell_supported(q->a_ell) && q->a_ell <= 16) q->a_mode = 2;
    if (const char *e = getenv("BSLS_SPMV_A")) q->a_mode = atoi(e);
    if (const char *e = getenv("BSLS_SPMV_AT")) q->t_mode = atoi(e);
    auto fail = [&](int rc) {
        bsls_lsq_destroy(q);
        return rc;
    };
#define TRY_OR_FAIL(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return fail(BSLS_ERR_CUDA);                                                      \
        }                                                                                    \
    } while (0)
    TRY_OR_FAIL(cudaMalloc(&q->r, sizeof(double) * (size_t)m));
    TRY_OR_FAIL(cudaEventCreate(&q->ev0));
    TRY_OR_FAIL(cudaEventCreate(&q->ev1));
#undef TRY_OR_FAIL
    if (int rc = bsls_ws_create(&q->ws)) return fail(rc);
    *out = q;
    return BSLS_OK;
}

int bsls_lsq_destroy(bsls_lsq *q) {
    if (!q) return BSLS_OK;
    if (q->r) cudaFree(q->r);
    if (q->r2) cudaFree(q->r2);
    if (q->partial) cudaFree(q->partial);
    for (int k = 0; k < 2; ++k) {
        if (q->sell_idx[k]) cudaFree(q->sell_idx[k]);
        if (q->sell_goff[k]) cudaFree(q->sell_goff[k]);
        if (q->sell_val[k]) cudaFree(q->sell_val[k]);
    }
    if (q->wz) cudaFree(q->wz);
    if (q->wg) cudaFree(q->wg);
    if (q->wxn) cudaFree(q->wxn);
    if (q->wgn) cudaFree(q->wgn);
    bsls_ws_destroy(q->ws);
    if (q->ev0) cudaEventDestroy(q->ev0);
    if (q->ev1) cudaEventDestroy(q->ev1);
    delete q;
    return BSLS_OK;
}

int bsls_lsq_set_comm(bsls_lsq *q, bsls_comm *c) {
    if (!q) return BSLS_ERR_ARG;
    q->ws->comm = c;
    return BSLS_OK;
}

bsls_ws *bsls_lsq_ws(bsls_lsq *q) { return q ? q->ws : nullptr; }

int bsls_lsq_set_b(bsls_lsq *q, const double *b) {
    if (!q) return BSLS_ERR_ARG;
    q->b = b;
    return BSLS_OK;
}

int bsls_lsq_set_modes(bsls_lsq *q, int a_mode, int at_mode) {
    if (!q) return BSLS_ERR_ARG;
    auto ok = [](int v) { return v == 0 || v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32; };
    if (!ok(a_mode) || !ok(at_mode)) {
        set_error("lsq_set_modes: mode must be 0, 1, 2, 4, 8, 16 or 32");
        return BSLS_ERR_ARG;
    }
    q->a_mode = a_mode ? a_mode : (ell_supported(q->a_ell) ? 2 : pick_mode(q->nnz, q->m));
    q->t_mode = at_mode ? at_mode : (ell_supported(q->t_ell) ? 2 : pick_mode(q->nnz, q->n));
    return BSLS_OK;
}

int bsls_lsq_set_panels(bsls_lsq *q, int panels, const int64_t *ptr, const int32_t *idx, const double *val) {
    if (!q || panels < 0) return BSLS_ERR_ARG;
    if (q->partial) cudaFree(q->partial);
    q->partial = nullptr;
    q->panels = 0;
    if (panels <= 1) return BSLS_OK;
    if (!ptr || !idx) {
        set_error("lsq_set_panels: null structure");
        return BSLS_ERR_ARG;
    }
    BSLS_CUDA_TRY(cudaMalloc(&q->partial, sizeof(double) * (size_t)q->m * (size_t)panels));
    q->panels = panels;
    q->p_ptr = ptr;
    q->p_idx = idx;
    q->p_val = (q->a_val == nullptr) ? nullptr : val;  // unit values were detected at creation: the copy needs none either
    q->p_mode = pick_mode(q->nnz, q->m * panels);
    if (const char *e = getenv("BSLS_SPMV_P")) q->p_mode = atoi(e);
    return BSLS_OK;
}

const double *bsls_lsq_residual_ptr(const bsls_lsq *q) { return q ? q->r : nullptr; }
double *bsls_lsq_scalar_ptr(const bsls_lsq *q) { return q ? q->ws->d_scal : nullptr; }

int bsls_dev_lsq_residual_f64(bsls_lsq *q, const double *x, bsls_stream_t s) {
    if (!q || !x || !q->b) {
        set_error("lsq_residual: null handle, vector or b");
        return BSLS_ERR_ARG;
    }
    return residual(q, x, q->r, q->b, (cudaStream_t)s);
}

int bsls_dev_lsq_gradient_f64(bsls_lsq *q, double *g, bsls_stream_t s) {
    if (!q || !g) return BSLS_ERR_ARG;
    EpiPlain epi{g};
    return launch_at(q, q->r, epi, (cudaStream_t)s);
}

int bsls_dev_lsq_gradient_bb_f64(bsls_lsq *q, double *g_new, const double *g, const double *x, const double *x_new, bsls_stream_t s) {
    if (!q || !g_new || !g || !x || !x_new) return BSLS_ERR_ARG;
    EpiGradBB epi{g_new, g, x, x_new};
    return launch_at(q, q->r, epi, (cudaStream_t)s);
}

int bsls_dev_lsq_matvec_f64(bsls_lsq *q, const double *v, double *out, bsls_stream_t s) {
    if (!q || !v || !out) return BSLS_ERR_ARG;
    return residual(q, v, out, nullptr, (cudaStream_t)s);
}

int bsls_dev_lsq_rmatvec_f64(bsls_lsq *q, const double *w, double *out, bsls_stream_t s) {
    if (!q || !w || !out) return BSLS_ERR_ARG;
    EpiPlain epi{out};
    return launch_at(q, w, epi, (cudaStream_t)s);
}

int bsls_lsq_scalars(bsls_lsq *q, double out[16], bsls_stream_t s) { return q ? bsls_ws_scalars(q->ws, out, s) : BSLS_ERR_ARG; }

int bsls_lsq_obj_f64(bsls_lsq *q, const double *x, double *g, double *f_host, bsls_stream_t s) {
    if (!q || !x || !g || !f_host) return BSLS_ERR_ARG;
    if (int rc = bsls_dev_lsq_residual_f64(q, x, s)) return rc;
    if (int rc = bsls_dev_lsq_gradient_f64(q, g, s)) return rc;
    if (int rc = fetch_scalars(q->ws, (cudaStream_t)s)) return rc;
    *f_host = q->ws->h_scal[kScalF];
    return BSLS_OK;
}

// ---------------------------------------------------------------------------------------------
// vector kernels
// ---------------------------------------------------------------------------------------------
int bsls_dev_axpby_f64(double *out, double a, const double *x, double b, const double *y, int64_t n, bsls_stream_t s) {
    if (!out || !x || !y || n < 0) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    if (n == 0) return BSLS_OK;
    axpby_kernel<<<grid_elems(n), 256, 0, (cudaStream_t)s>>>(out, a, x, b, y, n);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

int bsls_ws_dots_f64(bsls_ws *q, int count, const double *const *x, const double *const *y, int64_t n, int want_max,
                      double out[5], bsls_stream_t s) {
    if (!q || count < 1 || count > 4 || !x || !y || !out || n <= 0) return BSLS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)s;
    DotArgs a{};
    a.count = count;
    a.want_max = want_max;
    for (int k = 0; k < 4; ++k) {
        a.x[k] = x[k < count ? k : 0];
        a.y[k] = y[k < count ? k : 0];
    }
    RedCtx red = q->red;
    red.out = q->d_scal + kScalDot0;  // slots 6..9, maximum in slot 10
    dots_kernel<<<grid_elems(n), 256, 0, st>>>(a, n, red);
    BSLS_LAUNCH_CHECK();
    q->launches++;
    if (int rc = allreduce(q, q->d_scal + kScalDot0, 4, kNcclSum, st)) return rc;
    if (want_max)
        if (int rc = allreduce(q, q->d_scal + kScalMax0, 1, kNcclMax, st)) return rc;
    if (int rc = fetch_scalars(q, st)) return rc;
    for (int k = 0; k < 5; ++k) out[k] = q->h_scal[kScalDot0 + k];
    return BSLS_OK;
}

int bsls_ws_flow_metrics_f64(bsls_ws *q, const double *scaling, const double *x_true, const double *x_hat, int64_t n, double thresh,
                             double out[5], bsls_stream_t s) {
    if (!q || !x_true || !x_hat || !out || n <= 0) return BSLS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)s;
    RedCtx red = q->red;
    red.out = q->d_scal + kScalDot0;  // slots 6..10
    flow_metrics_kernel<<<grid_elems(n), 256, 0, st>>>(scaling, x_true, x_hat, thresh, n, red);
    BSLS_LAUNCH_CHECK();
    q->launches++;
    if (int rc = fetch_scalars(q, st)) return rc;
    for (int k = 0; k < 5; ++k) out[k] = q->h_scal[kScalDot0 + k];
    return BSLS_OK;
}

int bsls_dev_axpy_dot_f64(bsls_ws *q, double *d, double scale, const double *c0, const double *c1, const double *v,
                          const double *w, double *out, int64_t n, bsls_stream_t s) {
    if (!q || !d || n <= 0 || (w && !out)) return BSLS_ERR_ARG;
    cudaStream_t st = (cudaStream_t)s;
    DevCoef c{c0, c1, scale};
    RedCtx red = q->red;
    red.out = w ? out : q->d_scal + (kScalCount - 1);  // slot 15 is scratch when no dot is wanted
    axpy_dot_kernel<<<grid_elems(n), 256, 0, st>>>(d, c, v, w, n, red);
    BSLS_LAUNCH_CHECK();
    q->launches++;
    if (w)
        if (int rc = allreduce(q, out, 1, kNcclSum, st)) return rc;
    return BSLS_OK;
}

int bsls_dev_md_update_f64(bsls_ws *q, const bsls_plan *plan, double *x_new, const double *x, const double *g, double step,
                           int per_block_log, bsls_stream_t s) {
    if (!q || !plan || !x_new || !x || !g) return BSLS_ERR_ARG;
    if (int rc = md_update(q, plan, x_new, x, g, step, per_block_log, (cudaStream_t)s, nullptr)) return rc;
    return allreduce(q, q->d_scal + kScalMax0, 1, kNcclMax, (cudaStream_t)s);
}

int bsls_dev_block_scale_f64(const bsls_plan *plan, double *y, const double *f, int divide, bsls_stream_t s) {
    if (!plan || !y || !f) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    const int lanes = lanes_for(plan);
    DISPATCH_LANES(lanes, block_scale_kernel, grid_groups(plan->nb, lanes), (cudaStream_t)s, y, f, divide, layout_of(plan));
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

int bsls_dev_nz_f64(const bsls_plan *plan, double *x, const double *z, int add_x0, bsls_stream_t s) {
    if (!plan || !x || !z) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    const int lanes = lanes_for(plan);
    DISPATCH_LANES(lanes, nz_kernel, grid_groups(plan->nb, lanes), (cudaStream_t)s, x, z, add_x0, layout_of(plan));
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

int bsls_dev_ntv_f64(const bsls_plan *plan, double *zg, const double *v, bsls_stream_t s) {
    if (!plan || !zg || !v) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    const int lanes = lanes_for(plan);
    DISPATCH_LANES(lanes, ntv_kernel, grid_groups(plan->nb, lanes), (cudaStream_t)s, zg, v, layout_of(plan));
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

int bsls_dev_x2z_f64(const bsls_plan *plan, const double *x, double *z, bsls_stream_t s) {
    if (!plan || !x || !z) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    if (plan->first != 0) {  // c_extensions.pyx:203 asserts blocks[0] == 0
        set_error("x2z: blocks[0] must be 0");
        return BSLS_ERR_ARG;
    }
    if (plan->uniform >= 2 && plan->uniform <= kX2zTileMaxK) {
        const int K = plan->uniform;
        const size_t smem = (size_t)kX2zTileThreads * (K | 1) * sizeof(double);
        static thread_local PerDevice<size_t> attr_smem_pd;
        size_t &attr_smem = attr_smem_pd.get(0);
        if (smem > attr_smem) {
            BSLS_CUDA_TRY(cudaFuncSetAttribute(x2z_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem = smem;
        }
        const int ntiles = (plan->nb + kX2zTileThreads - 1) / kX2zTileThreads;
        const int grid = ntiles < 8 * num_sms() ? ntiles : 8 * num_sms();
        x2z_tile_kernel<<<grid, kX2zTileThreads, smem, (cudaStream_t)s>>>(x, z, plan->nb, K, make_fastdiv((uint32_t)K),
                                                                           make_fastdiv((uint32_t)(K - 1)));
        BSLS_LAUNCH_CHECK();
        return BSLS_OK;
    }
    x2z_kernel<<<grid_groups(plan->nb, 1), 256, 0, (cudaStream_t)s>>>(x, z, layout_of(plan));
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

int bsls_dev_z2x_f64(const bsls_plan *plan, double *x, const double *z, bsls_stream_t s) {
    if (!plan || !x || !z) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    if (plan->first != 0) {
        set_error("z2x: blocks[0] must be 0");
        return BSLS_ERR_ARG;
    }
    {
        // the differences are independent of each other: the element-parallel kernel of x = x0 + N z computes exactly
        // z_l - z_{l-1} and 1 + (0 - z_last) = 1 - z_last (the same IEEE operations as the serial loop of z2x_kernel)
        const int lanes = lanes_for(plan);
        DISPATCH_LANES(lanes, nz_kernel, grid_groups(plan->nb, lanes), (cudaStream_t)s, x, z, 1, layout_of(plan));
    }
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

// ---------------------------------------------------------------------------------------------
// BATCH.solve / solve_BB / solve_MD, x-space, sparse objective
// ---------------------------------------------------------------------------------------------
// One objective evaluation at x_new, fused with everything the loop needs from it:
//   r = A x_new - b, f_new = 0.5 <r,r>                               (kernel 1 [+ all-reduce])
//   g_new = A^T r, <dx,dg>, <dg,dg>, <g,dx>, max|dx|                  (kernel 2)
// and one 128-byte copy of the scalars to the host.
static int eval_new(bsls_lsq *q, const double *x, const double *g, const double *x_new, double *g_new, cudaStream_t st) {
    if (int rc = residual(q, x_new, q->r, q->b, st)) return rc;
    EpiGradBB epi{g_new, g, x, x_new};
    if (int rc = launch_at(q, q->r, epi, st)) return rc;
    if (int rc = allreduce(q->ws, q->ws->d_scal + kScalSxy, 4, kNcclSum, st)) return rc;
    if (int rc = allreduce(q->ws, q->ws->d_scal + kScalStep, 1, kNcclMax, st)) return rc;
    return fetch_scalars(q->ws, st);
}

// The round-1 loop: host-side step logic, one 128-byte read-back and one stream synchronisation per objective
// evaluation, a back-track re-evaluates the objective with two more products.  Kept behind BSLS_BATCH_LEGACY=1 as the
// A/B reference of the device-resident loop below.
static int batch_solve_legacy(bsls_lsq *q, const bsls_plan *plan, double *x, const bsls_batch_opts *o, bsls_batch_result *res,
                              double *progress_f, double *progress_t, int progress_cap, bsls_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    if (int rc = ensure_workspace(q)) return rc;
    const int64_t n = q->n;
    double *xc = x, *g = q->wg, *xn = q->wxn, *gn = q->wgn;  // roles rotate by pointer swap; no vector is copied
    const int launches0 = q->ws->launches;
    int extra_launches = 0;
    int evals = 0, backtracks = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };

    BSLS_CUDA_TRY(cudaEventRecord(q->ev0, st));
    // f = obj(x, g)
    double f = 0.0;
    if (int rc = bsls_lsq_obj_f64(q, xc, g, &f, s)) return rc;
    ++evals;
    int np = 0;
    if (progress_f && np < progress_cap) {
        progress_f[np] = f;
        if (progress_t) progress_t[np] = 0.0;
        ++np;
    }
    const double t_start = now();
    double f_old = INFINITY, sxy = 0.0, syy = 0.0;
    int i = 1, stop_code = 0;
    double stop_value = 0.0;
    for (;;) {
        // algorithm_utils.stopping (:158-172): later tests overwrite the reason of earlier ones
        bool flag = false;
        if (i == o->max_iter) {
            stop_code = 1;
            flag = true;
        }
        if (o->has_f_min && f - o->f_min < o->opt_tol) {
            stop_code = 2;
            stop_value = f - o->f_min;
            flag = true;
        }
        if (std::fabs(f_old - f) < o->prog_tol) {
            stop_code = 3;
            stop_value = std::fabs(f_old - f);
            flag = true;
        }
        if (flag) break;
        // ---- trial point ---------------------------------------------------------------------
        if (o->method == 2) {
            const double t = 1.0 / (o->min_eig * i + 1.0);  // decreasing_step_size(i, 1.0, min_eig)
            if (int rc = bsls_dev_md_update_f64(q->ws, plan, xn, xc, g, t, 0, s)) return rc;
        } else {
            double t;
            if (o->method == 1)
                t = (i == 1) ? 1.0 : sxy / syy;  // BATCH.py:87-91
            else
                t = 1.0 / (o->min_eig * i + 1.0);  // BATCH.py:38
            bool fused = false;
            if (o->proj_mode != 2)  // x_new = proj(x - t g) in ONE kernel where the layout allows it
                if (int rc = project_step_f64(plan, xc, g, t, xn, o->proj_mode, st, &fused)) return rc;
            if (!fused) {
                axpby_kernel<<<grid_elems(n), 256, 0, st>>>(xn, 1.0, xc, -t, g, n);
                BSLS_LAUNCH_CHECK();
                ++extra_launches;
                if (o->proj_mode == 2) {  // z-space: isotonic regression + clip to [0,1] (algorithm_utils.py:219-224)
                    if (int rc = pava_clip_f64(plan, xn, nullptr, 1, 1, st)) return rc;
                } else {
                    if (int rc = project_f64(plan, xn, o->proj_mode, st)) return rc;
                }
            }
            ++extra_launches;
        }
        if (int rc = eval_new(q, xc, g, xn, gn, st)) return rc;
        ++evals;
        double f_new = q->ws->h_scal[kScalF];
        // ---- line_search_np (algorithm_utils.py:113-137) ------------------------------------------
        const bool search = (o->method == 1) || (o->method == 0 && o->use_line_search);
        if (search) {
            double t = 1.0;
            double upper = f + 1e-4 * q->ws->h_scal[kScalGd];
            while (f_new > upper) {
                t *= .8;
                if (q->ws->h_scal[kScalStep] < 1e-12) {  // step too small: stay where we are
                    f_new = f;
                    BSLS_CUDA_TRY(cudaMemcpyAsync(gn, g, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
                    BSLS_CUDA_TRY(cudaMemcpyAsync(xn, xc, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
                    q->ws->h_scal[kScalSxy] = 0.0;
                    q->ws->h_scal[kScalSyy] = 0.0;
                    break;
                }
                axpby_kernel<<<grid_elems(n), 256, 0, st>>>(xn, 1.0 - t, xc, t, xn, n);
                BSLS_LAUNCH_CHECK();
                ++extra_launches;
                if (int rc = eval_new(q, xc, g, xn, gn, st)) return rc;
                ++evals;
                ++backtracks;
                f_new = q->ws->h_scal[kScalF];
                upper = f + 1e-4 * q->ws->h_scal[kScalGd];
            }
        }
        // ---- take the step: delta_x / delta_g only ever enter through their dot products -------------
        sxy = q->ws->h_scal[kScalSxy];
        syy = q->ws->h_scal[kScalSyy];
        f_old = f;
        f = f_new;
        double *tx = xc, *tg = g;
        xc = xn;
        g = gn;
        xn = tx;
        gn = tg;
        ++i;
        if (progress_f && np < progress_cap) {
            progress_f[np] = f;
            if (progress_t) progress_t[np] = now() - t_start;
            ++np;
        }
    }
    if (xc != x) BSLS_CUDA_TRY(cudaMemcpyAsync(x, xc, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    BSLS_CUDA_TRY(cudaEventRecord(q->ev1, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    BSLS_CUDA_TRY(cudaEventElapsedTime(&ms, q->ev0, q->ev1));
    res->f = f;
    res->iterations = i;
    res->stop_code = stop_code;
    res->stop_value = stop_value;
    res->obj_evals = evals;
    res->backtracks = backtracks;
    res->kernel_launches = (q->ws->launches - launches0) + extra_launches;
    res->device_ms = ms;
    return BSLS_OK;
}

// ---------------------------------------------------------------------------------------------
// device-resident loop
// ---------------------------------------------------------------------------------------------
// Buffers come in two sets (x, g, r)[0..1]; iteration k reads set k & 1 (the iterate) and writes set (k + 1) & 1 (the
// trial point).  The trial point always becomes the next iterate -- directly, or pulled back along the segment by
// commit_kernel after a back-tracked line search -- so the roles alternate by iteration number and no kernel argument
// depends on a decision taken on the device.  The host therefore enqueues iteration k + 1 BEFORE it learns how
// iteration k ended (one iteration of slack: the GPU never waits for the host), reads the 80-byte solver state of
// iteration k from pinned memory when its event fires, and stops enqueuing once `done` is set.  An iteration enqueued
// past the stop exits in every kernel's first instructions and touches only buffers that do not hold the result.
namespace {
struct LoopBuffers {
    double *x[2], *g[2], *r[2];
};

int ensure_loop_scratch(bsls_ws *w, int prog_cap) {
    const int nr = (w->comm && w->comm->nranks > 1) ? w->comm->nranks : 0;
    if (nr > w->gather_cap) {
        if (w->d_gather) cudaFree(w->d_gather);
        w->d_gather = nullptr;
        BSLS_CUDA_TRY(cudaMalloc(&w->d_gather, sizeof(double) * kStepScalars * (size_t)nr));
        w->gather_cap = nr;
    }
    if (prog_cap > w->prog_cap) {
        if (w->d_prog) cudaFree(w->d_prog);
        w->d_prog = nullptr;
        BSLS_CUDA_TRY(cudaMalloc(&w->d_prog, sizeof(double) * 2 * (size_t)prog_cap));
        w->prog_cap = prog_cap;
    }
    return BSLS_OK;
}

DevOpts make_dev_opts(const bsls_batch_opts *o, const bsls_ws *w, int cap) {
    DevOpts d{};
    d.method = o->method;
    d.corrections = o->corrections > 0 ? (o->corrections < 64 ? o->corrections : 64) : 50;
    d.search = (o->method == 1) || (o->method == 5) || (o->method == 0 && o->use_line_search);
    d.has_f_min = o->has_f_min;
    d.max_iter = o->max_iter;
    d.f_min = o->f_min;
    d.opt_tol = o->opt_tol;
    d.prog_tol = o->prog_tol;
    d.min_eig = o->min_eig;
    d.nranks = (w->comm && w->comm->nranks > 1) ? w->comm->nranks : 1;
    d.progress_cap = cap;
    return d;
}

void finish_result(bsls_batch_result *res, const DevState &fin, int launches, float ms, double *progress_t, int cap) {
    if (cap > 0 && progress_t) {  // device time stamps (ns) -> seconds since the first objective value
        const int np = fin.i < cap ? fin.i : cap;
        const double t0 = progress_t[0];
        for (int j = 0; j < np; ++j) progress_t[j] = (progress_t[j] - t0) * 1e-9;
    }
    res->f = fin.f;
    res->iterations = fin.i;
    res->stop_code = fin.done;
    res->stop_value = fin.stop_value;
    res->obj_evals = fin.evals;
    res->backtracks = fin.backtracks;
    res->kernel_launches = launches;
    res->device_ms = ms;
}

// sliced-ELL copy of one CSR side (solver_tiny.cuh)
int build_sell(const int64_t *ptr, const int32_t *idx, const double *val, int rows, int pad, int32_t **goff_out, int32_t **sidx_out,
               double **sval_out, std::vector<int32_t> *hgoff, cudaStream_t st) {
    const int groups = (rows + 31) / 32;
    BSLS_CUDA_TRY(cudaMalloc(goff_out, sizeof(int32_t) * ((size_t)groups + 1)));
    sell_widths_kernel<<<(groups + 127) / 128, 128, 0, st>>>(ptr, rows, groups, *goff_out);
    sell_scan_kernel<<<1, 32, 0, st>>>(*goff_out, groups);
    BSLS_LAUNCH_CHECK();
    hgoff->resize((size_t)groups + 1);
    BSLS_CUDA_TRY(cudaMemcpyAsync(hgoff->data(), *goff_out, sizeof(int32_t) * ((size_t)groups + 1), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    const int32_t total = hgoff->back();
    BSLS_CUDA_TRY(cudaMalloc(sidx_out, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)));
    if (val) BSLS_CUDA_TRY(cudaMalloc(sval_out, sizeof(double) * (size_t)(total > 0 ? total : 1)));
    sell_fill_kernel<<<(32 * groups + 255) / 256, 256, 0, st>>>(ptr, idx, val, rows, groups, *goff_out, *sidx_out, val ? *sval_out : nullptr, pad);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

// Problems whose vectors fit one SM's shared memory: the whole loop in one launch of one CTA (solver_tiny.cuh).
// Returns 0 when the problem does not qualify (the caller goes on), -1 when it was solved here, > 0 on error.
int solve_tiny(bsls_lsq *q, const bsls_plan *plan, double *x, const bsls_batch_opts *o, bsls_batch_result *res, double *progress_f,
               double *progress_t, int cap, cudaStream_t st) {
    const char *e = getenv("BSLS_NO_TINY");  // read per call: the tests run the same problems through both loops
    const bool off = e && atoi(e) != 0;
    bsls_ws *w = q->ws;
    const size_t smem = tiny_smem_bytes((int)q->n, (int)q->m);
    const bool single_fits = smem <= 220 * 1024;
    const bool cluster_may_fit = 16 * (size_t)(q->n + q->m + 2) <= 200 * 1024;  // x and r twice; the exact need follows below
    if (off || (w->comm && w->comm->nranks > 1) || o->method > 1 || o->proj_mode > 1 || plan->max_size > kTinyMaxBlock || plan->first != 0 ||
        (!single_fits && !cluster_may_fit) || q->n > (1 << 20) || q->nnz > (1 << 26))
        return 0;
    static thread_local PerDevice<bool> attr_pd;
    bool &attr = attr_pd.get(false);
    if (!attr) {
        BSLS_CUDA_TRY(cudaFuncSetAttribute(solver_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    if (!q->sell_ready) {
        if (int rc = build_sell(q->a_ptr, q->a_idx, q->a_val, (int)q->m, (int)q->n, &q->sell_goff[0], &q->sell_idx[0], &q->sell_val[0], &q->sell_hgoff[0], st)) return rc;
        if (int rc = build_sell(q->t_ptr, q->t_idx, q->t_val, (int)q->n, (int)q->m, &q->sell_goff[1], &q->sell_idx[1], &q->sell_val[1], &q->sell_hgoff[1], st)) return rc;
        q->sell_ready = true;
    }
    TinyArgs a{};
    a.n = (int)q->n;
    a.m = (int)q->m;
    a.nb = plan->nb;
    a.starts = plan->d_starts;
    a.A = SellMatrix{q->sell_idx[0], q->sell_val[0], q->sell_goff[0]};
    a.AT = SellMatrix{q->sell_idx[1], q->sell_val[1], q->sell_goff[1]};
    a.b = q->b;
    a.x = x;
    a.st = w->d_state;
    a.progress_f = cap > 0 ? w->d_prog : nullptr;
    a.progress_t = cap > 0 ? w->d_prog + w->prog_cap : nullptr;
    a.proj_mode = o->proj_mode;
    // A cluster of 8 CTAs (solver_cluster.cuh) when the shares fit and are worth it; BSLS_TINY_CLUSTER=0 / 1 forces one
    // kernel or the other where both fit (both are tested on every small problem).
    const char *ce = getenv("BSLS_TINY_CLUSTER");
    ClusterLayout lay{};
    size_t csmem = 0;
    const bool cluster_fits = cluster_layout(q, plan->max_size, &lay, &csmem);
    if (!single_fits && !cluster_fits) return 0;
    const bool clustered =
        cluster_fits && (!single_fits || (ce ? atoi(ce) != 0 : (plan->nb >= 4 * kClusterCtas && q->m >= 32 * kClusterCtas)));
    const DevOpts d = make_dev_opts(o, w, cap);
    const char *pe = getenv("BSLS_TINY_PROF");  // development: phase cycle counts of the cluster kernel on stderr
    long long *d_prof = nullptr;
    if (pe && atoi(pe)) {
        BSLS_CUDA_TRY(cudaMalloc(&d_prof, 16 * sizeof(long long)));
        BSLS_CUDA_TRY(cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), st));
        a.prof = d_prof;
    }
    BSLS_CUDA_TRY(cudaEventRecord(q->ev0, st));
    if (clustered) {
        if (int rc = launch_cluster(a, d, lay, csmem, q->sell_val[0] != nullptr, plan->max_size, st)) return rc;
    } else {
        solver_tiny_kernel<<<1, kTinyThreads, smem, st>>>(a, d);
    }
    BSLS_LAUNCH_CHECK();
    BSLS_CUDA_TRY(cudaEventRecord(q->ev1, st));
    BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[0], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    const DevState fin = w->h_state[0];
    if (d_prof) {
        long long h[16];
        BSLS_CUDA_TRY(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d_prof);
        static const char *names[12] = {"project", "sync_x", "rows_Ax", "publish_r", "rows_Atr", "publish_g", "decide", "sync_decide",
                                        "pullback", "-", "-", "-"};
        fprintf(stderr, "[tiny prof] iterations %d, cycles per iteration:", fin.i);
        for (int k = 0; k < 9; ++k) fprintf(stderr, " %s=%.0f", names[k], (double)h[k] / (fin.i > 1 ? fin.i - 1 : 1));
        fprintf(stderr, "\n");
    }
    if (!fin.done) {
        set_error("batch_solve (single-CTA loop): ended without a stop code (i=%d)", fin.i);
        return BSLS_ERR_CUDA;
    }
    if (cap > 0) {
        const int np = fin.i < cap ? fin.i : cap;
        BSLS_CUDA_TRY(cudaMemcpy(progress_f, w->d_prog, sizeof(double) * (size_t)np, cudaMemcpyDeviceToHost));
        if (progress_t) BSLS_CUDA_TRY(cudaMemcpy(progress_t, w->d_prog + w->prog_cap, sizeof(double) * (size_t)np, cudaMemcpyDeviceToHost));
    }
    float ms = 0.f;
    BSLS_CUDA_TRY(cudaEventElapsedTime(&ms, q->ev0, q->ev1));
    finish_result(res, fin, 1, ms, progress_t, cap);
    return -1;
}

// one iteration of BATCH.solve / solve_BB / solve_MD on the device: trial point, objective + gradient there, decision
int enqueue_iteration(bsls_lsq *q, const bsls_plan *plan, const LoopBuffers &B, int cur, const DevOpts &d, int proj_mode, cudaStream_t st,
                      int *extra_launches) {
    bsls_ws *w = q->ws;
    const int nxt = cur ^ 1;
    const int64_t n = q->n;
    DevState *S = w->d_state;
    const int *done = &S->done;
    const bool dist = w->comm && w->comm->nranks > 1;
    // ---- trial point --------------------------------------------------------------------------
    if (d.method == 2) {
        if (int rc = md_update(w, plan, B.x[nxt], B.x[cur], B.g[cur], 0.0, 0, st, S)) return rc;
    } else if (d.method == 5) {  // x_new = proj(x + d), d from the scalar two-loop recursion; the trial set still holds (x_prev, g_prev)
        lbfgs_step_kernel<<<grid_elems(n), 256, 0, st>>>(B.x[nxt], B.x[cur], B.x[nxt], B.g[cur], B.g[nxt], S, n);
        BSLS_LAUNCH_CHECK();
        ++*extra_launches;
        if (proj_mode == 2) {
            if (int rc = pava_clip_f64(plan, B.x[nxt], nullptr, 1, 1, st)) return rc;
        } else {
            if (int rc = project_f64(plan, B.x[nxt], proj_mode, st)) return rc;
        }
        ++*extra_launches;
    } else {
        bool fused = false;
        const StepCtl ctl{&S->t, done};
        if (proj_mode != 2)  // x_new = proj(x - t g) in ONE kernel where the layout allows it
            if (int rc = project_step_f64(plan, B.x[cur], B.g[cur], 0.0, B.x[nxt], proj_mode, st, &fused, &ctl)) return rc;
        if (!fused) {
            step_axpy_kernel<<<grid_elems(n), 256, 0, st>>>(B.x[nxt], B.x[cur], B.g[cur], S, n);
            BSLS_LAUNCH_CHECK();
            ++*extra_launches;
            if (proj_mode == 2) {  // z-space: isotonic regression + clip to [0,1] (algorithm_utils.py:219-224)
                if (int rc = pava_clip_f64(plan, B.x[nxt], nullptr, 1, 1, st)) return rc;
            } else {
                if (int rc = project_f64(plan, B.x[nxt], proj_mode, st)) return rc;
            }
        }
        ++*extra_launches;
    }
    // ---- objective and gradient at the trial point ------------------------------------------------
    const bool p2p = dist && use_p2p(q);
    if (int rc = residual(q, B.x[nxt], B.r[nxt], q->b, st, B.r[cur], done, p2p ? nxt : -1)) return rc;
    const bool need_dots = d.method == 1 || d.method == 5 || d.search;
    if (need_dots) {
        EpiGradBB epi{B.g[nxt], B.g[cur], B.x[cur], B.x[nxt]};
        if (int rc = launch_at(q, B.r[nxt], epi, st, done)) return rc;
        if (p2p) {  // every rank needs every rank's share of the step scalars: pushed over NVLink, flag C (p2p.cuh) ...
            p2p_post_kernel<<<1, 32, 0, st>>>(w->comm->view, w->comm->epoch, w->d_scal, done);
            BSLS_LAUNCH_CHECK();
            ++*extra_launches;
        } else if (dist) {  // ... or one small NCCL all-gather (slots 1..6)
            BSLS_NCCL_TRY(g_nccl.AllGather(w->d_scal + kScalSxy, w->d_gather, kStepScalars, kNcclFloat64, w->comm->comm, st));
        }
    } else {
        EpiPlain epi{B.g[nxt]};
        if (int rc = launch_at(q, B.r[nxt], epi, st, done)) return rc;
    }
    // ---- decision, and the pull-back of a back-tracked trial point ----------------------------------
    if (need_dots && p2p) {
        const bsls_comm *c = w->comm;
        decide_kernel<<<1, 32, 0, st>>>(S, w->d_scal, c->view.gath[c->rank], d, w->d_prog, w->d_prog ? w->d_prog + w->prog_cap : nullptr, 0,
                                        c->view.flags[c->rank] + 2 * c->nranks, c->epoch);
    } else {
        decide_kernel<<<1, 32, 0, st>>>(S, w->d_scal, (need_dots && dist) ? w->d_gather : nullptr, d, w->d_prog,
                                        w->d_prog ? w->d_prog + w->prog_cap : nullptr, 0);
    }
    BSLS_LAUNCH_CHECK();
    ++*extra_launches;
    if (d.search) {
        const int64_t big = n > q->m ? n : q->m;
        commit_kernel<<<grid_elems(big), 256, 0, st>>>(S, B.x[nxt], B.x[cur], B.g[nxt], B.g[cur], n, B.r[nxt], B.r[cur], q->m);
        BSLS_LAUNCH_CHECK();
        ++*extra_launches;
    }
    return BSLS_OK;
}
}  // namespace

int bsls_batch_solve_f64(bsls_lsq *q, const bsls_plan *plan, double *x, const bsls_batch_opts *o, bsls_batch_result *res,
                         double *progress_f, double *progress_t, int progress_cap, bsls_stream_t s) {
    if (!q || !plan || !x || !o || !res || !q->b) {
        set_error("batch_solve: null argument");
        return BSLS_ERR_ARG;
    }
    if (plan->n != q->n) {
        set_error("batch_solve: plan covers %d variables, A has %lld columns", plan->n, (long long)q->n);
        return BSLS_ERR_ARG;
    }
    if (!(o->method == 0 || o->method == 1 || o->method == 2 || o->method == 5)) return BSLS_ERR_ARG;
    if (int rc = ensure_workspace(q)) return rc;
    static const bool legacy = [] {
        const char *e = getenv("BSLS_BATCH_LEGACY");
        return e && atoi(e) != 0;
    }();
    if (legacy && o->method <= 2) return batch_solve_legacy(q, plan, x, o, res, progress_f, progress_t, progress_cap, s);

    cudaStream_t st = (cudaStream_t)s;
    bsls_ws *w = q->ws;
    const int cap = (progress_f && progress_cap > 0) ? progress_cap : 0;
    if (int rc = ensure_loop_scratch(w, cap)) return rc;
    if (int rc = solve_tiny(q, plan, x, o, res, progress_f, progress_t, cap, st)) return rc < 0 ? BSLS_OK : rc;
    const DevOpts d = make_dev_opts(o, w, cap);
    LoopBuffers B{{x, q->wxn}, {q->wg, q->wgn}, {q->r, q->r2}};
    if (use_p2p(q)) {  // the residual lives in the exchange region: the owners of the rows write it there
        B.r[0] = w->comm->view.rfull[w->comm->rank][0];
        B.r[1] = w->comm->view.rfull[w->comm->rank][1];
    }
    const int launches0 = w->launches;
    int extra = 0;

    BSLS_CUDA_TRY(cudaEventRecord(q->ev0, st));
    BSLS_CUDA_TRY(cudaMemsetAsync(w->d_state, 0, sizeof(DevState), st));
    // f = obj(x, g) at the starting point
    if (int rc = residual(q, B.x[0], B.r[0], q->b, st, nullptr, nullptr, use_p2p(q) ? 0 : -1)) return rc;
    {
        EpiPlain epi{B.g[0]};
        if (int rc = launch_at(q, B.r[0], epi, st)) return rc;
    }
    decide_kernel<<<1, 32, 0, st>>>(w->d_state, w->d_scal, nullptr, d, w->d_prog, w->d_prog ? w->d_prog + w->prog_cap : nullptr, 1);
    BSLS_LAUNCH_CHECK();
    ++extra;
    BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[1], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaEventRecord(w->ev_state[1], st));

    // iteration k may run only if the state after iteration k - 1 is not `done`; the host learns that one iteration late
    // the reference stops at i == max_iter, i starting at 1; max_iter <= 0 never matches: no iteration limit there either
    const int max_body = o->max_iter > 1 ? o->max_iter - 1 : (o->max_iter == 1 ? 0 : 0x7fffffff);
    int k = 0;
    for (; k < max_body; ++k) {
        if (int rc = enqueue_iteration(q, plan, B, k & 1, d, o->proj_mode, st, &extra)) return rc;
        BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[k & 1], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
        BSLS_CUDA_TRY(cudaEventRecord(w->ev_state[k & 1], st));
        // the state BEFORE this iteration (after iteration k - 1, or after the initial evaluation)
        BSLS_CUDA_TRY(cudaEventSynchronize(w->ev_state[(k + 1) & 1]));
        if (w->h_state[(k + 1) & 1].done) break;
    }
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    DevState fin;
    BSLS_CUDA_TRY(cudaMemcpy(&fin, w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost));
    if (!fin.done) {
        set_error("batch_solve: the device loop ended without a stop code (i=%d)", fin.i);
        return BSLS_ERR_CUDA;
    }
    if (fin.parity != 0) BSLS_CUDA_TRY(cudaMemcpyAsync(x, B.x[1], sizeof(double) * (size_t)q->n, cudaMemcpyDeviceToDevice, st));
    if (cap > 0) {
        const int np = fin.i < cap ? fin.i : cap;
        BSLS_CUDA_TRY(cudaMemcpyAsync(progress_f, w->d_prog, sizeof(double) * (size_t)np, cudaMemcpyDeviceToHost, st));
        if (progress_t) BSLS_CUDA_TRY(cudaMemcpyAsync(progress_t, w->d_prog + w->prog_cap, sizeof(double) * (size_t)np, cudaMemcpyDeviceToHost, st));
    }
    BSLS_CUDA_TRY(cudaEventRecord(q->ev1, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    BSLS_CUDA_TRY(cudaEventElapsedTime(&ms, q->ev0, q->ev1));
    finish_result(res, fin, (w->launches - launches0) + extra, ms, progress_t, cap);
    return BSLS_OK;
}

// mirror_descent.least_squares (python/mirror_descent.py:7-53) as a device-resident loop: per iteration the SpMV pair,
// one fused exponentiate / normalise / max-change kernel whose step sqrt(2 ln K) / (sqrt(k) Lf) is formed on the device
// from the iteration counter, [one scalar max all-reduce when sharded,] and the stop test max |x - x_prev| < tolerance
// taken by decide_kernel.  x: in = starting point (1 / K_block), out = result.
int bsls_md_least_squares_f64(bsls_lsq *q, const bsls_plan *plan, double *x, int iters, double tolerance, double Lf, bsls_batch_result *res,
                              bsls_stream_t s) {
    if (!q || !plan || !x || !res || !q->b || plan->n != q->n || !(Lf > 0)) {
        set_error("md_least_squares: bad argument");
        return BSLS_ERR_ARG;
    }
    if (int rc = ensure_workspace(q)) return rc;
    cudaStream_t st = (cudaStream_t)s;
    bsls_ws *w = q->ws;
    DevOpts d{};
    d.method = 4;
    d.max_iter = iters;
    d.tolerance = tolerance;
    d.Lf = Lf;
    d.nranks = (w->comm && w->comm->nranks > 1) ? w->comm->nranks : 1;
    double *X[2] = {x, q->wxn};
    double *g = q->wg;
    const int launches0 = w->launches;
    int extra = 0;
    BSLS_CUDA_TRY(cudaEventRecord(q->ev0, st));
    BSLS_CUDA_TRY(cudaMemsetAsync(w->d_state, 0, sizeof(DevState), st));
    decide_kernel<<<1, 32, 0, st>>>(w->d_state, w->d_scal, nullptr, d, nullptr, nullptr, 1);
    BSLS_LAUNCH_CHECK();
    BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[1], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaEventRecord(w->ev_state[1], st));
    const int *done = &w->d_state->done;
    for (int k = 0; k < iters; ++k) {
        const int cur = k & 1, nxt = cur ^ 1;
        if (int rc = residual(q, X[cur], q->r, q->b, st, nullptr, done)) return rc;
        EpiPlain epi{g};
        if (int rc = launch_at(q, q->r, epi, st, done)) return rc;
        if (int rc = md_update(w, plan, X[nxt], X[cur], g, Lf, 1, st, w->d_state)) return rc;
        if (int rc = allreduce(w, w->d_scal + kScalMax0, 1, kNcclMax, st)) return rc;
        decide_kernel<<<1, 32, 0, st>>>(w->d_state, w->d_scal, nullptr, d, nullptr, nullptr, 0);
        BSLS_LAUNCH_CHECK();
        ++extra;
        BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[k & 1], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
        BSLS_CUDA_TRY(cudaEventRecord(w->ev_state[k & 1], st));
        BSLS_CUDA_TRY(cudaEventSynchronize(w->ev_state[(k + 1) & 1]));  // the state one iteration back: the GPU keeps a full iteration queued
        if (w->h_state[(k + 1) & 1].done) break;
    }
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    DevState fin;
    BSLS_CUDA_TRY(cudaMemcpy(&fin, w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost));
    if (fin.parity != 0) BSLS_CUDA_TRY(cudaMemcpyAsync(x, X[1], sizeof(double) * (size_t)q->n, cudaMemcpyDeviceToDevice, st));
    BSLS_CUDA_TRY(cudaEventRecord(q->ev1, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    BSLS_CUDA_TRY(cudaEventElapsedTime(&ms, q->ev0, q->ev1));
    res->f = 0.0;
    res->iterations = fin.i - 1;  // update steps taken
    res->stop_code = fin.done;
    res->stop_value = fin.change;
    res->obj_evals = fin.evals;
    res->backtracks = 0;
    res->kernel_launches = (w->launches - launches0) + extra + 1;
    res->device_ms = ms;
    return BSLS_OK;
}

// BB.solve (python/BB.py:7-45) for the z-space problem of main.solve_in_z (python/main.py:47-65) as a device-resident loop:
//   f(z) = 0.5 |A (N z) - b'|^2 with b' = -target held by the handle, grad = N^T A^T r, proj = isotonic regression + clip.
// Per iteration: A^T r, N^T, one pass for the four sums, z - t g, the projection, N z, A x -- and the decision
// (BB step, the "no change in gradient" exit, solvers.stopping) by decide_kernel.  The reference evaluates A N z three
// times per iteration (gradient, objective, next gradient); here the residual at the new point serves the stopping test
// and the next gradient: two products.  Runs iterations i_start + 1 .. i_end (segments let the caller record states every
// `record_every` iterations as the reference's log callback does); z, z_prev, g_prev are updated in place.
int bsls_zbb_run_f64(bsls_lsq *q, const bsls_plan *xplan, const bsls_plan *zplan, double *z, double *z_prev, double *g_prev, int i_start,
                     int i_end, int max_iter, double opt_tol, bsls_batch_result *res, bsls_stream_t s) {
    if (!q || !xplan || !zplan || !z || !z_prev || !g_prev || !res || !q->b || xplan->n != q->n || i_end < i_start) {
        set_error("zbb_run: bad argument");
        return BSLS_ERR_ARG;
    }
    const int64_t n = q->n, nz = zplan->n;
    if (nz != n - xplan->nb) {
        set_error("zbb_run: z has %lld entries, expected n - numblocks = %lld", (long long)nz, (long long)(n - xplan->nb));
        return BSLS_ERR_ARG;
    }
    if (int rc = ensure_workspace(q)) return rc;
    if (!q->wz) BSLS_CUDA_TRY(cudaMalloc(&q->wz, sizeof(double) * (size_t)q->n));
    cudaStream_t st = (cudaStream_t)s;
    bsls_ws *w = q->ws;
    DevOpts d{};
    d.method = 3;
    d.max_iter = max_iter;
    d.opt_tol = opt_tol;
    d.nranks = (w->comm && w->comm->nranks > 1) ? w->comm->nranks : 1;
    double *Z[3] = {z, z_prev, q->wgn}, *G[2] = {g_prev, q->wz};
    double *xbuf = q->wxn, *gx = q->wg;
    const int lanes = lanes_for(xplan);
    const int ngrid = grid_groups(xplan->nb, lanes);
    auto n_dot = [&](double *x_out, const double *zz) -> int {  // x = N z
        DISPATCH_LANES(lanes, nz_kernel, ngrid, st, x_out, zz, 0, layout_of(xplan));
        BSLS_LAUNCH_CHECK();
        return BSLS_OK;
    };
    const int launches0 = w->launches;
    int extra = 0;
    BSLS_CUDA_TRY(cudaEventRecord(q->ev0, st));
    DevState init{};
    init.i = i_start;
    BSLS_CUDA_TRY(cudaMemcpyAsync(w->d_state, &init, sizeof(DevState), cudaMemcpyHostToDevice, st));
    const int *done = &w->d_state->done;
    // r = A N z - b' at the starting point of the segment
    if (int rc = n_dot(xbuf, Z[0])) return rc;
    if (int rc = residual(q, xbuf, q->r, q->b, st)) return rc;
    BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[1], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaEventRecord(w->ev_state[1], st));
    const int todo = i_end - i_start;
    for (int k = 0; k < todo; ++k) {
        const int cur = (2 * k) % 3, prev = (2 * k + 1) % 3, nxt = (2 * k + 2) % 3;  // cur_k = -k, prev_k = 1 - k, nxt_k = 2 - k (mod 3)
        const int gc = (k + 1) & 1, gp = k & 1;
        // g = N^T A^T r
        EpiPlain epi{gx};
        if (int rc = launch_at(q, q->r, epi, st, done)) return rc;
        DISPATCH_LANES(lanes, ntv_kernel, ngrid, st, G[gc], gx, layout_of(xplan));
        BSLS_LAUNCH_CHECK();
        zbb_dots_kernel<<<grid_elems(nz), 256, 0, st>>>(Z[cur], Z[prev], G[gc], G[gp], nz, w->red, done);
        BSLS_LAUNCH_CHECK();
        if (int rc = allreduce(w, w->d_scal + kScalSxy, 4, kNcclSum, st)) return rc;
        // z_new = proj(z - t g)
        zbb_step_kernel<<<grid_elems(nz), 256, 0, st>>>(Z[nxt], Z[cur], G[gc], w->d_scal, done, nz);
        BSLS_LAUNCH_CHECK();
        if (int rc = pava_clip_f64(zplan, Z[nxt], nullptr, 1, 1, st)) return rc;
        // f(z_new) -- and the residual the next gradient starts from
        if (int rc = n_dot(xbuf, Z[nxt])) return rc;
        if (int rc = residual(q, xbuf, q->r, q->b, st, nullptr, done)) return rc;
        decide_kernel<<<1, 32, 0, st>>>(w->d_state, w->d_scal, nullptr, d, nullptr, nullptr, 0);
        BSLS_LAUNCH_CHECK();
        extra += 6;
        BSLS_CUDA_TRY(cudaMemcpyAsync(&w->h_state[k & 1], w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, st));
        BSLS_CUDA_TRY(cudaEventRecord(w->ev_state[k & 1], st));
        BSLS_CUDA_TRY(cudaEventSynchronize(w->ev_state[(k + 1) & 1]));  // the state one iteration back
        if (w->h_state[(k + 1) & 1].done) break;
    }
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    DevState fin;
    BSLS_CUDA_TRY(cudaMemcpy(&fin, w->d_state, sizeof(DevState), cudaMemcpyDeviceToHost));
    // hand the three vectors back in the caller's buffers: K completed iterations rotated the roles K times
    const int K = fin.i - i_start;
    const size_t zb = sizeof(double) * (size_t)nz;
    if (K % 3 == 1) {         // cur = Z[2], prev = Z[0]
        BSLS_CUDA_TRY(cudaMemcpyAsync(Z[1], Z[0], zb, cudaMemcpyDeviceToDevice, st));
        BSLS_CUDA_TRY(cudaMemcpyAsync(Z[0], Z[2], zb, cudaMemcpyDeviceToDevice, st));
    } else if (K % 3 == 2) {  // cur = Z[1], prev = Z[2]
        BSLS_CUDA_TRY(cudaMemcpyAsync(Z[0], Z[1], zb, cudaMemcpyDeviceToDevice, st));
        BSLS_CUDA_TRY(cudaMemcpyAsync(Z[1], Z[2], zb, cudaMemcpyDeviceToDevice, st));
    }
    if (K & 1) BSLS_CUDA_TRY(cudaMemcpyAsync(G[0], G[1], zb, cudaMemcpyDeviceToDevice, st));
    BSLS_CUDA_TRY(cudaEventRecord(q->ev1, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    BSLS_CUDA_TRY(cudaEventElapsedTime(&ms, q->ev0, q->ev1));
    res->f = fin.f;
    res->iterations = fin.i;
    res->stop_code = fin.done;
    res->stop_value = fin.t;
    res->obj_evals = fin.evals;
    res->backtracks = 0;
    res->kernel_launches = (w->launches - launches0) + extra;
    res->device_ms = ms;
    return BSLS_OK;
}

}  // extern "C"
