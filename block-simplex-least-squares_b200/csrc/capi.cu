// capi.cu -- the C ABI of libbsls_b200.so (see include/bsls_b200.h).
//
// Host-buffer entry points keep the reference's native signatures
// (python/c_extensions/c_extensions.pyx:15-19,52-61) and do H2D -> kernels -> D2H.
// Device entry points are asynchronous on the caller's stream.  There is no CPU
// fallback anywhere: without an sm_100 device every call fails with BSLS_ERR_NO_DEVICE.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "kernels.h"

namespace bsls {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static thread_local PerDevice<int> sms_pd;
    int &sms = sms_pd.get(0);
    if (sms <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200
    }
    return sms;
}

int device_ok() {
    static thread_local PerDevice<int> cached_pd;
    int &cached = cached_pd.get(-1);
    if (cached == BSLS_OK) return BSLS_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device (%s); libbsls_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        cudaGetLastError();
        return BSLS_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0;
    BSLS_CUDA_TRY(cudaGetDevice(&dev));
    BSLS_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major < 10) {
        set_error("device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
        return BSLS_ERR_NO_DEVICE;
    }
    cached = BSLS_OK;
    return BSLS_OK;
}

}  // namespace bsls

using namespace bsls;


template <typename T>
static int pava_seq(int variant, T *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int min_size, int update, int cold, int clip,
                    cudaStream_t stream) {
    if constexpr (sizeof(T) == 8)
        return pava_seq_f64(variant, (double *)y, w, starts, ids, count, min_size, update, cold, clip, stream);
    else
        return pava_seq_f32(variant, (float *)y, w, starts, ids, count, min_size, update, cold, clip, stream);
}

// variant 1: the parallel kernels (bit-identical to isotonic_regression.h:13-58); blocks beyond their 8192-entry window and
// variants 2 / 3 (isotonic_regression.h:61-82,105-155) run the reference's routine as written, one thread per block.
template <typename T>
static int dev_pava(const bsls_plan *plan, T *y, int32_t *weight, int update, int clip01, cudaStream_t stream, int variant = 1) {
    if (int rc = device_ok()) return rc;
    if (!plan || !y) {
        set_error("isotonic regression: null plan or buffer");
        return BSLS_ERR_ARG;
    }
    if (variant == 2) return pava_seq<T>(2, y, nullptr, plan->d_starts, nullptr, plan->nb, 0, 1, 0, clip01, stream);
    if (variant == 3) {
        if (!weight) {
            set_error("isotonic_regression_3: a weight array is required on device buffers (pass ones, as c_extensions.pyx:118 does)");
            return BSLS_ERR_ARG;
        }
        return pava_seq<T>(3, y, weight, plan->d_starts, nullptr, plan->nb, 0, update, 0, clip01, stream);
    }
    const bool has_long = plan->max_size > kPlanPavaLargeMax;  // blocks the shared-memory kernels cannot hold
    int32_t *seq_w = weight ? weight : plan->d_seq_w;
    if (has_long && !seq_w) {
        set_error("isotonic regression: plan lacks the scratch for blocks longer than %d entries", kPlanPavaLargeMax);
        return BSLS_ERR_ARG;
    }
    const int cta_max = has_long ? kPlanPavaLargeMax : plan->max_size;
    static const int words_min = [] {  // smallest uniform block size the word-per-lane kernel takes (BSLS_PAVA_WORDS_MIN: experiments)
        const char *e = getenv("BSLS_PAVA_WORDS_MIN");
        const int v = e ? atoi(e) : 0;
        return v > 32 ? v : kPlanPavaSmallMax + 1;
    }();
    if (plan->uniform >= words_min && plan->uniform <= kPlanWordsMax) {  // packs of blocks per warp
        const int bpp = 32 / ((plan->uniform + 31) >> 5);
        const int npacks = (plan->nb + bpp - 1) / bpp;
        if constexpr (sizeof(T) == 8)
            return pava_words_f64((double *)y, weight, nullptr, nullptr, nullptr, npacks, plan->first, plan->nb, plan->uniform, update, clip01, 0, stream);
        else
            return pava_words_f32((float *)y, weight, nullptr, nullptr, nullptr, npacks, plan->first, plan->nb, plan->uniform, update, clip01, 0, stream);
    }
    if (plan->uniform > kPlanPavaLargeMax)  // every block is too long for shared memory
        return pava_seq<T>(1, y, seq_w, plan->d_starts, nullptr, plan->nb, 0, update, weight ? 0 : 1, clip01, stream);
    if (plan->uniform > kPlanWordsMax) {  // one CTA per block
        if constexpr (sizeof(T) == 8)
            return pava_words_cta_f64((double *)y, weight, plan->d_starts, nullptr, plan->nb, plan->max_size, update, clip01, 0, stream);
        else
            return pava_words_cta_f32((float *)y, weight, plan->d_starts, nullptr, plan->nb, plan->max_size, update, clip01, 0, stream);
    }
    if (plan->uniform > 0 && plan->uniform <= kPlanPavaSmallMax) {
        if constexpr (sizeof(T) == 8)
            return pava_small_f64((double *)y, weight, plan->first, plan->nb, plan->uniform, update, clip01, stream);
        else
            return pava_small_f32((float *)y, weight, plan->first, plan->nb, plan->uniform, update, clip01, stream);
    }
    // Ragged layout, on the tile grid of the projection: blocks of at most kPlanMidMin entries as rows inside tiles
    // (pava_tile_rows_kernel), up to kPlanTileMaxBlock in packs per warp (d_mid_ids -> pava_words_kernel), longer ones
    // by one CTA each (d_large_ids -> pava_words_cta_kernel).  The three kernels own disjoint blocks: fork onto two
    // auxiliary streams, join.
    BSLS_CUDA_TRY(cudaEventRecord(plan->ev_fork, stream));
    // CTAs per SM of the (tile, words, cta) grids: 0 = as many as fit.  Capping them so that the three kernels are
    // resident side by side measured slower on C3 than letting each fill the GPU (tools/c3_caps.py); BSLS_PAVA_CAPS
    // is kept for that experiment.
    int cap_tile = 0, cap_words = 0, cap_cta = 0;
    if (const char *e = getenv("BSLS_PAVA_CAPS")) sscanf(e, "%d,%d,%d", &cap_tile, &cap_words, &cap_cta);
    int rc = BSLS_OK;
    if (plan->large > 0) {  // the long blocks first: few CTAs with a long critical path each
        BSLS_CUDA_TRY(cudaStreamWaitEvent(plan->aux[1], plan->ev_fork, 0));
        if constexpr (sizeof(T) == 8)
            rc = pava_words_cta_f64((double *)y, weight, plan->d_starts, plan->d_large_ids, plan->large, cta_max, update, clip01, cap_cta, plan->aux[1]);
        else
            rc = pava_words_cta_f32((float *)y, weight, plan->d_starts, plan->d_large_ids, plan->large, cta_max, update, clip01, cap_cta, plan->aux[1]);
        if (rc) return rc;
        if (has_long)  // blocks beyond the window: skipped by the kernel above, served here
            if (int rc2 = pava_seq<T>(1, y, seq_w, plan->d_starts, plan->d_large_ids, plan->large, kPlanPavaLargeMax, update, weight ? 0 : 1, clip01,
                                      plan->aux[1]))
                return rc2;
        BSLS_CUDA_TRY(cudaEventRecord(plan->ev_join[1], plan->aux[1]));
    }
    if (plan->mid > 0) {
        BSLS_CUDA_TRY(cudaStreamWaitEvent(plan->aux[0], plan->ev_fork, 0));
        if constexpr (sizeof(T) == 8)
            rc = pava_words_f64((double *)y, weight, plan->d_starts, plan->d_mid_ids, plan->d_mid_pack, plan->mid_packs, 0, plan->nb, 0, update, clip01, cap_words, plan->aux[0]);
        else
            rc = pava_words_f32((float *)y, weight, plan->d_starts, plan->d_mid_ids, plan->d_mid_pack, plan->mid_packs, 0, plan->nb, 0, update, clip01, cap_words, plan->aux[0]);
        if (rc) return rc;
        BSLS_CUDA_TRY(cudaEventRecord(plan->ev_join[0], plan->aux[0]));
    }
    if constexpr (sizeof(T) == 8)
        rc = pava_tile_f64((double *)y, weight, plan->d_starts, plan->d_tile_first, plan->tiles, update, clip01, cap_tile, stream);
    else
        rc = pava_tile_f32((float *)y, weight, plan->d_starts, plan->d_tile_first, plan->tiles, update, clip01, cap_tile, stream);
    if (rc) return rc;
    if (plan->mid > 0) BSLS_CUDA_TRY(cudaStreamWaitEvent(stream, plan->ev_join[0], 0));
    if (plan->large > 0) BSLS_CUDA_TRY(cudaStreamWaitEvent(stream, plan->ev_join[1], 0));
    return BSLS_OK;
}

// ------------------------------------------------------------------------------------
// device entry points: projections
// ------------------------------------------------------------------------------------
template <typename T> static int dev_project(const bsls_plan *plan, T *y, int mode, cudaStream_t stream) {
    if (int rc = device_ok()) return rc;
    if (!plan || !y) {
        set_error("projection: null plan or buffer");
        return BSLS_ERR_ARG;
    }
    if (plan->uniform > 0 && plan->uniform <= 512) {
        if constexpr (sizeof(T) == 8)
            return proj_uniform_f64((double *)y, plan->first, plan->nb, plan->uniform, mode, plan->d_slow, stream);
        else
            return proj_uniform_f32((float *)y, plan->first, plan->nb, plan->uniform, mode, plan->d_slow, stream);
    }
    if (plan->max_size > kPlanLargeMaxBlock && sizeof(T) != 8) {
        // longer blocks stage only their candidates; the rounding margin of that selection is too wide in fp32
        set_error("projection (fp32): a block of %d entries exceeds the %d-entry limit (use the fp64 entry point)", plan->max_size, kPlanLargeMaxBlock);
        return BSLS_ERR_UNSUPPORTED;
    }
    const int ntiles = plan->ragged ? plan->tiles : 0;
    const int32_t *ids = plan->ragged ? plan->d_large_ids : nullptr;  // uniform large blocks: all of them
    const int nlarge = plan->ragged ? plan->large : plan->nb;
    const RaggedStreams rs = {{plan->aux[0], plan->aux[1]}, plan->ev_fork, {plan->ev_join[0], plan->ev_join[1]}};
    if constexpr (sizeof(T) == 8)
        return proj_ragged_f64((double *)y, plan->d_starts, plan->d_tile_first, ntiles, plan->d_mid_ids, plan->mid, ids, nlarge,
                               plan->max_size, mode, plan->d_slow, plan->nb, rs, plan->d_huge, plan->huge_cap, plan->d_huge_lock, stream);
    else
        return proj_ragged_f32((float *)y, plan->d_starts, plan->d_tile_first, ntiles, plan->d_mid_ids, plan->mid, ids, nlarge,
                               plan->max_size, mode, plan->d_slow, plan->nb, rs, plan->d_huge, plan->huge_cap, plan->d_huge_lock, stream);
}


namespace {

struct HostWorkspace {  // grow-only device staging owned by the calling thread
    double *d_y = nullptr;
    size_t y_cap = 0;
    int32_t *d_blocks = nullptr;
    size_t b_cap = 0;
    int32_t *d_w = nullptr;
    size_t w_cap = 0;
    cudaStream_t stream = nullptr;
    int reserve_w(size_t n) {
        if (n > w_cap) {
            if (d_w) cudaFree(d_w);
            d_w = nullptr;
            w_cap = 0;
            BSLS_CUDA_TRY(cudaMalloc(&d_w, n * sizeof(int32_t)));
            w_cap = n;
        }
        return BSLS_OK;
    }
    int reserve(size_t n, size_t nb) {
        if (!stream) BSLS_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (n > y_cap) {
            if (d_y) cudaFree(d_y);
            d_y = nullptr;
            y_cap = 0;
            BSLS_CUDA_TRY(cudaMalloc(&d_y, n * sizeof(double)));
            y_cap = n;
        }
        if (nb > b_cap) {
            if (d_blocks) cudaFree(d_blocks);
            d_blocks = nullptr;
            b_cap = 0;
            BSLS_CUDA_TRY(cudaMalloc(&d_blocks, nb * sizeof(int32_t)));
            b_cap = nb;
        }
        return BSLS_OK;
    }
};
thread_local PerDevice<HostWorkspace> g_ws_pd;  // staging is device memory: one set per device
#define g_ws (g_ws_pd.get())

// Ragged layouts on host buffers: the analysed layout of the previous call is kept (per calling thread) and reused when
// the block starts are the same -- solvers call with one layout thousands of times, and bsls_plan_create synchronises.
struct HostPlanCache {
    bsls_plan *plan = nullptr;
    int nb = 0, device = -1;
    long long span = 0;
    uint64_t hash = 0;
    int get(const int *blocks, int numblocks, int first, long long span_, uint64_t hash_, cudaStream_t st, bsls_plan **out) {
        int dev = 0;
        BSLS_CUDA_TRY(cudaGetDevice(&dev));
        if (plan && nb == numblocks && span == span_ && hash == hash_ && device == dev) {
            *out = plan;
            return BSLS_OK;
        }
        if (plan) bsls_plan_destroy(plan);
        plan = nullptr;
        std::vector<int32_t> rebased((size_t)numblocks);  // the staged span begins at 0
        for (int i = 0; i < numblocks; ++i) rebased[i] = blocks[i] - first;
        BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_blocks, rebased.data(), (size_t)numblocks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (int rc = bsls_plan_create(g_ws.d_blocks, numblocks, (int)span_, st, &plan)) return rc;  // synchronises st: `rebased` may go
        nb = numblocks;
        span = span_;
        hash = hash_;
        device = dev;
        *out = plan;
        return BSLS_OK;
    }
};
thread_local PerDevice<HostPlanCache> g_plan_cache_pd;
#define g_plan_cache (g_plan_cache_pd.get())

// the reference's Python-side asserts (c_extensions.pyx:33-34), checked on the host copy; the same pass finds the
// common block size.  *uniform = K when every block has K entries, else 0.
int validate_blocks(const int *blocks, int numblocks, int n, int *uniform = nullptr, uint64_t *hash = nullptr) {
    if (!blocks || numblocks <= 0 || n <= 0) {
        set_error("need numblocks > 0 and n > 0");
        return BSLS_ERR_ARG;
    }
    if (blocks[0] < 0 || blocks[numblocks - 1] >= n) {
        set_error("block starts out of range");
        return BSLS_ERR_ARG;
    }
    const int K = (numblocks > 1 ? blocks[1] : n) - blocks[0];
    int bad = 0, diff = 0;  // branch-free scan (vectorised); the position is looked up only on failure
    uint64_t h = (uint64_t)(uint32_t)blocks[0] * 0x9e3779b97f4a7c15ull;
    for (int i = 1; i < numblocks; ++i) {
        const int d = blocks[i] - blocks[i - 1];
        bad |= (d <= 0);
        diff |= d ^ K;
        h += (uint64_t)(uint32_t)d * ((uint64_t)i * 0x9e3779b97f4a7c15ull + 0x7f4a7c15ull);  // position-weighted sum of the sizes
    }
    if (hash) *hash = h ^ ((uint64_t)(uint32_t)numblocks << 32) ^ (uint64_t)(uint32_t)n;
    if (bad) {
        int i = 1;
        while (i < numblocks && blocks[i] > blocks[i - 1]) ++i;
        set_error("block starts not strictly increasing at %d", i);
        return BSLS_ERR_ARG;
    }
    if (uniform) *uniform = (diff == 0 && n - blocks[numblocks - 1] == K) ? K : 0;
    return BSLS_OK;
}

// Uniform layouts need no plan, which lets the host entry points pipeline: the span is cut into
// chunks of whole blocks and chunk c's upload, kernel and download run on stream c mod 3, so the
// two copy engines and the SMs work at the same time (PCIe is the bound of these entry points).
constexpr int kPipeStreams = 3;
struct HostPipeline {
    cudaStream_t s[kPipeStreams] = {nullptr, nullptr, nullptr};
    int32_t *slow[kPipeStreams] = {nullptr, nullptr, nullptr};
    size_t slow_cap = 0;
    int prepare(size_t chunk_blocks) {
        for (int k = 0; k < kPipeStreams; ++k)
            if (!s[k]) BSLS_CUDA_TRY(cudaStreamCreateWithFlags(&s[k], cudaStreamNonBlocking));
        if (chunk_blocks + 1 > slow_cap) {
            for (int k = 0; k < kPipeStreams; ++k) {
                if (slow[k]) cudaFree(slow[k]);
                slow[k] = nullptr;
                BSLS_CUDA_TRY(cudaMalloc(&slow[k], (chunk_blocks + 1) * sizeof(int32_t)));
            }
            slow_cap = chunk_blocks + 1;
        }
        return BSLS_OK;
    }
    int sync() {
        for (int k = 0; k < kPipeStreams; ++k) BSLS_CUDA_TRY(cudaStreamSynchronize(s[k]));
        return BSLS_OK;
    }
};
thread_local PerDevice<HostPipeline> g_pipe_pd;
#define g_pipe (g_pipe_pd.get())

size_t pipeline_chunk_blocks(int K, int numblocks) {
    // chunk size: 1/16 of the span, between 8 MB (short fill / drain for small arrays) and 32 MB (long transfers keep
    // both copy engines nearer the link rate: 512 MB each way took 14.0 ms in 8 MB chunks, 12.4 ms in 32 MB chunks);
    // BSLS_PIPE_MB overrides, for tuning
    static const size_t forced_mb = [] {
        const char *e = getenv("BSLS_PIPE_MB");
        const int v = e ? atoi(e) : 0;
        return (size_t)(v > 0 ? v : 0);
    }();
    const size_t span_bytes = (size_t)numblocks * K * sizeof(double);
    size_t bytes = span_bytes / 16;
    if (bytes < ((size_t)8 << 20)) bytes = (size_t)8 << 20;
    if (bytes > ((size_t)32 << 20)) bytes = (size_t)32 << 20;
    if (forced_mb) bytes = forced_mb << 20;
    size_t cb = bytes / ((size_t)K * sizeof(double));
    if (cb < 1024) cb = 1024;
    if (cb > (size_t)numblocks) cb = (size_t)numblocks;
    return cb;
}

int host_project_uniform(double *y, int first, int numblocks, int K, int mode) {
    const size_t span = (size_t)numblocks * K;
    if (int rc = g_ws.reserve(span, 1)) return rc;
    const size_t cb = pipeline_chunk_blocks(K, numblocks);
    if (int rc = g_pipe.prepare(cb)) return rc;
    auto chunk = [&](size_t b0, int c) -> int {
        const size_t nbc = std::min(cb, (size_t)numblocks - b0);
        const size_t off = b0 * K, cnt = nbc * K;
        cudaStream_t st = g_pipe.s[c % kPipeStreams];
        BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_y + off, y + first + off, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
        if (int rc = proj_uniform_f64(g_ws.d_y, (long long)off, (int)nbc, K, mode, g_pipe.slow[c % kPipeStreams], st)) return rc;
        BSLS_CUDA_TRY(cudaMemcpyAsync(y + first + off, g_ws.d_y + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
        return BSLS_OK;
    };
    int c = 0;
    for (size_t b0 = 0; b0 < (size_t)numblocks; b0 += cb, ++c)
        if (int rc = chunk(b0, c)) {
            g_pipe.sync();  // copies into the caller's buffer may still be in flight
            return rc;
        }
    return g_pipe.sync();
}

int host_project(double *y, const int *blocks, int numblocks, int n, int mode) {
    if (!y) {
        set_error("null buffer");
        return BSLS_ERR_ARG;
    }
    int K = 0;
    uint64_t hash = 0;
    if (int rc = validate_blocks(blocks, numblocks, n, &K, &hash)) return rc;  // argument errors first, as the reference's asserts
    if (int rc = device_ok()) return rc;
    const int first = blocks[0];
    const size_t span = (size_t)n - first;  // entries before blocks[0] never leave the host
    if (K > 0 && K <= 512) return host_project_uniform(y, first, numblocks, K, mode);
    if (int rc = g_ws.reserve(span, (size_t)numblocks)) return rc;
    cudaStream_t st = g_ws.stream;
    BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_y, y + first, span * sizeof(double), cudaMemcpyHostToDevice, st));
    bsls_plan *plan = nullptr;
    if (int rc = g_plan_cache.get(blocks, numblocks, first, (long long)span, hash, st, &plan)) return rc;
    if (int rc = dev_project<double>(plan, g_ws.d_y, mode, st)) return rc;
    BSLS_CUDA_TRY(cudaMemcpyAsync(y + first, g_ws.d_y, span * sizeof(double), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    return BSLS_OK;
}

// isotonic regression on host buffers; `weight` may be NULL (all ones in, result dropped)
int host_pava(double *y, const int *blocks, int numblocks, int n, int *weight, int update, int variant = 1) {
    if (!y) {
        set_error("null buffer");
        return BSLS_ERR_ARG;
    }
    int K = 0;
    uint64_t hash = 0;
    if (int rc = validate_blocks(blocks, numblocks, n, &K, &hash)) return rc;
    if (int rc = device_ok()) return rc;
    const int first = blocks[0];
    const size_t span = (size_t)n - first;
    if (int rc = g_ws.reserve(span, (size_t)numblocks)) return rc;
    if (weight)
        if (int rc = g_ws.reserve_w(span)) return rc;
    if (variant == 1 && K > 0 && K <= kPlanPavaSmallMax) {
        // pipelined like host_project_uniform: chunks of whole blocks on three streams
        const size_t cb = pipeline_chunk_blocks(K, numblocks);
        if (int rc = g_pipe.prepare(cb)) return rc;
        auto chunk = [&](size_t b0, int c) -> int {
            const size_t nbc = std::min(cb, (size_t)numblocks - b0);
            const size_t off = b0 * K, cnt = nbc * K;
            cudaStream_t st = g_pipe.s[c % kPipeStreams];
            BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_y + off, y + first + off, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
            if (weight) BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_w + off, weight + first + off, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            if (int rc = pava_small_f64(g_ws.d_y, weight ? g_ws.d_w : nullptr, (long long)off, (int)nbc, K, update, 0, st)) return rc;
            BSLS_CUDA_TRY(cudaMemcpyAsync(y + first + off, g_ws.d_y + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
            if (weight) BSLS_CUDA_TRY(cudaMemcpyAsync(weight + first + off, g_ws.d_w + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            return BSLS_OK;
        };
        int c = 0;
        for (size_t b0 = 0; b0 < (size_t)numblocks; b0 += cb, ++c)
            if (int rc = chunk(b0, c)) {
                g_pipe.sync();
                return rc;
            }
        return g_pipe.sync();
    }
    cudaStream_t st = g_ws.stream;
    BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_y, y + first, span * sizeof(double), cudaMemcpyHostToDevice, st));
    if (weight) BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_w, weight + first, span * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    int32_t *dw = weight ? g_ws.d_w : nullptr;
    if (variant == 3 && !weight) {  // weight=None of the Python layer: ones in, result dropped (c_extensions.pyx:118-119)
        if (int rc = g_ws.reserve_w(span)) return rc;
        std::vector<int32_t> ones(span, 1);
        BSLS_CUDA_TRY(cudaMemcpyAsync(g_ws.d_w, ones.data(), span * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        BSLS_CUDA_TRY(cudaStreamSynchronize(st));  // `ones` leaves scope
        dw = g_ws.d_w;
    }
    bsls_plan *plan = nullptr;
    if (int rc = g_plan_cache.get(blocks, numblocks, first, (long long)span, hash, st, &plan)) return rc;
    if (int rc = dev_pava<double>(plan, g_ws.d_y, dw, update, 0, st, variant)) return rc;
    BSLS_CUDA_TRY(cudaMemcpyAsync(y + first, g_ws.d_y, span * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (weight) BSLS_CUDA_TRY(cudaMemcpyAsync(weight + first, g_ws.d_w, span * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BSLS_CUDA_TRY(cudaStreamSynchronize(st));
    return BSLS_OK;
}

int host_pava_single(double *y, int start, int end, int *weight, int update, int variant = 1) {
    if (start >= end) return BSLS_OK;  // c_extensions.pyx:66
    if (start < 0) {
        set_error("isotonic_regression: start < 0");
        return BSLS_ERR_ARG;
    }
    const int one = start;
    return host_pava(y, &one, 1, end, weight, update, variant);
}

}  // namespace

namespace bsls {
int project_f64(const bsls_plan *plan, double *y, int mode, cudaStream_t stream) { return dev_project<double>(plan, y, mode, stream); }
int project_step_f64(const bsls_plan *plan, const double *x, const double *g, double t, double *x_new, int mode, cudaStream_t stream, bool *fused,
                     const StepCtl *ctl) {
    *fused = false;
    if (!plan || !x || !g || !x_new) return BSLS_ERR_ARG;
    if (plan->first != 0 || plan->uniform <= 0 || plan->uniform > 512 || !proj_step_fuses(plan->uniform)) return BSLS_OK;
    *fused = true;
    return proj_step_uniform_f64(x, g, t, x_new, 0, plan->nb, plan->uniform, mode, stream, ctl);
}
int pava_clip_f64(const bsls_plan *plan, double *y, int32_t *weight, int update, int clip01, cudaStream_t stream) {
    return dev_pava<double>(plan, y, weight, update, clip01, stream);
}
}  // namespace bsls

extern "C" {

const char *bsls_last_error(void) { return g_err; }
const char *bsls_version(void) { return "bsls_b200 0.1 (sm_100a)"; }
int bsls_device_ok(void) { return device_ok(); }

// ------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------
int bsls_plan_create(const int32_t *d_blocks, int numblocks, int n, bsls_stream_t stream_, bsls_plan **out) {
    if (!out) return BSLS_ERR_ARG;
    *out = nullptr;
    if (int rc = device_ok()) return rc;
    if (!d_blocks || numblocks <= 0 || n <= 0) {
        set_error("plan_create: need numblocks > 0 and n > 0 (got %d, %d)", numblocks, n);
        return BSLS_ERR_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    bsls_plan *p = new bsls_plan();
    p->nb = numblocks;
    p->n = n;
    LayoutStats *d_stats = nullptr;
    LayoutStats h_stats = {0x7fffffff, 0, 0};
    int first = 0;
    auto fail = [&](int rc) {
        if (d_stats) cudaFree(d_stats);
        bsls_plan_destroy(p);
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
    TRY_OR_FAIL(cudaMalloc(&p->d_starts, sizeof(int32_t) * ((size_t)numblocks + 1)));
    TRY_OR_FAIL(cudaMalloc(&d_stats, sizeof(LayoutStats)));
    TRY_OR_FAIL(cudaMemcpyAsync(p->d_starts, d_blocks, sizeof(int32_t) * (size_t)numblocks, cudaMemcpyDeviceToDevice, stream));
    TRY_OR_FAIL(cudaMemcpyAsync(p->d_starts + numblocks, &n, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    TRY_OR_FAIL(cudaMemcpyAsync(d_stats, &h_stats, sizeof(LayoutStats), cudaMemcpyHostToDevice, stream));
    if (int rc = plan_layout_stats(p->d_starts, numblocks, n, d_stats, stream)) return fail(rc);
    TRY_OR_FAIL(cudaMemcpyAsync(&h_stats, d_stats, sizeof(LayoutStats), cudaMemcpyDeviceToHost, stream));
    TRY_OR_FAIL(cudaMemcpyAsync(&first, p->d_starts, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    TRY_OR_FAIL(cudaStreamSynchronize(stream));
    cudaFree(d_stats);
    d_stats = nullptr;
    if (h_stats.bad || first < 0 || first >= n) {
        // the reference's asserts (c_extensions.pyx:33-34): strictly increasing, within [0, n)
        set_error("plan_create: block starts must be strictly increasing with 0 <= blocks[0] and blocks[-1] < n");
        return fail(BSLS_ERR_ARG);
    }
    p->first = first;
    p->min_size = h_stats.min_size;
    p->max_size = h_stats.max_size;
    p->uniform = (h_stats.min_size == h_stats.max_size) ? h_stats.min_size : 0;
    if (!p->uniform) {
        // ragged: fixed tile grid over [first, n) + list of blocks too long for a tile
        p->ragged = true;
        p->tiles = (int)(((long long)n - first + kPlanTileElems - 1) / kPlanTileElems);
        int *d_count = nullptr;
        int h_count = 0;
        TRY_OR_FAIL(cudaMalloc(&p->d_tile_first, sizeof(int32_t) * ((size_t)p->tiles + 1)));
        TRY_OR_FAIL(cudaMalloc(&d_count, sizeof(int)));
        TRY_OR_FAIL(cudaMemsetAsync(d_count, 0, sizeof(int), stream));
        if (int rc = plan_tile_first(p->d_starts, numblocks, first, kPlanTileElems, p->d_tile_first, p->tiles, stream)) return fail(rc);
        if (p->max_size > kPlanTileMaxBlock) {
            if (int rc = plan_large_list(p->d_starts, numblocks, kPlanTileMaxBlock, nullptr, d_count, stream)) return fail(rc);
            TRY_OR_FAIL(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            TRY_OR_FAIL(cudaStreamSynchronize(stream));
            TRY_OR_FAIL(cudaMalloc(&p->d_large_ids, sizeof(int32_t) * (size_t)h_count));
            TRY_OR_FAIL(cudaMemsetAsync(d_count, 0, sizeof(int), stream));
            if (int rc = plan_large_list(p->d_starts, numblocks, kPlanTileMaxBlock, p->d_large_ids, d_count, stream)) return fail(rc);
            p->large = h_count;
        }
        if (p->max_size > kPlanMidMin) {
            int h_mid = 0;
            TRY_OR_FAIL(cudaMemsetAsync(d_count, 0, sizeof(int), stream));
            if (int rc = plan_large_list(p->d_starts, numblocks, kPlanMidMin, nullptr, d_count, stream, kPlanTileMaxBlock)) return fail(rc);
            TRY_OR_FAIL(cudaMemcpyAsync(&h_mid, d_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            TRY_OR_FAIL(cudaStreamSynchronize(stream));
            if (h_mid > 0) {
                TRY_OR_FAIL(cudaMalloc(&p->d_mid_ids, sizeof(int32_t) * (size_t)h_mid));
                TRY_OR_FAIL(cudaMemsetAsync(d_count, 0, sizeof(int), stream));
                if (int rc = plan_large_list(p->d_starts, numblocks, kPlanMidMin, p->d_mid_ids, d_count, stream, kPlanTileMaxBlock)) return fail(rc);
            }
            p->mid = h_mid;
        }
        if (p->mid > 0) {  // packs of the mid list for the word-per-lane isotonic regression
            int h_np = 0;
            TRY_OR_FAIL(cudaMalloc(&p->d_mid_pack, sizeof(int32_t) * ((size_t)p->mid + 1)));
            if (int rc = plan_pack_words(p->d_starts, p->d_mid_ids, p->mid, p->d_mid_pack, d_count, stream)) return fail(rc);
            TRY_OR_FAIL(cudaMemcpyAsync(&h_np, d_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            TRY_OR_FAIL(cudaStreamSynchronize(stream));
            p->mid_packs = h_np;
        }
        TRY_OR_FAIL(cudaStreamSynchronize(stream));
        cudaFree(d_count);
    }
    // Everything a call may need is allocated HERE: the plan is immutable afterwards.  Its scratch (queue of dense blocks,
    // candidate buffer of giant blocks, auxiliary streams) is shared by all calls on the plan, hence "one stream at a
    // time per plan" (include/bsls_b200.h).
    if (p->ragged || p->uniform > 16)  // queue between the selection kernels and the sorter
        TRY_OR_FAIL(cudaMalloc(&p->d_slow, sizeof(int32_t) * ((size_t)numblocks + 1)));
    if (p->ragged || p->uniform > 512) {  // fork/join of kernels that own disjoint blocks
        TRY_OR_FAIL(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            TRY_OR_FAIL(cudaStreamCreateWithFlags(&p->aux[k], cudaStreamNonBlocking));
            TRY_OR_FAIL(cudaEventCreateWithFlags(&p->ev_join[k], cudaEventDisableTiming));
        }
    }
    if (p->max_size > kPlanLargeMaxBlock) {  // projection: candidates of blocks beyond the shared-memory window
        int cap = 1;
        while (cap < p->max_size) cap <<= 1;
        TRY_OR_FAIL(cudaMalloc(&p->d_huge, (size_t)2 * cap * sizeof(double)));
        TRY_OR_FAIL(cudaMalloc(&p->d_huge_lock, sizeof(int)));
        TRY_OR_FAIL(cudaMemsetAsync(p->d_huge_lock, 0, sizeof(int), stream));
        p->huge_cap = cap;
    }
    if (p->max_size > kPlanPavaLargeMax)  // isotonic regression of such blocks without a caller weight array
        TRY_OR_FAIL(cudaMalloc(&p->d_seq_w, sizeof(int32_t) * (size_t)n));
    TRY_OR_FAIL(cudaStreamSynchronize(stream));
#undef TRY_OR_FAIL
    *out = p;
    return BSLS_OK;
}

int bsls_plan_destroy(bsls_plan *plan) {
    if (!plan) return BSLS_OK;
    if (plan->d_starts) cudaFree(plan->d_starts);
    if (plan->d_tile_first) cudaFree(plan->d_tile_first);
    if (plan->d_large_ids) cudaFree(plan->d_large_ids);
    if (plan->d_mid_pack) cudaFree(plan->d_mid_pack);
    if (plan->d_slow) cudaFree(plan->d_slow);
    if (plan->d_huge) cudaFree(plan->d_huge);
    if (plan->d_huge_lock) cudaFree(plan->d_huge_lock);
    if (plan->d_seq_w) cudaFree(plan->d_seq_w);
    if (plan->d_mid_ids) cudaFree(plan->d_mid_ids);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    for (int k = 0; k < 2; ++k) {
        if (plan->ev_join[k]) cudaEventDestroy(plan->ev_join[k]);
        if (plan->aux[k]) cudaStreamDestroy(plan->aux[k]);
    }
    delete plan;
    return BSLS_OK;
}

int bsls_plan_info(const bsls_plan *plan, int64_t info[8]) {
    if (!plan || !info) return BSLS_ERR_ARG;
    info[0] = plan->nb;
    info[1] = plan->n;
    info[2] = plan->first;
    info[3] = plan->uniform;
    info[4] = plan->min_size;
    info[5] = plan->max_size;
    info[6] = plan->tiles;
    info[7] = plan->large;
    return BSLS_OK;
}

// device entry points: projections
int bsls_dev_proj_multi_simplex_f64(const bsls_plan *plan, double *y, bsls_stream_t s) { return dev_project<double>(plan, y, 0, (cudaStream_t)s); }
int bsls_dev_proj_multi_ball_f64(const bsls_plan *plan, double *y, bsls_stream_t s) { return dev_project<double>(plan, y, 1, (cudaStream_t)s); }
int bsls_dev_proj_multi_simplex_f32(const bsls_plan *plan, float *y, bsls_stream_t s) { return dev_project<float>(plan, y, 0, (cudaStream_t)s); }
int bsls_dev_proj_multi_ball_f32(const bsls_plan *plan, float *y, bsls_stream_t s) { return dev_project<float>(plan, y, 1, (cudaStream_t)s); }

// ------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------
int bsls_proj_simplex(double *y, int start, int end) {
    // single block [start, end): argument checks are the caller's (c_extensions.pyx:24-25);
    // an empty range is a no-op exactly as there
    if (start >= end) return BSLS_OK;
    if (start < 0) {
        set_error("proj_simplex: start < 0");
        return BSLS_ERR_ARG;
    }
    const int one = start;
    return host_project(y, &one, 1, end, 0);
}

int bsls_proj_multi_simplex(double *y, const int *blocks, int numblocks, int n) { return host_project(y, blocks, numblocks, n, 0); }
int bsls_proj_multi_ball(double *y, const int *blocks, int numblocks, int n) { return host_project(y, blocks, numblocks, n, 1); }

// isotonic regression, host buffers.  Variant 1 runs the parallel kernels; variants 2 and 3 (a different order of
// operations and, for 3, a different weight array: isotonic_regression.h:61-82,105-155) run the reference's routines as
// written, one thread per block (pava_seq.cuh), so that values AND weights are the reference's bits.
int bsls_isotonic_regression(double *y, int start, int end, int *weight, int update) { return host_pava_single(y, start, end, weight, update); }
int bsls_isotonic_regression_multi(double *y, const int *blocks, int numblocks, int n, int *weight, int update) {
    return host_pava(y, blocks, numblocks, n, weight, update);
}
int bsls_isotonic_regression_2(double *y, int start, int end) { return host_pava_single(y, start, end, nullptr, 1, 2); }
int bsls_isotonic_regression_multi_2(double *y, const int *blocks, int numblocks, int n) { return host_pava(y, blocks, numblocks, n, nullptr, 1, 2); }
int bsls_isotonic_regression_3(double *y, int start, int end, int *weight, int update) { return host_pava_single(y, start, end, weight, update, 3); }
int bsls_isotonic_regression_multi_3(double *y, const int *blocks, int numblocks, int n, int *weight, int update) {
    return host_pava(y, blocks, numblocks, n, weight, update, 3);
}

int bsls_dev_isotonic_regression_multi_2_f64(const bsls_plan *plan, double *y, bsls_stream_t s) {
    return dev_pava<double>(plan, y, nullptr, 1, 0, (cudaStream_t)s, 2);
}
int bsls_dev_isotonic_regression_multi_3_f64(const bsls_plan *plan, double *y, int32_t *weight, int update, bsls_stream_t s) {
    return dev_pava<double>(plan, y, weight, update, 0, (cudaStream_t)s, 3);
}

int bsls_dev_isotonic_regression_multi_f64(const bsls_plan *plan, double *y, int32_t *weight, int update, int clip01, bsls_stream_t s) {
    return dev_pava<double>(plan, y, weight, update, clip01, (cudaStream_t)s);
}
int bsls_dev_isotonic_regression_multi_f32(const bsls_plan *plan, float *y, int32_t *weight, int update, int clip01, bsls_stream_t s) {
    return dev_pava<float>(plan, y, weight, update, clip01, (cudaStream_t)s);
}

int bsls_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return BSLS_ERR_ARG;
    if (int rc = device_ok()) return rc;
    BSLS_CUDA_TRY(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return BSLS_OK;
}

int bsls_host_free(void *ptr) {
    if (!ptr) return BSLS_OK;
    BSLS_CUDA_TRY(cudaFreeHost(ptr));
    return BSLS_OK;
}

}  // extern "C"
