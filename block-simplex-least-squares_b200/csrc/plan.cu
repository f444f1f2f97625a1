// plan.cu -- device-side analysis of a block layout (run once per layout, not per call).
#include "kernels.h"

namespace bsls {

__global__ void layout_stats_kernel(const int32_t *__restrict__ starts, int nb, int n, LayoutStats *out) {
    int lo = 0x7fffffff, hi = 0, bad = 0;
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
        const int s = starts[b];
        const int e = (b + 1 < nb) ? starts[b + 1] : n;
        const int sz = e - s;
        if (sz <= 0 || s < 0) bad = 1;
        lo = min(lo, sz);
        hi = max(hi, sz);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out->min_size, lo);
        atomicMax(&out->max_size, hi);
        if (bad) atomicOr(&out->bad, 1);
    }
}

static int grid_for(long long items, int threads) {
    long long want = (items + threads - 1) / threads;
    if (want < 1) want = 1;
    return (int)(want < 8 * num_sms() ? want : 8 * num_sms());
}

int plan_layout_stats(const int32_t *starts, int nb, int n, LayoutStats *d_out, cudaStream_t stream) {
    layout_stats_kernel<<<grid_for(nb, 256), 256, 0, stream>>>(starts, nb, n, d_out);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

// tile_first[t] = first block whose start lies in tile t or later; tile_first[ntiles] = nb.
__global__ void tile_first_kernel(const int32_t *__restrict__ starts, int nb, int first, int pitch,
                                  int32_t *__restrict__ tile_first, int ntiles) {
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
        const int t_here = (starts[b] - first) / pitch;
        const int t_prev = b ? (starts[b - 1] - first) / pitch : -1;
        for (int t = t_prev + 1; t <= t_here; ++t) tile_first[t] = (int)b;
        if (b == nb - 1)
            for (int t = t_here + 1; t <= ntiles; ++t) tile_first[t] = nb;
    }
}

int plan_tile_first(const int32_t *starts, int nb, int first, int pitch, int32_t *tile_first, int ntiles, cudaStream_t stream) {
    tile_first_kernel<<<grid_for(nb, 256), 256, 0, stream>>>(starts, nb, first, pitch, tile_first, ntiles);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

// blocks with threshold < size <= upper
__global__ void large_list_kernel(const int32_t *__restrict__ starts, int nb, int threshold, int upper, int32_t *__restrict__ ids, int *count) {
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
        const int sz = starts[b + 1] - starts[b];
        if (sz > threshold && sz <= upper) {
            const int slot = atomicAdd(count, 1);
            if (ids) ids[slot] = (int)b;
        }
    }
}

int plan_large_list(const int32_t *starts, int nb, int threshold, int32_t *ids, int *d_count, cudaStream_t stream, int upper) {
    large_list_kernel<<<grid_for(nb, 256), 256, 0, stream>>>(starts, nb, threshold, upper, ids, d_count);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

// One thread walks the list once per layout: consecutive ids are packed while they fit 32 words / 32 blocks.
__global__ void pack_words_kernel(const int32_t *__restrict__ starts, const int32_t *__restrict__ ids, int count,
                                  int32_t *__restrict__ pack_first, int *npacks) {
    if (threadIdx.x || blockIdx.x) return;
    int acc = 0, np = 0, opened = 0;
    for (int i = 0; i < count; ++i) {
        const int b = ids[i];
        const int words = (starts[b + 1] - starts[b] + 31) >> 5;
        if (i == 0 || acc + words > 32 || i - opened >= 32) {
            pack_first[np++] = i;
            opened = i;
            acc = 0;
        }
        acc += words;
    }
    pack_first[np] = count;
    *npacks = np;
}

int plan_pack_words(const int32_t *starts, const int32_t *ids, int count, int32_t *pack_first, int *d_npacks, cudaStream_t stream) {
    pack_words_kernel<<<1, 32, 0, stream>>>(starts, ids, count, pack_first, d_npacks);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

}  // namespace bsls
