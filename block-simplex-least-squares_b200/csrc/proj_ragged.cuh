// proj_ragged.cuh -- segmented simplex / l1-ball projection for blocks of mixed sizes.
//
// Replaces proj_multi_simplex / proj_multi_ball (python/c_extensions/proj_simplex.h:37-74)
// for ragged layouts (BASELINE config 3: power-law sizes 2..4096).
//
// Work split (decided once per layout, see plan.cu):
//   * TILE kernel: the element range is cut into fixed tiles of kTileElems; a CTA owns the
//     blocks that START in its tile (at most kTileMaxBlock long, so they fit the CTA's
//     shared-memory window of kTileElems + kTileMaxBlock values).  Inside the CTA blocks
//     are binned by size class and every class is processed by the register code of
//     simplex_core.cuh with G = 1..32 lanes per block.  HBM traffic: read y once, write y
//     once, read the int32 starts once.
//   * LARGE kernel: one CTA per block longer than kTileMaxBlock (up to kLargeMaxBlock):
//     bitonic sort in shared memory, the reference's left-to-right running sum by one
//     thread, candidates tested by all threads.
// Both reproduce the reference's arithmetic order, so results are bit-identical to it.
#pragma once
#include "proj_uniform.cuh"

namespace bsls {

constexpr int kTileElems = 2048;     // tile grid pitch (elements)
constexpr int kTileMaxBlock = 512;   // longest block the tile kernel handles
constexpr int kTileThreads = 256;
constexpr int kLargeMaxBlock = 8192; // longest block the one-CTA kernel handles
constexpr int kLargeThreads = 512;
constexpr int kNumClasses = 8;

__device__ __forceinline__ int size_class(int K) {
    // 0:<=4 1:<=8 2:<=16 3:<=32 4:<=64 5:<=128 6:<=256 7:<=512
    return K <= 4 ? 0 : (30 - __clz(K - 1));  // ceil(log2 K) - 2 for K > 4
}

// One size class: G lanes per block, E registers per lane.
template <typename T, int E, int G, int MODE>
__device__ __noinline__ void process_class(T *ybuf, const int *sstart, const uint16_t *list, int count, int tile_lo) {
    constexpr int GROUPS = kTileThreads / G;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int sub = lane & (G - 1);
    const int grp = tid / G;
    for (int base = 0; base < count; base += GROUPS) {  // uniform trip count across the CTA
        const int idx = base + grp;
        const bool live = idx < count;
        int K = 1;
        T *blkp = ybuf;
        if (live) {
            const int b = list[idx];
            const int s = sstart[b];
            K = sstart[b + 1] - s;
            blkp = ybuf + (s - tile_lo);
        }
        bool project = true;
        if (MODE == kBall) {
            T total = T(0);
            if (live && sub == 0)
                for (int k = 0; k < K; ++k) {
                    const T x = blkp[k];
                    if (!(x < T(0))) total += x;
                }
            if (G > 1) total = __shfl_sync(0xffffffffu, total, lane & ~(G - 1));
            project = total > T(1);
        }
        T v[E];
        load_block_regs<T, E, G, MODE>(v, blkp, K, lane, live, false);
        sort_desc_group<T, E, G>(v, lane);
        T shift = simplex_shift_sorted<T, E, G>(v, K, lane);
        if (MODE == kBall && !project) shift = T(0);
        // every lane rewrites exactly the elements it loaded (same rotation as the load)
        if (live) {
            int q = lane % K;
            const int step = G % K;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int slot = e * G + sub;
                if (slot < K) {
                    T x = blkp[q];
                    if (MODE == kBall) x = clip_neg(x);
                    x = shift + x;
                    blkp[q] = (x < T(0)) ? T(0) : x;
                }
                q += step;
                if (q >= K) q -= K;
            }
        }
    }
}

// tile_first[t] = index of the first block whose start lies in tile t (tile_first[ntiles] = nb)
template <typename T, int MODE>
__global__ void __launch_bounds__(kTileThreads, 2)
proj_tile_kernel(T *__restrict__ y, const int32_t *__restrict__ starts /* nb+1, last = n */,
                 const int32_t *__restrict__ tile_first, int ntiles) {
    __shared__ __align__(16) T ybuf[kTileElems + kTileMaxBlock];
    __shared__ int sstart[kTileElems + 1];
    __shared__ uint16_t list[kTileElems];
    __shared__ int cnt[kNumClasses], off[kNumClasses + 1], fill[kNumClasses];

    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int fb = tile_first[tile];
        const int nblk = tile_first[tile + 1] - fb;
        if (nblk <= 0) continue;  // uniform across the CTA
        if (tid < kNumClasses) {
            cnt[tid] = 0;
            fill[tid] = 0;
        }
        for (int i = tid; i <= nblk; i += kTileThreads) sstart[i] = starts[fb + i];
        __syncthreads();
        const int tile_lo = sstart[0];
        // the window ends with the last block that is not "large"
        int tile_hi = sstart[nblk];
        if (tile_hi - sstart[nblk - 1] > kTileMaxBlock) tile_hi = sstart[nblk - 1];
        const int nel = tile_hi - tile_lo;
        for (int i = tid; i < nel; i += kTileThreads) ybuf[i] = y[(size_t)tile_lo + i];
        // ---- bin the blocks by size class (counting sort on shared counters) ---------------
        for (int i = tid; i < nblk; i += kTileThreads) {
            const int K = sstart[i + 1] - sstart[i];
            if (K <= kTileMaxBlock) atomicAdd(&cnt[size_class(K)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int c = 0; c < kNumClasses; ++c) {
                off[c] = acc;
                acc += cnt[c];
            }
            off[kNumClasses] = acc;
        }
        __syncthreads();
        for (int i = tid; i < nblk; i += kTileThreads) {
            const int K = sstart[i + 1] - sstart[i];
            if (K <= kTileMaxBlock) {
                const int c = size_class(K);
                list[off[c] + atomicAdd(&fill[c], 1)] = (uint16_t)i;
            }
        }
        __syncthreads();
        // ---- per class ------------------------------------------------------------------------
        process_class<T, 4, 1, MODE>(ybuf, sstart, list + off[0], cnt[0], tile_lo);
        process_class<T, 8, 1, MODE>(ybuf, sstart, list + off[1], cnt[1], tile_lo);
        process_class<T, 16, 1, MODE>(ybuf, sstart, list + off[2], cnt[2], tile_lo);
        process_class<T, 32, 1, MODE>(ybuf, sstart, list + off[3], cnt[3], tile_lo);
        process_class<T, 16, 4, MODE>(ybuf, sstart, list + off[4], cnt[4], tile_lo);
        process_class<T, 16, 8, MODE>(ybuf, sstart, list + off[5], cnt[5], tile_lo);
        process_class<T, 16, 16, MODE>(ybuf, sstart, list + off[6], cnt[6], tile_lo);
        process_class<T, 16, 32, MODE>(ybuf, sstart, list + off[7], cnt[7], tile_lo);
        __syncthreads();
        // ---- coalesced write-back (large blocks inside the window are rewritten unchanged;
        //      the LARGE kernel runs afterwards on the same stream) ----------------------------
        for (int i = tid; i < nel; i += kTileThreads) y[(size_t)tile_lo + i] = ybuf[i];
        __syncthreads();
    }
}

// ---- one CTA per large block ---------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kLargeThreads)
proj_large_kernel(T *__restrict__ y, const int32_t *__restrict__ starts, const int32_t *__restrict__ ids, int count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_last;
    __shared__ T s_shift, s_total;
    const int tid = threadIdx.x;
    for (int it = blockIdx.x; it < count; it += gridDim.x) {
        const int b = ids ? ids[it] : it;
        const int lo = starts[b];
        const int K = starts[b + 1] - lo;
        int KP = 1;
        while (KP < K) KP <<= 1;
        T *srt = reinterpret_cast<T *>(smem_raw);  // KP sorted values
        T *pre = srt + KP;                         // KP running sums
        T *gy = y + (size_t)lo;
        const T ninf = Num<T>::neg_inf();
        for (int i = tid; i < KP; i += kLargeThreads) {
            T x = ninf;
            if (i < K) {
                x = gy[i];
                if (MODE == kBall) x = clip_neg(x);
            }
            srt[i] = x;
        }
        if (tid == 0) s_last = 0;
        if (MODE == kBall) {
            if (tid == 0) {  // the reference's left-to-right sum over the kept entries
                T total = T(0);
                for (int k = 0; k < K; ++k) {
                    const T x = gy[k];
                    if (!(x < T(0))) total += x;
                }
                s_total = total;
            }
        }
        __syncthreads();
        const bool project = (MODE == kBall) ? (s_total > T(1)) : true;
        if (project) {
            // bitonic sort, descending, -inf sentinels at the tail
            for (int size = 2; size <= KP; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int t = tid; t < (KP >> 1); t += kLargeThreads) {
                        const int i = 2 * t - (t & (stride - 1));
                        const int j = i + stride;
                        const bool desc = (i & size) == 0;
                        const T a = srt[i], c = srt[j];
                        if ((a < c) == desc) {
                            srt[i] = c;
                            srt[j] = a;
                        }
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) {  // running sum, strictly left to right (proj_simplex.h:24-28)
                T run = srt[0];
                pre[0] = run;
                for (int k = 1; k < K; ++k) {
                    run += srt[k];
                    pre[k] = run;
                }
            }
            __syncthreads();
            int mine = 0;  // last passing position among this thread's candidates
            for (int k = tid; k < K; k += kLargeThreads) {
                if (k == 0) continue;
                const T cand = (T(1) - pre[k]) / (T(k) + T(1));
                if (srt[k] + cand > T(0)) mine = k;
            }
            if (mine) atomicMax(&s_last, mine);
            __syncthreads();
            if (tid == 0) {
                const int k = s_last;
                s_shift = (k == 0) ? (T(1) - pre[0]) : (T(1) - pre[k]) / (T(k) + T(1));
            }
            __syncthreads();
        }
        const T shift = project ? s_shift : T(0);
        for (int i = tid; i < K; i += kLargeThreads) {
            T x = gy[i];
            if (MODE == kBall) x = clip_neg(x);
            x = shift + x;
            gy[i] = (x < T(0)) ? T(0) : x;
        }
        __syncthreads();
    }
}

template <typename T, int MODE>
int launch_proj_ragged(T *y, const int32_t *starts, const int32_t *tile_first, int ntiles, const int32_t *large_ids,
                       int nlarge, int max_large, cudaStream_t stream) {
    int dev = 0, num_sm = kNumSM;
    BSLS_CUDA_TRY(cudaGetDevice(&dev));
    BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
    if (ntiles > 0) {
        auto kern = proj_tile_kernel<T, MODE>;
        static thread_local int per_sm = 0;
        if (!per_sm) {
            BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, 0));
            if (per_sm < 1) per_sm = 1;
        }
        const int grid = ntiles < num_sm * per_sm ? ntiles : num_sm * per_sm;
        kern<<<grid, kTileThreads, 0, stream>>>(y, starts, tile_first, ntiles);
        BSLS_LAUNCH_CHECK();
    }
    if (nlarge > 0) {
        auto kern = proj_large_kernel<T, MODE>;
        int KP = 1;
        while (KP < max_large) KP <<= 1;
        const size_t smem = (size_t)2 * KP * sizeof(T);
        static thread_local bool attr_set = false;
        if (!attr_set) {
            BSLS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kLargeMaxBlock * sizeof(T))));
            attr_set = true;
        }
        const int grid = nlarge < 2 * num_sm ? nlarge : 2 * num_sm;
        kern<<<grid, kLargeThreads, smem, stream>>>(y, starts, large_ids, nlarge);
        BSLS_LAUNCH_CHECK();
    }
    return BSLS_OK;
}

}  // namespace bsls
