// proj_ragged.cuh -- segmented simplex / l1-ball projection for blocks of mixed sizes.
//
// Replaces proj_multi_simplex / proj_multi_ball (python/c_extensions/proj_simplex.h:37-74)
// for ragged layouts (BASELINE config 3: power-law sizes 2..4096).
//
// Work split (decided once per layout, see plan.cu):
//   * TILE kernel: the element range is cut into fixed tiles of kTileElems; a CTA owns the
//     blocks that START in its tile (at most kTileMaxBlock long, so they fit the CTA's
//     shared-memory window of kTileElems + kTileMaxBlock values).  Inside the CTA the blocks
//     are binned by size class (so that the lanes of a warp do similar work) and then
//       - blocks of at most 8 values: one thread, sorting network in registers;
//       - 9..32 values: one thread, candidate selection (select_core.cuh);
//       - longer blocks, and the few whose support is too dense for one thread: one warp,
//         candidate selection with up to 128 candidates in the warp's registers;
//       - blocks even a warp cannot settle are queued for the LARGE kernel.
//     Results are written into the window in place; the tile goes back with coalesced stores.
//     HBM traffic: read y once, write y once, read the int32 starts once.
//   * LARGE kernel: one CTA per block longer than kTileMaxBlock (up to kLargeMaxBlock) and per
//     queued block: the block is staged in shared memory, warp 0 tries candidate selection,
//     and only a dense support falls back to a bitonic sort with the reference's serial sum.
// All paths reproduce the reference's arithmetic order, so results are bit-identical to it.
#pragma once
#include <cstdio>

#include "proj_uniform.cuh"

namespace bsls {

constexpr int kTileElems = 2048;     // tile grid pitch (elements)
constexpr int kTileMaxBlock = 512;   // longest block the tile kernel handles
constexpr int kTileThreads = 128;
constexpr int kTileThreadMax = 32;   // longest block one thread takes inside a tile (== kPlanMidMin)
constexpr int kLargeMaxBlock = 8192; // longest block the one-CTA kernel handles
constexpr int kLargeThreads = 512;
constexpr int kNumClasses = 8;

__device__ __forceinline__ int size_class(int K) {
    // 0:<=4 1:<=8 2:<=16 3:<=32 4:<=64 5:<=128 6:<=256 7:<=512
    return K <= 4 ? 0 : (30 - __clz(K - 1));  // ceil(log2 K) - 2 for K > 4
}

// y <- max(y + shift, 0) over one block held in shared memory, by one thread
template <typename T, int MODE> __device__ __forceinline__ void apply_shift_thread(T *blk, int K, T shift) {
    for (int j = 0; j < K; ++j) {
        T x = blk[j];
        if (MODE == kBall) x = clip_neg(x);
        x = shift + x;
        blk[j] = (x < T(0)) ? T(0) : x;
    }
}

// l1-ball: clip, and project only when the clipped block sums to more than one; the sum runs in
// index order as in the reference (proj_simplex.h:54-62).  Returns false when the block is final.
template <typename T> __device__ __forceinline__ bool ball_needs_projection(T *blk, int K) {
    T total = T(0);
    for (int k = 0; k < K; ++k) {
        const T x = blk[k];
        if (!(x < T(0))) total += x;
    }
    return total > T(1);
}

// Blocks of at most E values: registers, sorting network, the reference's loop.
template <typename T, int E, int MODE> __device__ __forceinline__ void tiny_block(T *blk, int K) {
    T v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        T x = Num<T>::neg_inf();
        if (e < K) {
            x = blk[e];
            if (MODE == kBall) x = clip_neg(x);
        }
        v[e] = x;
    }
    sort_desc_regs<T, E>(v);
    const T shift = simplex_shift_sorted<T, E, 1>(v, K, 0);
    apply_shift_thread<T, MODE>(blk, K, shift);
}

// tile_first[t] = index of the first block whose start lies in tile t (tile_first[ntiles] = nb)
// slow: queue of block ids for the LARGE kernel, count at slow[nb]
template <typename T, int MODE>
__global__ void __launch_bounds__(kTileThreads, 5)
proj_tile_kernel(T *__restrict__ y, const int32_t *__restrict__ starts /* nb+1, last = n */,
                 const int32_t *__restrict__ tile_first, int ntiles, int32_t *__restrict__ slow, int nb) {
    constexpr int NW = kTileThreads / 32;
    // the window holds the blocks this kernel owns: they start inside the tile and have at most kTileThreadMax values;
    // what lies beyond (the body of a longer last block) is neither staged nor written back
    constexpr int WIN = kTileElems + kTileThreadMax;
    constexpr int NCOV = WIN / 32 + 2;
    __shared__ __align__(16) T ybuf[WIN];
    __shared__ uint32_t cov[NCOV];                    // bit i: entry i was projected here (is written back)
    __shared__ uint16_t sstart[kTileElems + 2];       // block starts relative to the window
    __shared__ uint16_t list[kTileElems];
    __shared__ uint16_t wlist[(kTileElems + kTileMaxBlock) / 17 + 8];  // single-thread failures (17..32 values only)
    __shared__ __align__(16) T tcand[kSelMaxCand * kTileThreads];      // thread phase; reused by the warp phase
    static_assert(kSelMaxCand * kTileThreads >= NW * kSelWarpCand, "warp candidate lists alias the thread slots");
    T(*wcand)[kSelWarpCand] = reinterpret_cast<T(*)[kSelWarpCand]>(tcand);
    __shared__ int cnt[kNumClasses], off[kNumClasses + 1], fill[kNumClasses];
    __shared__ int s_nwarp;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int fb = tile_first[tile];
        const int nblk = tile_first[tile + 1] - fb;
        if (nblk <= 0) continue;  // uniform across the CTA
        if (tid < kNumClasses) {
            cnt[tid] = 0;
            fill[tid] = 0;
        }
        if (tid == 0) s_nwarp = 0;
        const int tile_lo = starts[fb];
        // relative starts; a "large" last block (it ends the window) is clamped, it is never touched here
        for (int i = tid; i <= nblk; i += kTileThreads) sstart[i] = (uint16_t)min(starts[fb + i] - tile_lo, 65535);
        __syncthreads();
        const int nel = min((int)sstart[nblk], WIN);
        {
            const T *src = y + (size_t)tile_lo + tid;
            for (int i = tid; i < nel; i += kTileThreads, src += kTileThreads) cp_async_elem<sizeof(T)>(&ybuf[i], src);
            cp_async_commit();
        }
        for (int i = tid; i < NCOV; i += kTileThreads) cov[i] = 0;
        cp_async_wait<0>();
        __syncthreads();
        // ---- bin the blocks by size class (counting sort on shared counters) ---------------
        for (int i = tid; i < nblk; i += kTileThreads) {
            const int K = sstart[i + 1] - sstart[i];
            if (K <= kTileThreadMax) atomicAdd(&cnt[size_class(K)], 1);  // longer: a block of proj_mid_kernel / proj_large_kernel
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
            if (K <= kTileThreadMax) {
                const int c = size_class(K);
                list[off[c] + atomicAdd(&fill[c], 1)] = (uint16_t)i;
            }
        }
        __syncthreads();
        // ---- thread per block: classes 0..3 (at most 32 values) ------------------------------
        const int nthread = off[4];
        for (int p = tid; p < nthread; p += kTileThreads) {
            const int b = list[p];
            const int s = sstart[b];
            const int K = sstart[b + 1] - s;
            T *blk = ybuf + s;
            {  // mark the block for the write-back (a block the LARGE kernel ends up with goes back unchanged)
                const unsigned long long span = (K == 32 ? 0xffffffffull : ((1ull << K) - 1ull)) << (s & 31);
                atomicOr(&cov[s >> 5], (uint32_t)span);
                if (span >> 32) atomicOr(&cov[(s >> 5) + 1], (uint32_t)(span >> 32));
            }
            if (MODE == kBall && !ball_needs_projection<T>(blk, K)) {
                for (int j = 0; j < K; ++j) blk[j] = clip_neg(blk[j]);
                continue;
            }
            if (K <= 4) {
                tiny_block<T, 4, MODE>(blk, K);
            } else if (K <= 8) {
                tiny_block<T, 8, MODE>(blk, K);
            } else {
                T shift;
                if (select_shift_thread<T, MODE == kBall, 0>(blk, K, lane, false, tcand + tid, kTileThreads, shift))
                    apply_shift_thread<T, MODE>(blk, K, shift);
                else
                    wlist[atomicAdd(&s_nwarp, 1)] = (uint16_t)b;
            }
        }
        __syncthreads();
        // ---- warp per block: the few blocks a single thread gave up on (dense support) -------------------
        const int nwork = s_nwarp;
        for (int p = wid; p < nwork; p += NW) {
            const int b = wlist[p];
            const int s = sstart[b];
            const int K = sstart[b + 1] - s;
            T *blk = ybuf + s;
            bool project = true;
            if (MODE == kBall) {
                int flag = 0;
                if (lane == 0) flag = ball_needs_projection<T>(blk, K) ? 1 : 0;
                project = __shfl_sync(0xffffffffu, flag, 0) != 0;
            }
            T shift = T(0);
            bool ok = true;
            if (project) ok = select_shift_warp<T, MODE == kBall>(blk, K, lane, wcand[wid], shift);
            if (ok) {
                for (int j = lane; j < K; j += 32) {
                    T x = blk[j];
                    if (MODE == kBall) x = clip_neg(x);
                    x = shift + x;
                    blk[j] = (x < T(0)) ? T(0) : x;
                }
            } else if (lane == 0) {
                slow[atomicAdd(&slow[nb], 1)] = fb + b;  // dense support: the LARGE kernel sorts it (block left untouched)
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- coalesced write-back (large blocks inside the window are rewritten unchanged;
        //      the LARGE kernel runs afterwards on the same stream) ----------------------------
        {
            T *dst = y + (size_t)tile_lo + tid;
            for (int i = tid; i < nel; i += kTileThreads, dst += kTileThreads)
                if ((cov[i >> 5] >> (i & 31)) & 1u) *dst = ybuf[i];
        }
        __syncthreads();
    }
}

// Blocks of kTileThreadMax < K <= kTileMaxBlock: one WARP per block (ids from the plan), staged in the
// warp's slice of shared memory, candidate selection with up to 128 candidates in registers.  Its own
// launch puts all of them in flight at once (inside a tile there are only a handful).  Dense supports
// are queued for proj_large_kernel.
constexpr int kMidWarps = 4;
template <typename T, int MODE>
__global__ void __launch_bounds__(kMidWarps * 32)
proj_mid_kernel(T *__restrict__ y, const int32_t *__restrict__ starts, const int32_t *__restrict__ ids, int count,
                int32_t *__restrict__ slow, int nb) {
    __shared__ __align__(16) T ys[kMidWarps][kTileMaxBlock];
    __shared__ __align__(16) T wcand[kMidWarps][kSelWarpCand];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int it = blockIdx.x * kMidWarps + wid; it < count; it += gridDim.x * kMidWarps) {
        const int b = ids[it];
        const int lo = starts[b];
        const int K = starts[b + 1] - lo;
        T *gy = y + (size_t)lo;
        T *blk = ys[wid];
        for (int i = lane; i < K; i += 32) blk[i] = gy[i];
        __syncwarp();
        bool project = true;
        if (MODE == kBall) {
            int flag = 0;
            if (lane == 0) flag = ball_needs_projection<T>(blk, K) ? 1 : 0;
            project = __shfl_sync(0xffffffffu, flag, 0) != 0;
        }
        T shift = T(0);
        bool ok = true;
        if (project) ok = select_shift_warp<T, MODE == kBall>(blk, K, lane, wcand[wid], shift);
        if (ok) {
            for (int j = lane; j < K; j += 32) {
                T x = blk[j];
                if (MODE == kBall) x = clip_neg(x);
                x = shift + x;
                gy[j] = (x < T(0)) ? T(0) : x;
            }
        } else if (lane == 0) {
            slow[atomicAdd(&slow[nb], 1)] = b;
        }
        __syncwarp();
    }
}

// ---- one CTA per large block ---------------------------------------------------------------------
// block-wide reductions for kLargeThreads threads (result to every thread)
template <typename T> __device__ __forceinline__ T cta_reduce_max(T v, T *s_red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = s_red[0];
    for (int w = 1; w < kLargeThreads / 32; ++w) r = (s_red[w] > r) ? s_red[w] : r;
    return r;
}
template <typename T> __device__ __forceinline__ T cta_reduce_sum(T v, T *s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = s_red[0];
    for (int w = 1; w < kLargeThreads / 32; ++w) r += s_red[w];
    return r;
}

// Blocks longer than kLargeMaxBlock never fit shared memory.  Only their CANDIDATES are staged: the
// whole CTA scans the block in global memory for its maximum, tightens the threshold bound with
// Michelot rounds (select_core.cuh) and compacts the survivors into srt[]; everything not staged is
// provably inactive, so the reference's algorithm run on the staged values alone yields the same
// shift.  Returns the number of staged values (traps when they exceed `cap`: a support that dense in a
// block that long is outside this revision).
// Scratch in global memory for long blocks whose candidates do not even fit the shared-memory window
// (e.g. 10^6 uniform random values: the rounding margin of the selection keeps ~2 % of them).  One
// such block at a time: `lock` serialises the CTAs that need it.
template <typename T> struct HugeScratch {
    T *buf;        // 2 * cap values
    int cap;       // power of two
    int *lock;     // 0 = free
};

template <typename T, int MODE>
__device__ int stage_candidates(const T *gy, int K, T *&srt, int cap, const HugeScratch<T> &huge, bool &used_scratch, T *s_red,
                                int *s_cnt) {
    const int tid = threadIdx.x;
    T umax = Num<T>::neg_inf();
    for (int i = tid; i < K; i += kLargeThreads) {
        T x = gy[i];
        if (MODE == kBall) x = clip_neg(x);
        umax = (x > umax) ? x : umax;
    }
    umax = cta_reduce_max<T>(umax, s_red);
    const T delta = select_delta<T>(K, umax);
    T tau = (umax - T(1)) - delta;
    int c = K, c_prev = K + 1;
    for (int round = 0; round < 24; ++round) {
        T sl = T(0), cl = T(0);
        for (int i = tid; i < K; i += kLargeThreads) {
            T x = gy[i];
            if (MODE == kBall) x = clip_neg(x);
            if (x >= tau) {
                sl += x;
                cl += T(1);
            }
        }
        const T ssum = cta_reduce_sum<T>(sl, s_red);
        c = (int)cta_reduce_sum<T>(cl, s_red);  // exact: counts stay far below 2^24
        if (c <= 8 || c >= c_prev) break;
        c_prev = c;
        const T t2 = (ssum - T(1)) / (T)c - delta;
        if (!(t2 > tau)) break;
        tau = t2;
    }
    used_scratch = false;
    if (c > cap) {
        if (!huge.buf || c > huge.cap) {
            if (tid == 0)
                printf("libbsls_b200: a block of %d values keeps %d candidates (limit %d): support too dense for the long-block path\n", K, c,
                       huge.buf ? huge.cap : cap);
            __syncthreads();
            asm volatile("trap;");
        }
        if (tid == 0) {
            while (atomicCAS(huge.lock, 0, 1) != 0) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
        srt = huge.buf;
        used_scratch = true;
    }
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    for (int i = tid; i < K; i += kLargeThreads) {
        T x = gy[i];
        if (MODE == kBall) x = clip_neg(x);
        if (x >= tau) srt[atomicAdd(s_cnt, 1)] = x;  // order is irrelevant: the values are sorted next
    }
    __syncthreads();
    return *s_cnt;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kLargeThreads)
proj_large_kernel(T *__restrict__ y, const int32_t *__restrict__ starts, const int32_t *__restrict__ ids, int count_host,
                  const int32_t *__restrict__ count_dev, HugeScratch<T> huge) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_last, s_done, s_cnt;
    __shared__ T s_shift, s_total;
    __shared__ T s_red[kLargeThreads / 32];
    __shared__ __align__(16) T wcand[kSelWarpCand];
    const int tid = threadIdx.x;
    const int count = count_dev ? *count_dev : count_host;
    for (int it = blockIdx.x; it < count; it += gridDim.x) {
        const int b = ids ? ids[it] : it;
        const int lo = starts[b];
        const int Kfull = starts[b + 1] - lo;
        T *gy = y + (size_t)lo;
        const T ninf = Num<T>::neg_inf();
        T *srt = reinterpret_cast<T *>(smem_raw);  // staged values (sorted later)
        // K: number of staged values -- the whole block, or only its candidates when it is too long
        int K = Kfull;
        bool used_scratch = false;
        if (Kfull > kLargeMaxBlock) K = stage_candidates<T, MODE>(gy, Kfull, srt, kLargeMaxBlock, huge, used_scratch, s_red, &s_cnt);
        int KP = 1;
        while (KP < K) KP <<= 1;
        T *pre = srt + KP;                         // KP running sums
        if (Kfull > kLargeMaxBlock) {
            for (int i = K + tid; i < KP; i += kLargeThreads) srt[i] = ninf;
        } else {
            for (int i = tid; i < KP; i += kLargeThreads) {
                T x = ninf;
                if (i < K) {
                    x = gy[i];
                    if (MODE == kBall) x = clip_neg(x);
                }
                srt[i] = x;
            }
        }
        if (tid == 0) s_last = 0;
        if (MODE == kBall) {
            if (tid == 0) {  // the reference's left-to-right sum over the kept entries
                T total = T(0);
                for (int k = 0; k < Kfull; ++k) {
                    const T x = gy[k];
                    if (!(x < T(0))) total += x;
                }
                s_total = total;
            }
        }
        __syncthreads();
        const bool project = (MODE == kBall) ? (s_total > T(1)) : true;
        // candidate selection by warp 0 on the staged copy (srt[0..K) is still in input order)
        if (project) {
            if (tid < 32) {
                T shift;
                const bool ok = select_shift_warp<T, false>(srt, K, tid, wcand, shift);  // srt is already clipped in ball mode
                if (tid == 0) {
                    s_done = ok ? 1 : 0;
                    if (ok) s_shift = shift;
                }
            }
            __syncthreads();
        }
        if (project && !s_done) {
            // dense support: bitonic sort, descending, -inf sentinels at the tail
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
        for (int i = tid; i < Kfull; i += kLargeThreads) {
            T x = gy[i];
            if (MODE == kBall) x = clip_neg(x);
            x = shift + x;
            gy[i] = (x < T(0)) ? T(0) : x;
        }
        __syncthreads();
        if (used_scratch && tid == 0) {
            __threadfence();
            atomicExch(huge.lock, 0);
        }
    }
}

template <typename T, int MODE>
int launch_proj_ragged(T *y, const int32_t *starts, const int32_t *tile_first, int ntiles, const int32_t *mid_ids, int nmid,
                       const int32_t *large_ids, int nlarge, int max_large, int32_t *slow, int nb, const RaggedStreams &rs,
                       void *huge_buf, int huge_cap, int *huge_lock, cudaStream_t stream) {
    const HugeScratch<T> huge = {reinterpret_cast<T *>(huge_buf), huge_cap, huge_lock};
    int dev = 0, num_sm = num_sms();
    BSLS_CUDA_TRY(cudaGetDevice(&dev));
    BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
    auto large = proj_large_kernel<T, MODE>;
    static thread_local PerDevice<bool> attr_set_pd;
    bool &attr_set = attr_set_pd.get(false);
    if (!attr_set) {
        BSLS_CUDA_TRY(cudaFuncSetAttribute(large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kLargeMaxBlock * sizeof(T))));
        attr_set = true;
    }
    const bool tiled = ntiles > 0;
    if (tiled) BSLS_CUDA_TRY(cudaMemsetAsync(slow + nb, 0, sizeof(int32_t), stream));
    BSLS_CUDA_TRY(cudaEventRecord(rs.fork, stream));
    if (tiled && nmid > 0) {
        BSLS_CUDA_TRY(cudaStreamWaitEvent(rs.aux[0], rs.fork, 0));
        auto mid = proj_mid_kernel<T, MODE>;
        static thread_local PerDevice<int> mid_full_pd;
        int &mid_full = mid_full_pd.get(0);
        if (!mid_full) {
            int per = 1;
            BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, mid, kMidWarps * 32, 0));
            mid_full = num_sm * (per < 1 ? 1 : per);
        }
        const int want = (nmid + kMidWarps - 1) / kMidWarps;
        mid<<<want < mid_full ? want : mid_full, kMidWarps * 32, 0, rs.aux[0]>>>(y, starts, mid_ids, nmid, slow, nb);
        BSLS_LAUNCH_CHECK();
        BSLS_CUDA_TRY(cudaEventRecord(rs.join[0], rs.aux[0]));
    }
    if (nlarge > 0) {
        BSLS_CUDA_TRY(cudaStreamWaitEvent(rs.aux[1], rs.fork, 0));
        int KP = 1;
        while (KP < max_large && KP < kLargeMaxBlock) KP <<= 1;   // longer blocks stage only their candidates
        const size_t smem = (size_t)2 * KP * sizeof(T);
        int per = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, large, kLargeThreads, smem) != cudaSuccess || per < 1) per = 1;
        const int grid = nlarge < per * num_sm ? nlarge : per * num_sm;
        large<<<grid, kLargeThreads, smem, rs.aux[1]>>>(y, starts, large_ids, nlarge, nullptr, huge);
        BSLS_LAUNCH_CHECK();
        BSLS_CUDA_TRY(cudaEventRecord(rs.join[1], rs.aux[1]));
    }
    if (tiled) {
        auto kern = proj_tile_kernel<T, MODE>;
        static thread_local PerDevice<int> per_sm_pd;
        int &per_sm = per_sm_pd.get(0);
        if (!per_sm) {
            BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, 0));
            if (per_sm < 1) per_sm = 1;
        }
        const int grid = ntiles < num_sm * per_sm ? ntiles : num_sm * per_sm;
        kern<<<grid, kTileThreads, 0, stream>>>(y, starts, tile_first, ntiles, slow, nb);
        BSLS_LAUNCH_CHECK();
    }
    if (tiled && nmid > 0) BSLS_CUDA_TRY(cudaStreamWaitEvent(stream, rs.join[0], 0));
    if (nlarge > 0) BSLS_CUDA_TRY(cudaStreamWaitEvent(stream, rs.join[1], 0));
    if (tiled) {
        // blocks whose support was too dense for a warp (count known only on the device; usually zero)
        large<<<num_sm, kLargeThreads, (size_t)2 * kTileMaxBlock * sizeof(T), stream>>>(y, starts, slow, 0, slow + nb, HugeScratch<T>{nullptr, 0, nullptr});
        BSLS_LAUNCH_CHECK();
    }
    return BSLS_OK;
}

}  // namespace bsls
