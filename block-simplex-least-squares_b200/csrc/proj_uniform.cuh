// proj_uniform.cuh -- segmented simplex / l1-ball projection, all blocks of one size K.
//
// Replaces proj_multi_simplex / proj_multi_ball (python/c_extensions/proj_simplex.h:37-74)
// for the layouts of BASELINE configs 1, 2, 4, 5 (K = 5, 4/16/64, 20, 16).
//
// HBM-bound streaming kernel, one pass over y (read 1x, write 1x):
//   * persistent CTAs (grid = SMs x resident CTAs), each looping over tiles of TB blocks;
//   * a tile is pulled into shared memory with ONE 1-D bulk copy (TMA, UBLKCP) signalled
//     on an mbarrier; two stages, so the next tile is in flight while this one is sorted;
//   * every block is sorted in the registers of G lanes (simplex_core.cuh); shared-memory
//     reads are 128-bit and rotated by the lane so that equal strides do not bank-conflict;
//   * the output pass re-reads the tile from shared memory and writes y with 128-bit
//     streaming stores, fully coalesced.
#pragma once
#include <cstdlib>

#include "select_core.cuh"

namespace bsls {

enum ProjMode { kSimplex = 0, kBall = 1 };

template <typename T> struct Vec16;  // one 16-byte granule
template <> struct Vec16<double> {
    static constexpr int N = 2;
    using type = double2;
};
template <> struct Vec16<float> {
    static constexpr int N = 4;
    using type = float4;
};


// Loads the K values of one block (at blkp, in shared memory) into the registers of its G
// lanes.  Which lane/register receives which element is irrelevant to the sort, so the
// lanes interleave: granule slot (e2*G + sub) reads granule (e2*G + lane) mod KV.  Inside a
// block that is a rotation (a bijection); across the quarter-warp that serves one 128-bit
// shared-memory wavefront the eight lanes hit eight different 16-byte bank groups even
// when every block starts on the same bank (K*sizeof(T) a multiple of 128 bytes).
template <typename T, int E, int G, int MODE>
__device__ __forceinline__ void load_block_regs(T (&v)[E], const T *blkp, int K, int lane, bool live, bool vec_ok) {
    constexpr int VN = Vec16<T>::N;
    using VT = typename Vec16<T>::type;
    const T ninf = Num<T>::neg_inf();
    const int sub = lane & (G - 1);
    if (vec_ok && (E % VN == 0)) {
        const int KV = K / VN;  // granules per block (vec_ok: K % VN == 0)
        int q = lane % KV;
        const int step = G % KV;
#pragma unroll
        for (int e2 = 0; e2 < E / VN; ++e2) {
            const int slot = e2 * G + sub;
            T tmp[VN];
#pragma unroll
            for (int j = 0; j < VN; ++j) tmp[j] = ninf;
            if (live && slot < KV) {
                const VT g = *reinterpret_cast<const VT *>(blkp + q * VN);
                if constexpr (VN == 2) {
                    tmp[0] = g.x;
                    tmp[1] = g.y;
                } else {
                    tmp[0] = g.x;
                    tmp[1] = g.y;
                    tmp[2] = g.z;
                    tmp[3] = g.w;
                }
                if (MODE == kBall) {
#pragma unroll
                    for (int j = 0; j < VN; ++j) tmp[j] = clip_neg(tmp[j]);
                }
            }
            q += step;
            if (q >= KV) q -= KV;
#pragma unroll
            for (int j = 0; j < VN; ++j) v[e2 * VN + j] = tmp[j];
        }
    } else {
        int q = lane % K;
        const int step = G % K;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int slot = e * G + sub;
            T x = ninf;
            if (live && slot < K) {
                x = blkp[q];
                if (MODE == kBall) x = clip_neg(x);
            }
            q += step;
            if (q >= K) q -= K;
            v[e] = x;
        }
    }
}

// Kernel configuration (all compile time):
//   E, G     registers per lane / lanes per block (E*G >= K)
//   THREADS  CTA size;  BPT blocks per lane-group per tile  =>  tile = THREADS/G*BPT blocks
//   STAGES   2: the next tile is in flight while this one is processed (fewer, fatter CTAs)
//            1: one buffer, latency hidden by the other resident CTAs (more warps per SM)
//   MINB     minimum resident CTAs per SM the register allocator must allow
//   KC       block size when known at compile time (0: runtime K <= E*G, -inf padded)
template <typename T, int E, int G, int THREADS, int MINB, int BPT, int STAGES, int KC, int MODE>
__global__ void __launch_bounds__(THREADS, MINB)
proj_uniform_kernel(T *__restrict__ y, long long first, int nb, int Krt, FastDiv kdiv, int aligned, const T *__restrict__ gsrc,
                    T tstep, T *__restrict__ yout, StepCtl ctl) {
    // ctl.t != null (device-resident solver loop): the step is read from the device and nothing happens once the
    // solver has stopped
    if (ctl.t) {
        if (*ctl.done) return;
        tstep = (T)*ctl.t;
    }
    // gsrc != null: FUSED projected-gradient step: the tile is formed as y + (-tstep) * gsrc on the
    // way in (np.add(x, -t*g, x_new), python/BATCH.py:91: product and sum rounded separately) and the
    // result goes to yout -- x, g are read once, x_new written once, no intermediate vector.
    const bool fused = gsrc != nullptr;
    constexpr int GROUPS = THREADS / G;
    constexpr int TB = GROUPS * BPT;  // blocks per tile
    constexpr int VN = Vec16<T>::N;
    using VT = typename Vec16<T>::type;
    const int K = KC ? KC : Krt;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_elems = TB * K;
    T *stage0 = reinterpret_cast<T *>(smem_raw);
    T *stage1 = stage0 + (STAGES == 2 ? tile_elems : 0);
    T *lam = stage1 + tile_elems;
    uint64_t *bar = reinterpret_cast<uint64_t *>(lam + TB + (TB & 1));  // 8-byte aligned for float too

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int grp = tid / G;
    const int sub = tid & (G - 1);
    const int ntiles = (nb + TB - 1) / TB;
    T *ybase = y + first;
    const bool vec_ok = (K % VN) == 0;  // block starts are 16-byte aligned inside the tile

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        if (STAGES == 2) mbar_init(&bar[1], 1);
        mbar_init_fence();
    }
    __syncthreads();

    auto issue = [&](int t, int s) {  // thread 0: start the bulk copy of a full tile
        if (!fused && aligned && (nb - t * TB) >= TB) {
            const uint32_t bytes = (uint32_t)(tile_elems * sizeof(T));
            mbar_expect_tx(&bar[s], bytes);
            bulk_g2s(s ? stage1 : stage0, ybase + (size_t)t * tile_elems, bytes, &bar[s]);
        }
    };
    auto kdivide = [&](uint32_t e) -> uint32_t { return KC ? e / (uint32_t)(KC ? KC : 1) : fdiv(e, kdiv); };

    int tile = blockIdx.x;
    if (STAGES == 2 && tile < ntiles && tid == 0) issue(tile, 0);

    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = (STAGES == 2) ? (it & 1) : 0;
        if (STAGES == 2) {
            const int nxt = tile + gridDim.x;
            if (nxt < ntiles && tid == 0) issue(nxt, s ^ 1);
        } else if (tid == 0) {
            issue(tile, 0);
        }

        const int nblk = min(TB, nb - tile * TB);
        const int nel = nblk * K;
        T *buf = s ? stage1 : stage0;
        const T *gin = ybase + (size_t)tile * tile_elems;
        T *gy = (fused && yout ? yout + first : ybase) + (size_t)tile * tile_elems;

        if (fused) {
            const T *gg = gsrc + first + (size_t)tile * tile_elems;
            const T nt = -tstep;
            for (int i = tid; i < nel; i += THREADS) buf[i] = gin[i] + nt * gg[i];
            __syncthreads();
        } else if (aligned && nblk == TB) {
            mbar_wait(&bar[s], (STAGES == 2) ? ((it >> 1) & 1) : (it & 1));
        } else {  // ragged last tile or unaligned base: plain coalesced loads
            for (int i = tid; i < nel; i += THREADS) buf[i] = gin[i];
            __syncthreads();
        }

        // ---- shift of every block -------------------------------------------------------
#pragma unroll 1
        for (int j = 0; j < BPT; ++j) {
            const int blk = j * GROUPS + grp;
            const bool live = blk < nblk;
            const T *blkp = buf + blk * K;
            T v[E];
            load_block_regs<T, E, G, MODE>(v, blkp, K, lane, live, vec_ok);
            bool project = true;
            if (MODE == kBall) {
                // clip negatives; project only when the clipped block sums to more than 1.  The reference adds the
                // kept entries left to right (proj_simplex.h:54-62); all terms are >= 0, so any order of summation
                // is within (K-1) eps of the exact sum: the lanes add their registers (padding entries are -inf:
                // skipped), and only a total within 4 K eps of 1 is re-added in the reference's order by one lane.
                T part = T(0);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (v[e] > T(0)) part += v[e];
#pragma unroll
                for (int o = G / 2; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                const T margin = T(4 * K) * (sizeof(T) == 8 ? T(2.220446049250313e-16) : T(1.1920929e-07)) * (part > T(1) ? part : T(1));
                project = part > T(1);
                const bool unsure = !(part > T(1) + margin) && !(part < T(1) - margin);
                if (__any_sync(0xffffffffu, unsure)) {
                    T total = T(0);
                    if (unsure && live && sub == 0)
                        for (int k = 0; k < K; ++k) {
                            const T x = blkp[k];
                            if (!(x < T(0))) total += x;
                        }
                    if (G > 1) total = __shfl_sync(0xffffffffu, total, lane & ~(G - 1));
                    if (unsure) project = total > T(1);
                }
            }
            sort_desc_group<T, E, G>(v, lane);
            T shift = simplex_shift_sorted<T, E, G>(v, K, lane);
            if (MODE == kBall && !project) shift = T(0);
            if (live && sub == 0) lam[blk] = shift;
        }
        __syncthreads();

        // ---- output pass: y <- max(y + shift, 0), coalesced 16-byte stores -------------------
        if (aligned) {
            const int nvec = nel / VN;
            for (int i = tid; i < nvec; i += THREADS) {
                const VT g = reinterpret_cast<const VT *>(buf)[i];
                T x[VN];
                if constexpr (VN == 2) {
                    x[0] = g.x;
                    x[1] = g.y;
                } else {
                    x[0] = g.x;
                    x[1] = g.y;
                    x[2] = g.z;
                    x[3] = g.w;
                }
                const uint32_t e0 = (uint32_t)i * VN;
                const uint32_t b0 = kdivide(e0);
#pragma unroll
                for (int j = 0; j < VN; ++j) {
                    const uint32_t b = vec_ok ? b0 : kdivide(e0 + j);
                    T t = x[j];
                    if (MODE == kBall) t = clip_neg(t);
                    t = lam[b] + t;
                    x[j] = (t < T(0)) ? T(0) : t;
                }
                if constexpr (VN == 2)
                    st_stream_v2(gy + e0, x[0], x[1]);
                else
                    st_stream_v4(gy + e0, x[0], x[1], x[2], x[3]);
            }
            for (int i = nvec * VN + tid; i < nel; i += THREADS) {
                T t = buf[i];
                if (MODE == kBall) t = clip_neg(t);
                t = lam[kdivide((uint32_t)i)] + t;
                gy[i] = (t < T(0)) ? T(0) : t;
            }
        } else {
            for (int i = tid; i < nel; i += THREADS) {
                T t = buf[i];
                if (MODE == kBall) t = clip_neg(t);
                t = lam[kdivide((uint32_t)i)] + t;
                gy[i] = (t < T(0)) ? T(0) : t;
            }
        }
        __syncthreads();  // the stage and lam[] are free again
    }
}

// ---- SELECT kernel: one thread per block, no full sort (select_core.cuh) -----------------------------
// Same streaming skeleton as proj_uniform_kernel (persistent CTAs, one bulk copy per tile, coalesced
// 128-bit output pass); a tile holds THREADS blocks and every thread finds the shift of its own
// block by candidate selection.  Blocks with a dense support (more than kSelMaxCand candidates)
// are left untouched and queued in `slow` (ids from slow[0], count in slow[nb]); proj_list_kernel
// sorts them afterwards.  Keeping the sorter out of this kernel keeps its register count low.
template <typename T, int THREADS, int MINB, int STAGES, int KC, int MODE>
__global__ void __launch_bounds__(THREADS, MINB)
proj_select_kernel(T *__restrict__ y, long long first, int nb, int Krt, FastDiv kdiv, int aligned, int32_t *__restrict__ slow) {
    constexpr int TB = THREADS;  // blocks per tile
    constexpr int VN = Vec16<T>::N;
    using VT = typename Vec16<T>::type;
    const int K = KC ? KC : Krt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_elems = TB * K;
    T *stage0 = reinterpret_cast<T *>(smem_raw);
    T *stage1 = stage0 + (STAGES == 2 ? tile_elems : 0);
    T *lam = stage1 + tile_elems;
    T *cand = lam + TB;                                     // kSelMaxCand x THREADS, slot-major
    uint64_t *bar = reinterpret_cast<uint64_t *>(cand + (size_t)kSelMaxCand * THREADS + ((TB & 1) ? 1 : 0));
    __shared__ int s_nslow2[2], s_base;               // queue counters, one per tile parity (see the reset below)
    __shared__ uint16_t s_slow[THREADS];
    __shared__ uint8_t s_keep[THREADS];                     // 1: block goes to the sorter, write it back unchanged

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int ntiles = (nb + TB - 1) / TB;
    T *ybase = y + first;
    const bool vec_ok = (K % VN) == 0;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        if (STAGES == 2) mbar_init(&bar[1], 1);
        mbar_init_fence();
        s_nslow2[0] = 0;
        s_nslow2[1] = 0;
    }
    __syncthreads();
    auto issue = [&](int t, int s) {
        if (aligned && (nb - t * TB) >= TB) {
            const uint32_t bytes = (uint32_t)(tile_elems * sizeof(T));
            mbar_expect_tx(&bar[s], bytes);
            bulk_g2s(s ? stage1 : stage0, ybase + (size_t)t * tile_elems, bytes, &bar[s]);
        }
    };
    auto kdivide = [&](uint32_t e) -> uint32_t { return KC ? e / (uint32_t)(KC ? KC : 1) : fdiv(e, kdiv); };
    int tile = blockIdx.x;
    if (STAGES == 2 && tile < ntiles && tid == 0) issue(tile, 0);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        int &s_nslow = s_nslow2[it & 1];
        const int s = (STAGES == 2) ? (it & 1) : 0;
        if (STAGES == 2) {
            const int nxt = tile + gridDim.x;
            if (nxt < ntiles && tid == 0) issue(nxt, s ^ 1);
        } else if (tid == 0) {
            issue(tile, 0);
        }
        const int nblk = min(TB, nb - tile * TB);
        const int nel = nblk * K;
        T *buf = s ? stage1 : stage0;
        T *gy = ybase + (size_t)tile * tile_elems;
        if (aligned && nblk == TB) {
            mbar_wait(&bar[s], (STAGES == 2) ? ((it >> 1) & 1) : (it & 1));
        } else {
            for (int i = tid; i < nel; i += THREADS) buf[i] = gy[i];
            __syncthreads();
        }
        // ---- shift of every block: thread per block ---------------------------------------------------
        if (tid < nblk) {
            const T *blkp = buf + tid * K;
            bool project = true;
            if (MODE == kBall) {
                T total = T(0);  // in index order, as the reference sums (proj_simplex.h:54-62)
                for (int k = 0; k < K; ++k) {
                    const T x = blkp[k];
                    if (!(x < T(0))) total += x;
                }
                project = total > T(1);
            }
            T shift = T(0);
            bool ok = true;
            if (project) ok = select_shift_thread<T, MODE == kBall, KC>(blkp, K, lane, vec_ok, cand + tid, THREADS, shift);
            if (!ok) s_slow[atomicAdd(&s_nslow, 1)] = (uint16_t)tid;
            s_keep[tid] = ok ? 0 : 1;
            lam[tid] = shift;
        }
        __syncthreads();
        const int nslow = s_nslow;
        // the NEXT tile's counter is cleared here, between two barriers of this tile: its first increment
        // comes after this tile's closing barrier, and nobody reads it before then (a full tile is awaited
        // on the mbarrier, without a CTA barrier, so clearing "at the top of the loop" would race)
        if (tid == 0) s_nslow2[(it + 1) & 1] = 0;
        if (nslow > 0) {  // hand the dense blocks to the sorter: one global atomic per tile
            if (tid == 0) s_base = atomicAdd(&slow[nb], nslow);
            __syncthreads();
            for (int i = tid; i < nslow; i += THREADS) slow[s_base + i] = tile * TB + (int)s_slow[i];
        }
        // ---- output pass ----------------------------------------------------------------------------------------
        if (aligned) {
            const int nvec = nel / VN;
            for (int i = tid; i < nvec; i += THREADS) {
                const VT g = reinterpret_cast<const VT *>(buf)[i];
                T x[VN];
                if constexpr (VN == 2) {
                    x[0] = g.x;
                    x[1] = g.y;
                } else {
                    x[0] = g.x;
                    x[1] = g.y;
                    x[2] = g.z;
                    x[3] = g.w;
                }
                const uint32_t e0 = (uint32_t)i * VN;
                const uint32_t b0 = kdivide(e0);
#pragma unroll
                for (int j = 0; j < VN; ++j) {
                    const uint32_t b = vec_ok ? b0 : kdivide(e0 + j);
                    T t = x[j];
                    if (MODE == kBall) t = clip_neg(t);
                    t = lam[b] + t;
                    t = (t < T(0)) ? T(0) : t;
                    x[j] = s_keep[b] ? x[j] : t;
                }
                if constexpr (VN == 2)
                    st_stream_v2(gy + e0, x[0], x[1]);
                else
                    st_stream_v4(gy + e0, x[0], x[1], x[2], x[3]);
            }
            for (int i = nvec * VN + tid; i < nel; i += THREADS) {
                const T x = buf[i];
                const uint32_t b = kdivide((uint32_t)i);
                T t = x;
                if (MODE == kBall) t = clip_neg(t);
                t = lam[b] + t;
                t = (t < T(0)) ? T(0) : t;
                gy[i] = s_keep[b] ? x : t;
            }
        } else {
            for (int i = tid; i < nel; i += THREADS) {
                const T x = buf[i];
                const uint32_t b = kdivide((uint32_t)i);
                T t = x;
                if (MODE == kBall) t = clip_neg(t);
                t = lam[b] + t;
                t = (t < T(0)) ? T(0) : t;
                gy[i] = s_keep[b] ? x : t;
            }
        }
        __syncthreads();
    }
}

// The sorter for the blocks proj_select_kernel queued: G lanes per block, values straight from
// global memory into registers, full sort, reference replay, result written in place.
template <typename T, int E, int G, int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS)
proj_list_kernel(T *__restrict__ y, long long first, int K, const int32_t *__restrict__ slow, int nb) {
    const int count = slow[nb];
    if (count == 0) return;
    constexpr int GROUPS = THREADS / G;
    const int tid = threadIdx.x, lane = tid & 31, sub = tid & (G - 1), grp = tid / G;
    const T ninf = Num<T>::neg_inf();
    for (int base = blockIdx.x * GROUPS; base < count; base += gridDim.x * GROUPS) {  // uniform trip count per warp
        const int idx = base + grp;
        const bool live = idx < count;
        T *blk = y + first + (size_t)(live ? slow[idx] : 0) * K;
        T v[E], raw[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int pos = e * G + sub;  // lane-interleaved: consecutive lanes read consecutive values
            T x = ninf;
            if (live && pos < K) x = blk[pos];
            raw[e] = x;
            v[e] = (MODE == kBall && x < T(0)) ? T(0) : x;
        }
        sort_desc_group<T, E, G>(v, lane);
        const T shift = simplex_shift_sorted<T, E, G>(v, K, lane);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int pos = e * G + sub;
            if (live && pos < K) {
                T t = raw[e];
                if (MODE == kBall) t = clip_neg(t);
                t = shift + t;
                blk[pos] = (t < T(0)) ? T(0) : t;
            }
        }
    }
}

template <typename T, int E, int G, int THREADS, int MINB, int STAGES, int KC, int MODE>
int launch_proj_select_cfg(T *y, long long first, int nb, int K, int32_t *slow, cudaStream_t stream) {
    constexpr int TB = THREADS;
    auto kern = proj_select_kernel<T, THREADS, MINB, STAGES, KC, MODE>;
    const size_t smem = (size_t)STAGES * TB * K * sizeof(T) + (size_t)TB * sizeof(T) + (size_t)kSelMaxCand * THREADS * sizeof(T) +
                        sizeof(T) + 2 * sizeof(uint64_t);
    static thread_local PerDevice<int> cached_blocks_per_sm_pd;
    int &cached_blocks_per_sm = cached_blocks_per_sm_pd.get(-1);
    static thread_local PerDevice<size_t> cached_smem_pd;
    size_t &cached_smem = cached_smem_pd.get(0);
    static thread_local PerDevice<int> num_sm_pd;
    int &num_sm = num_sm_pd.get(0);
    if (cached_blocks_per_sm < 0 || cached_smem != smem) {
        BSLS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        BSLS_CUDA_TRY(cudaGetDevice(&dev));
        BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
        int per = 0;
        BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, THREADS, smem));
        if (per < 1) {
            set_error("proj_select: block size K=%d does not fit shared memory (%zu B)", K, smem);
            return BSLS_ERR_ARG;
        }
        cached_blocks_per_sm = per;
        cached_smem = smem;
    }
    const int ntiles = (nb + TB - 1) / TB;
    const int grid = ntiles < num_sm * cached_blocks_per_sm ? ntiles : num_sm * cached_blocks_per_sm;
    const int aligned = ((reinterpret_cast<uintptr_t>(y + first) % 16) == 0) ? 1 : 0;
    BSLS_CUDA_TRY(cudaMemsetAsync(slow + nb, 0, sizeof(int32_t), stream));
    kern<<<grid, THREADS, smem, stream>>>(y, first, nb, K, make_fastdiv((uint32_t)K), aligned, slow);
    BSLS_LAUNCH_CHECK();
    // the sorter reads the count on the device: no host round trip; an empty queue costs one tiny launch
    constexpr int LT = 128;
    const int lgrid = num_sm * 4;
    proj_list_kernel<T, E, G, LT, MODE><<<lgrid, LT, 0, stream>>>(y, first, K, slow, nb);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

// ---- host side: pick a configuration for K and launch ------------------------------------------
template <typename T, int E, int G, int THREADS, int MINB, int BPT, int STAGES, int KC, int MODE>
int launch_proj_uniform_cfg(T *y, long long first, int nb, int K, cudaStream_t stream, const T *gsrc = nullptr, T tstep = T(0),
                            T *yout = nullptr, const StepCtl *ctl = nullptr) {
    constexpr int TB = THREADS / G * BPT;
    auto kern = proj_uniform_kernel<T, E, G, THREADS, MINB, BPT, STAGES, KC, MODE>;
    const size_t smem = (size_t)STAGES * TB * K * sizeof(T) + (size_t)(TB + (TB & 1)) * sizeof(T) + 2 * sizeof(uint64_t);
    static thread_local PerDevice<int> cached_blocks_per_sm_pd;
    int &cached_blocks_per_sm = cached_blocks_per_sm_pd.get(-1);
    static thread_local PerDevice<size_t> cached_smem_pd;
    size_t &cached_smem = cached_smem_pd.get(0);
    static thread_local PerDevice<int> num_sm_pd;
    int &num_sm = num_sm_pd.get(0);
    if (cached_blocks_per_sm < 0 || cached_smem != smem) {
        BSLS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        BSLS_CUDA_TRY(cudaGetDevice(&dev));
        BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
        int per = 0;
        BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, THREADS, smem));
        if (per < 1) {
            set_error("proj_uniform: block size K=%d does not fit shared memory (%zu B)", K, smem);
            return BSLS_ERR_ARG;
        }
        cached_blocks_per_sm = per;
        cached_smem = smem;
    }
    const int ntiles = (nb + TB - 1) / TB;
    const int grid = ntiles < num_sm * cached_blocks_per_sm ? ntiles : num_sm * cached_blocks_per_sm;
    const int aligned = ((reinterpret_cast<uintptr_t>((gsrc && yout ? yout : y) + first) % 16) == 0) ? 1 : 0;
    kern<<<grid, THREADS, smem, stream>>>(y, first, nb, K, make_fastdiv((uint32_t)K), aligned, gsrc, tstep, yout, ctl ? *ctl : StepCtl{nullptr, nullptr});
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

constexpr int kUniformMaxK = 512;  // larger uniform blocks go through the large-block kernel

// BSLS_TUNE=<n> (development aid) selects alternative configurations for the hot sizes.
inline int tune_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BSLS_TUNE");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// gsrc != null: fused step yout = proj(y - tstep * gsrc); only the sorting kernels fuse (the caller
// checks proj_uniform_fuses(K) first)
inline bool proj_uniform_fuses(int K) { return K < 32 || K > 128; }

template <typename T, int MODE>
int launch_proj_uniform(T *y, long long first, int nb, int K, int32_t *slow, cudaStream_t stream, const T *gsrc = nullptr, T tstep = T(0),
                        T *yout = nullptr, const StepCtl *ctl = nullptr) {
    if (nb <= 0) return BSLS_OK;
    const int tv = tune_variant();
    if (gsrc) slow = nullptr;  // fused mode: sorting kernels only
    // sizes the BASELINE configs name, with K known at compile time
#define CFG(E, G, TH, MINB, BPT, ST, KC) return launch_proj_uniform_cfg<T, E, G, TH, MINB, BPT, ST, KC, MODE>(y, first, nb, K, stream, gsrc, tstep, yout, ctl)
#define SEL(E, G, TH, MINB, ST, KC) return launch_proj_select_cfg<T, E, G, TH, MINB, ST, KC, MODE>(y, first, nb, K, slow, stream)
    // candidate selection instead of a full sort (BSLS_TUNE=100 keeps the sorting kernels everywhere)
    if (tv != 100 && slow != nullptr && K >= 32 && K <= 128) {
        if (K == 32) SEL(32, 1, 128, 5, 1, 32);
        if (K == 64) SEL(16, 4, 64, 5, 1, 64);
        if (K == 128) SEL(16, 8, 64, 3, 1, 128);
        if (K <= 64) SEL(16, 4, 64, 4, 1, 0);
        SEL(16, 8, 64, 2, 1, 0);
    }
    if (K == 4) CFG(4, 1, 256, 4, 4, 2, 4);
    if (K == 16) {
        if (tv == 1) CFG(16, 1, 128, 6, 1, 1, 16);
        if (tv == 2) CFG(16, 1, 256, 3, 1, 1, 16);
        if (tv == 3) CFG(16, 1, 128, 6, 1, 2, 16);
        if (tv == 4) CFG(16, 1, 256, 3, 1, 2, 0);
        if (tv == 5) CFG(16, 1, 128, 6, 1, 1, 0);
        CFG(16, 1, 256, 3, 1, 2, 16);
    }
    if (K == 64) {
        if (tv == 1) CFG(16, 4, 128, 4, 1, 1, 64);
        if (tv == 2) CFG(16, 4, 256, 2, 1, 1, 64);
        if (tv == 3) CFG(32, 2, 128, 2, 1, 1, 64);
        if (tv == 4) CFG(16, 4, 256, 2, 1, 2, 0);
        if (tv == 5) CFG(16, 4, 128, 4, 1, 1, 0);
        CFG(16, 4, 256, 2, 1, 2, 64);
    }
    if (K <= 4) CFG(4, 1, 256, 4, 4, 2, 0);
    if (K <= 8) CFG(8, 1, 256, 4, 2, 2, 0);
    if (K <= 12) CFG(12, 1, 256, 3, 1, 2, 0);
    if (K <= 16) CFG(16, 1, 256, 3, 1, 2, 0);
    if (K <= 20) CFG(20, 1, 128, 5, 1, 2, 0);
    if (K <= 24) CFG(24, 1, 128, 4, 1, 2, 0);
    if (K <= 32) CFG(32, 1, 128, 4, 1, 2, 0);
    if (K <= 64) CFG(16, 4, 256, 2, 1, 2, 0);
    if (K <= 128) CFG(16, 8, 256, 2, 1, 2, 0);
    if (K <= 256) CFG(16, 16, 256, 2, 1, 2, 0);
    if (K <= 512) CFG(16, 32, 256, 2, 1, 2, 0);
#undef CFG
#undef SEL
    set_error("launch_proj_uniform: K=%d above %d", K, kUniformMaxK);
    return BSLS_ERR_ARG;
}

}  // namespace bsls
