// pava_words.cuh -- isotonic regression of blocks of 33 .. 1024 entries: one LANE per 32-entry word.
//
// Same replay of the reference's sweeps as pava_block_runs (pava.cuh; isotonic_regression.h:13-58),
// with the head / run-start bit masks of a block spread over the lanes of a warp: lane j of a
// block owns entries 32j .. 32j+31, their `alive` word (pool heads) and their `S` word (heads that
// start a run of the current sweep).  A warp regresses a PACK of blocks that together fill at most
// 32 words, so short and long blocks use the same kernel and no lane idles by construction.
//
// One sweep:
//   A  every lane publishes its words (frozen copies A0 / St for run finding, A for the kills).
//   B  every lane merges the runs whose HEAD lies in its word.  A run may continue into later
//      words: followers there are found through the published words, summed in the reference's
//      order (0 + y*w products left to right, weights = gaps between head positions, :33-39) and,
//      when the run pools (first != last, :29), removed with atomicAnd on the owning word.
//      Runs of one sweep are independent (the reference reads only values the sweep has not
//      rewritten, :23-28), so all lanes work at once.  A merge marks two heads DIRTY: its own head
//      and the head that follows the run.
//   C  every lane re-evaluates the run-start bit of its dirty heads against the previous head
//      (possibly in an earlier word) on the new values -- the state the next sweep starts from.
// until a sweep merges nothing anywhere in the warp (a block that is finished has no run left
// that pools, so idling through the others' sweeps does not change it).
//
// WMEM = false: cold start (all weights 1, no weight array: the configuration of main.py:64): pool sizes are the gaps
//               between head bits.
// WMEM = true:  a weight array is given (in / out): the heads are the entries reached by i += weight[i]
//               (isotonic_regression.h:22), weights are read from and written to the array as the reference does,
//               so that values, pool sizes and the stale interior entries come out identical.
#pragma once
#include "pava.cuh"

namespace bsls {

constexpr int kWordsWarps = 2;             // warps per CTA
constexpr int kWordsMaxBlock = 1024;       // 32 words
constexpr int kWordsRcp = kPavaSmallMaxBlock + 1;

template <typename T, bool WMEM> struct WordsWarpSmem {
    T y[1024];
    uint32_t A0[32], A[32], St[32], D[32];
    uint16_t w[WMEM ? 1024 : 2];  // pool sizes (block lengths are <= 8192)
};

struct WordsWarpSync {  // the lanes of one warp
    __device__ __forceinline__ static void sync() { __syncwarp(); }
    __device__ __forceinline__ static bool any(bool p) { return __any_sync(0xffffffffu, p); }
};
struct WordsCtaSync {  // all threads of the CTA (one long block per CTA)
    __device__ __forceinline__ static void sync() { __syncthreads(); }
    __device__ __forceinline__ static bool any(bool p) { return __syncthreads_or(p) != 0; }
};

// y: the block, linear; A0 / A / St / D: its word arrays (LW entries each).  lane state: j = word index inside the
// block (j < 0: the lane has no word), K = entries of the block.
template <typename T, typename P, bool WMEM>
__device__ __forceinline__ void pava_words_engine(T *y, uint16_t *w, uint32_t *A0, uint32_t *A, uint32_t *St, uint32_t *D, int lane, int j, int LW,
                                                  int K, const T *rcp) {
    const bool have = j >= 0;
    const int base = have ? 32 * j : 0;
    const int nloc = have ? min(32, K - base) : 0;
    const uint32_t full_bits = nloc >= 32 ? ~0u : ((1u << nloc) - 1u);
    uint32_t alive = full_bits;
    uint32_t S = 0;
    bool unit = true;  // every weight of the block is 1 on entry: the cold-start state
    if (WMEM) {
        // Weights of all ones (the usual call: a fresh array that is to receive the pool sizes) make every entry a
        // head, as on a cold start.  Otherwise the heads are the chain i += weight[i] from the block's first entry
        // (isotonic_regression.h:22), walked by the lane of word 0.
        bool ones = true;
        if (have) {
            int r = lane & 31;
            for (int t = 0; t < 32; ++t) {
                if (r < nloc) ones &= (w[base + r] == 1);
                r = (r + 1) & 31;
            }
            A[j] = 0;
            if (j == 0) D[0] = 0;
        }
        P::sync();
        if (have && !ones) D[0] = 1;
        P::sync();
        unit = have ? (D[0] == 0) : true;
        if (have && !unit && j == 0)
            for (int i = 0; i < K; i += max(1, (int)w[i])) A[i >> 5] |= 1u << (i & 31);
        P::sync();
        if (!unit) alive = have ? A[j] : 0u;
    }
    if (unit) {
        // run-start bits of the first sweep; entries visited in a lane-rotated order so that the lanes of a
        // warp read different banks of the linear layout
        if (have) {
            int r = lane & 31;
            T prev = (base + r > 0) ? y[base + r - 1] : T(0);
#pragma unroll 4
            for (int t = 0; t < 32; ++t) {
                if (r < nloc) {
                    const T v = y[base + r];
                    const bool st = (base + r == 0) || !(v <= prev);
                    S |= (uint32_t)st << r;
                    prev = v;
                }
                r = (r + 1) & 31;
                if (r == 0 && base > 0) prev = y[base - 1];
            }
        }
    } else {
        // run-start bit of every head against the head before it (possibly in an earlier word)
        uint32_t rem = alive;
        while (rem) {
            const uint32_t kb = rem & (~rem + 1);
            rem ^= kb;
            const int k = base + __ffs((int)kb) - 1;
            const uint32_t low = alive & (kb - 1);
            bool st = true;
            if (low) {
                st = !(y[k] <= y[base + 31 - __clz((int)low)]);
            } else {
                for (int jj = j - 1; jj >= 0; --jj) {
                    const uint32_t a = A[jj];
                    if (a) {
                        st = !(y[k] <= y[32 * jj + 31 - __clz((int)a)]);
                        break;
                    }
                }
            }
            if (st) S |= kb;
        }
    }
    if (WMEM) P::sync();
    for (;;) {
        if (have) {
            A0[j] = alive;
            A[j] = alive;
            St[j] = S & alive;
            D[j] = 0;
        }
        P::sync();
        bool merged = false;
        uint32_t dirty = 0;
        if (alive) {
            const uint32_t Sal = S & alive;
            const uint32_t NS = alive & ~S;
            // first follower of every run that has followers in this word: a carry started at a run start
            // ripples through the dead positions above it and stops at the next head if that is a follower
            uint32_t ff = ((~alive | Sal) + Sal) & NS;
            // the run through the end of this word continues if the next non-empty word opens with a follower
            bool tail = false;
            int hs = 0;
            if (Sal) {
                hs = 31 - __clz((int)Sal);                            // highest run start of the word
                if (!(alive & ((~1u) << hs))) {                        // ... with no follower in this word
                    for (int jj = j + 1; jj < LW; ++jj) {
                        const uint32_t a = A0[jj];
                        if (a) {
                            tail = !(St[jj] & (a & (~a + 1)));
                            break;
                        }
                    }
                }
            }
            uint32_t kill = 0;
            while (ff || tail) {
                int p;
                if (ff) {
                    const uint32_t fb = ff & (~ff + 1);
                    ff ^= fb;
                    p = 31 - __clz((int)(alive & (fb - 1)));
                } else {
                    p = hs;
                    tail = false;
                }
                const uint32_t pb = 1u << p;
                const uint32_t above = (~1u) << p;
                const uint32_t Sab = Sal & above;
                const uint32_t eb = Sab & (~Sab + 1);
                const uint32_t fol = alive & (eb - 1) & above;
                const T first = y[base + p];
                T num = T(0), vprev = first;
                int kprev = base + p;
                uint32_t rem = fol;
                int den = 0;
                while (rem) {
                    const int k = base + __ffs((int)rem) - 1;
                    rem &= rem - 1;
                    const int wp = WMEM ? (int)w[kprev] : k - kprev;
                    num += vprev * small_int_to(T(0), wp);  // -fmad=false: product and sum round separately
                    den += wp;
                    kprev = k;
                    vprev = y[k];
                }
                int e = K;
                if (eb) {
                    e = base + __ffs((int)eb) - 1;
                } else {
                    for (int jj = j + 1; jj < LW; ++jj) {
                        const uint32_t a = A0[jj];
                        if (!a) continue;
                        const uint32_t st = St[jj];
                        uint32_t fm = st ? (a & ((st & (~st + 1)) - 1)) : a;
                        while (fm) {
                            const int k = 32 * jj + __ffs((int)fm) - 1;
                            fm &= fm - 1;
                            const int wp = WMEM ? (int)w[kprev] : k - kprev;
                            num += vprev * small_int_to(T(0), wp);
                            den += wp;
                            kprev = k;
                            vprev = y[k];
                        }
                        if (st) {
                            e = 32 * jj + __ffs((int)st) - 1;
                            break;
                        }
                    }
                }
                {
                    const int wp = WMEM ? (int)w[kprev] : e - kprev;
                    num += vprev * small_int_to(T(0), wp);
                    den += wp;
                }
                if (kprev != base + p && first != vprev) {
                    y[base + p] = div_small(num, den, rcp, kWordsRcp);
                    if (WMEM) w[base + p] = (uint16_t)den;
                    kill |= fol;
                    dirty |= pb | eb;
                    merged = true;
                    if (!eb) {
                        for (int jj = j + 1; jj < LW; ++jj) {
                            const uint32_t a = A0[jj];
                            if (!a) continue;
                            const uint32_t st = St[jj];
                            const uint32_t fm = st ? (a & ((st & (~st + 1)) - 1)) : a;
                            if (fm) atomicAnd(&A[jj], ~fm);
                            if (st) {
                                atomicOr(&D[jj], st & (~st + 1));
                                break;
                            }
                        }
                    }
                }
            }
            if (kill) atomicAnd(&A[j], ~kill);
        }
        const bool any = P::any(merged);
        P::sync();
        if (!any) break;
        if (have) {
            alive = A[j];
            dirty = (dirty | D[j]) & alive;
            while (dirty) {
                const uint32_t kb = dirty & (~dirty + 1);
                dirty ^= kb;
                const int k = base + __ffs((int)kb) - 1;
                const uint32_t low = alive & (kb - 1);
                bool st = true;  // no head before it: the block opens here
                if (low) {
                    st = !(y[k] <= y[base + 31 - __clz((int)low)]);
                } else {
                    for (int jj = j - 1; jj >= 0; --jj) {
                        const uint32_t a = A[jj];
                        if (a) {
                            st = !(y[k] <= y[32 * jj + 31 - __clz((int)a)]);
                            break;
                        }
                    }
                }
                S = st ? (S | kb) : (S & ~kb);
            }
        }
        P::sync();
    }
    if (have) A[j] = alive;
    P::sync();
}

// A pack = consecutive entries of `ids` (ragged layouts; pack_first[] from plan_pack_words) or consecutive blocks of a
// uniform layout (starts == nullptr: block b covers [first + b*Kuni, first + (b+1)*Kuni)).
template <typename T, bool CLIP, bool WMEM>
__global__ void __launch_bounds__(kWordsWarps * 32)
pava_words_kernel(T *__restrict__ yg, int32_t *__restrict__ wg, const int32_t *__restrict__ starts, const int32_t *__restrict__ ids,
                  const int32_t *__restrict__ pack_first, int npacks, long long first, int nb, int Kuni, FastDiv kdiv, int update) {
    __shared__ __align__(16) WordsWarpSmem<T, WMEM> smem[kWordsWarps];
    __shared__ T rcp[kWordsRcp];
    for (int i = threadIdx.x + 1; i < kWordsRcp; i += kWordsWarps * 32) rcp[i] = T(1) / (T)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WordsWarpSmem<T, WMEM> &sm = smem[wid];
    const int wpb = starts ? 0 : (Kuni + 31) >> 5;   // uniform: words per block
    const int bpp = starts ? 0 : 32 / wpb;           // uniform: blocks per pack
    for (int pack = blockIdx.x * kWordsWarps + wid; pack < npacks; pack += gridDim.x * kWordsWarps) {
        // lane b < cnt describes block b of the pack: global start, length, first word slot
        int cnt, g0 = 0, Kb = 0, w0 = 0;
        if (starts) {
            const int pf = pack_first[pack];
            cnt = pack_first[pack + 1] - pf;
            if (lane < cnt) {
                const int b = ids[pf + lane];
                g0 = starts[b];
                Kb = starts[b + 1] - g0;
            }
            const int words = (Kb + 31) >> 5;
            int incl = words;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            w0 = incl - words;
        } else {
            cnt = min(bpp, nb - pack * bpp);
            if (lane < cnt) {
                g0 = (int)((long long)(pack * bpp + lane) * Kuni);  // relative to `first`
                Kb = Kuni;
                w0 = lane * wpb;
            }
        }
        // which block / word this lane serves
        const unsigned startmask = __reduce_or_sync(0xffffffffu, (lane < cnt) ? (1u << w0) : 0u);
        const int total_words = __shfl_sync(0xffffffffu, w0 + ((Kb + 31) >> 5), cnt - 1);
        const int myb = __popc(startmask & ((2u << lane) - 1u)) - 1;
        const int my_w0 = __shfl_sync(0xffffffffu, w0, myb < 0 ? 0 : myb);
        const int my_K = __shfl_sync(0xffffffffu, Kb, myb < 0 ? 0 : myb);
        const bool have = lane < total_words;
        const int j = have ? lane - my_w0 : -1;
        const int LW = (my_K + 31) >> 5;
        if (!starts) {
            // uniform layout: the pack is one contiguous span; entry e of it is entry e % K of block e / K
            const int nel = cnt * Kuni;
            const T *src = yg + first + (long long)pack * bpp * Kuni + lane;
            if (Kuni == 32 * wpb) {  // whole words: the shared image is the span itself
                for (int e = lane; e < nel; e += 32, src += 32) cp_async_elem<sizeof(T)>(&sm.y[e], src);
            } else {
                for (int e = lane; e < nel; e += 32, src += 32) {
                    const int b = (int)fdiv((uint32_t)e, kdiv);
                    cp_async_elem<sizeof(T)>(&sm.y[32 * wpb * b + (e - b * Kuni)], src);
                }
            }
            if (WMEM) {
                const int32_t *wsrc = wg + first + (long long)pack * bpp * Kuni;
                for (int e = lane; e < nel; e += 32) {
                    const int b = (int)fdiv((uint32_t)e, kdiv);
                    sm.w[32 * wpb * b + (e - b * Kuni)] = (uint16_t)min(max(wsrc[e], 0), 65535);
                }
            }
        } else {
            // fetch: block by block, coalesced
            for (int b = 0; b < cnt; ++b) {
                const int bg = __shfl_sync(0xffffffffu, g0, b), bK = __shfl_sync(0xffffffffu, Kb, b), bw = __shfl_sync(0xffffffffu, w0, b);
                const T *src = yg + first + bg;
                for (int i = lane; i < bK; i += 32) cp_async_elem<sizeof(T)>(&sm.y[32 * bw + i], src + i);
                if (WMEM) {
                    const int32_t *wsrc = wg + first + bg;
                    for (int i = lane; i < bK; i += 32) sm.w[32 * bw + i] = (uint16_t)min(max(wsrc[i], 0), 65535);
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        {
            const int wb = have ? my_w0 : 0;
            pava_words_engine<T, WordsWarpSync, WMEM>(sm.y + 32 * wb, sm.w + (WMEM ? 32 * wb : 0), sm.A0 + wb, sm.A + wb, sm.St + wb, sm.D + wb, lane, j,
                                                      LW, my_K, rcp);
        }
        // store: every entry takes the value of the head at or below it
        if (!starts) {
            const int nel = cnt * Kuni;
            T *dst = yg + first + (long long)pack * bpp * Kuni + lane;
            for (int e = lane; e < nel; e += 32, dst += 32) {
                const int b = (int)fdiv((uint32_t)e, kdiv);
                const int wbase = wpb * b, i = e - b * Kuni;
                T v;
                if (update) {
                    int w = i >> 5;
                    uint32_t m = sm.A[wbase + w] & ((2u << (i & 31)) - 1u);
                    while (m == 0) m = sm.A[wbase + --w];
                    v = sm.y[32 * (wbase + w) + 31 - __clz((int)m)];
                } else {
                    v = sm.y[32 * wbase + i];
                }
                if (CLIP) v = clip01(v);
                *dst = v;
                if (WMEM) wg[first + (long long)pack * bpp * Kuni + e] = (int32_t)sm.w[32 * wbase + i];
            }
        } else
        for (int b = 0; b < cnt; ++b) {
            const int bg = __shfl_sync(0xffffffffu, g0, b), bK = __shfl_sync(0xffffffffu, Kb, b), bw = __shfl_sync(0xffffffffu, w0, b);
            T *dst = yg + first + bg;
            for (int i = lane; i < bK; i += 32) {
                T v;
                if (update) {
                    int w = i >> 5;
                    uint32_t m = sm.A[bw + w] & ((2u << (i & 31)) - 1u);
                    while (m == 0) m = sm.A[bw + --w];
                    v = sm.y[32 * (bw + w) + 31 - __clz((int)m)];
                } else {
                    v = sm.y[32 * bw + i];
                }
                if (CLIP) v = clip01(v);
                dst[i] = v;
                if (WMEM) wg[first + bg + i] = (int32_t)sm.w[32 * bw + i];
            }
        }
        __syncwarp();
    }
}

// Blocks longer than a warp's pack: one CTA per block, one thread per word (K <= 32 * kWordsCtaThreads).
constexpr int kWordsCtaThreads = 256;
inline size_t pava_words_cta_smem(int K, size_t elem) {
    const size_t words = ((size_t)K + 31) / 32;
    return (((size_t)K * elem + 15) & ~size_t(15)) + 4 * words * sizeof(uint32_t) + kWordsRcp * elem + 2 * (size_t)K + 64;
}
template <typename T, bool CLIP, bool WMEM>
__global__ void __launch_bounds__(kWordsCtaThreads)
pava_words_cta_kernel(T *__restrict__ yg, int32_t *__restrict__ wg, const int32_t *__restrict__ starts, const int32_t *__restrict__ ids, int count,
                      int max_block, int update) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int max_words = (max_block + 31) >> 5;
    T *y = reinterpret_cast<T *>(smem_raw);
    uint32_t *A0 = reinterpret_cast<uint32_t *>(smem_raw + (((size_t)max_block * sizeof(T) + 15) & ~size_t(15)));
    uint32_t *A = A0 + max_words, *St = A + max_words, *D = St + max_words;
    T *rcp = reinterpret_cast<T *>((reinterpret_cast<uintptr_t>(D + max_words) + 15) & ~uintptr_t(15));
    uint16_t *wsm = reinterpret_cast<uint16_t *>(rcp + kWordsRcp);  // max_block entries (with a weight array)
    for (int i = tid + 1; i < kWordsRcp; i += kWordsCtaThreads) rcp[i] = T(1) / (T)i;
    for (int it = blockIdx.x; it < count; it += gridDim.x) {
        const int b = ids ? ids[it] : it;  // ids == nullptr: every block of the layout
        const int g0 = starts[b];
        const int K = starts[b + 1] - g0;
        if (K > max_block) continue;  // beyond the shared-memory window: served by pava_seq_kernel (uniform over the CTA)
        const int LW = (K + 31) >> 5;
        T *gy = yg + g0;
        for (int i = tid; i < K; i += kWordsCtaThreads) cp_async_elem<sizeof(T)>(&y[i], gy + i);
        cp_async_commit();
        if (WMEM)
            for (int i = tid; i < K; i += kWordsCtaThreads) wsm[i] = (uint16_t)min(max(wg[g0 + i], 0), 65535);
        cp_async_wait<0>();
        __syncthreads();
        pava_words_engine<T, WordsCtaSync, WMEM>(y, wsm, A0, A, St, D, tid, tid < LW ? tid : -1, LW, K, rcp);
        for (int i = tid; i < K; i += kWordsCtaThreads) {
            T v;
            if (update) {
                int w = i >> 5;
                uint32_t m = A[w] & ((2u << (i & 31)) - 1u);
                while (m == 0) m = A[--w];
                v = y[32 * w + 31 - __clz((int)m)];
            } else {
                v = y[i];
            }
            if (CLIP) v = clip01(v);
            gy[i] = v;
            if (WMEM) wg[g0 + i] = (int32_t)wsm[i];
        }
        __syncthreads();
    }
}

template <typename T, bool CLIP, bool WMEM>
int launch_pava_words_cta_cfg(T *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int max_block, int update, int cap_per_sm,
                              cudaStream_t stream) {
    auto k = pava_words_cta_kernel<T, CLIP, WMEM>;
    const size_t smem = pava_words_cta_smem(max_block, sizeof(T));
    int dev = 0, num_sm = num_sms(), per_sm = 1;
    BSLS_CUDA_TRY(cudaGetDevice(&dev));
    BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
    BSLS_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pava_words_cta_smem(32 * kWordsCtaThreads, sizeof(T))));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kWordsCtaThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
    k<<<count < per_sm * num_sm ? count : per_sm * num_sm, kWordsCtaThreads, smem, stream>>>(y, w, starts, ids, count, max_block, update);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

template <typename T>
int launch_pava_words_cta(T *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int max_block, int update, int clip, int cap_per_sm,
                          cudaStream_t stream) {
    if (count <= 0) return BSLS_OK;
    if (max_block > 32 * kWordsCtaThreads) {
        set_error("pava_words_cta: block of %d entries exceeds %d", max_block, 32 * kWordsCtaThreads);
        return BSLS_ERR_ARG;
    }
    if (w) {
        if (clip) return launch_pava_words_cta_cfg<T, true, true>(y, w, starts, ids, count, max_block, update, cap_per_sm, stream);
        return launch_pava_words_cta_cfg<T, false, true>(y, w, starts, ids, count, max_block, update, cap_per_sm, stream);
    }
    if (clip) return launch_pava_words_cta_cfg<T, true, false>(y, w, starts, ids, count, max_block, update, cap_per_sm, stream);
    return launch_pava_words_cta_cfg<T, false, false>(y, w, starts, ids, count, max_block, update, cap_per_sm, stream);
}

template <typename T, bool CLIP, bool WMEM>
int launch_pava_words_cfg(T *y, int32_t *w, const int32_t *starts, const int32_t *ids, const int32_t *pack_first, int npacks, long long first, int nb,
                          int Kuni, int update, int cap_per_sm, cudaStream_t stream) {
    auto k = pava_words_kernel<T, CLIP, WMEM>;
    static thread_local PerDevice<int> full_pd;
    int &full = full_pd.get(0);
    if (!full) {
        int dev = 0, num_sm = num_sms(), per_sm = 1;
        BSLS_CUDA_TRY(cudaGetDevice(&dev));
        BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
        BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kWordsWarps * 32, 0));
        full = num_sm * (per_sm < 1 ? 1 : per_sm);
    }
    const int want = (npacks + kWordsWarps - 1) / kWordsWarps;
    int grid = want < full ? want : full;
    if (cap_per_sm > 0 && grid > cap_per_sm * num_sms()) grid = cap_per_sm * num_sms();
    k<<<grid, kWordsWarps * 32, 0, stream>>>(y, w, starts, ids, pack_first, npacks, first, nb, Kuni, make_fastdiv((uint32_t)(Kuni > 0 ? Kuni : 1)), update);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

template <typename T>
int launch_pava_words(T *y, int32_t *w, const int32_t *starts, const int32_t *ids, const int32_t *pack_first, int npacks, long long first, int nb,
                      int Kuni, int update, int clip, int cap_per_sm, cudaStream_t stream) {
    if (npacks <= 0) return BSLS_OK;
    if (w) {
        if (clip) return launch_pava_words_cfg<T, true, true>(y, w, starts, ids, pack_first, npacks, first, nb, Kuni, update, cap_per_sm, stream);
        return launch_pava_words_cfg<T, false, true>(y, w, starts, ids, pack_first, npacks, first, nb, Kuni, update, cap_per_sm, stream);
    }
    if (clip) return launch_pava_words_cfg<T, true, false>(y, w, starts, ids, pack_first, npacks, first, nb, Kuni, update, cap_per_sm, stream);
    return launch_pava_words_cfg<T, false, false>(y, w, starts, ids, pack_first, npacks, first, nb, Kuni, update, cap_per_sm, stream);
}

}  // namespace bsls
