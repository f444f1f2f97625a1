// pava.cuh -- segmented isotonic regression (pool adjacent violators) on sm_100a.
//
// Replaces isotonic_regression / isotonic_regression_multi, the variant the reference's
// hot path binds (python/c_extensions/isotonic_regression.h:13-58,85-92; called from
// python/main.py:64 and through python/c_extensions/c_extensions.pyx:63-89).
//
// The reference sweeps each block until nothing pools.  In one sweep it walks the pool heads,
// groups them into maximal runs in which each head is <= its predecessor, and replaces a run
// whose first and last value differ by the size-weighted mean, summed left to right (:23-44).
// Decisions inside a sweep only read values the sweep has not touched yet, so the runs of a
// sweep are independent, and only runs with followers ever change anything.  The kernels keep
// the pool heads and the run-start decisions of a block as BIT MASKS and replay exactly those
// runs (pava_block_runs below; pava_words.cuh spreads the masks over lanes for long blocks).
// The arrays y[] and weight[] are the reference's own representation (value / pool size
// stored at the head, stale entries elsewhere), so the result -- values, pool sizes and even
// the stale interior entries -- is bit-identical.
//
//   uniform layouts, K <= 64      pava_small_kernel        one thread per row of 1..16 blocks
//   ragged layouts, blocks <= 32  pava_tile_rows_kernel    rows found from a block-start bitmap
//   blocks of 33 .. 1024          pava_words_kernel        one lane per 32-entry word, packs per warp
//   blocks up to 8192             pava_words_cta_kernel    one CTA per block
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace bsls {

constexpr int kPavaLargeMaxBlock = 8192; // longest block (one CTA, pava_words.cuh)

struct PavaFlags {
    int update;      // copy the head value over its pool at the end (reference `update`)
    int clip01;      // clamp to [0,1] afterwards (python/main.py:65)
    int has_weight;  // weight array given (in/out); otherwise all ones in, result dropped
};


template <typename T> __device__ __forceinline__ T clip01(T v) {
    v = (v > T(0)) ? v : T(0);  // np.maximum(0., x)
    v = (v < T(1)) ? v : T(1);  // np.minimum(1., x)
    return v;
}

// ---------------------------------------------------------------------------------------------
// one THREAD per block: the reference's sweeps replayed run by run, driven by bit masks
// ---------------------------------------------------------------------------------------------
// Small non-negative integer -> floating point without the conversion unit: I2F.F64 issues at a
// quarter of the fp64 add rate on sm_100a; 2^52 + w assembled from bits, minus 2^52, is exact
// for 0 <= w < 2^31.
// The routines of this section also compile for the host (tools/pava_host_check.cu runs this very
// source against the oracle on the CPU); intrinsics get plain-C stand-ins there.
#define BSLS_HD __host__ __device__ __forceinline__
BSLS_HD double bits_to_double(unsigned hi, unsigned lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    const unsigned long long b = ((unsigned long long)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
BSLS_HD unsigned double_hi(double d) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(d);
#else
    unsigned long long b;
    memcpy(&b, &d, 8);
    return (unsigned)(b >> 32);
#endif
}
BSLS_HD double small_int_to(double, int w) { return bits_to_double(0x43300000u, (unsigned)w) - 4503599627370496.0; }
BSLS_HD float small_int_to(float, int w) { return (float)w; }

#ifdef __CUDA_ARCH__
BSLS_HD int bit_lo(uint32_t m) { return __ffs((int)m) - 1; }
BSLS_HD int bit_lo(uint64_t m) { return __ffsll((long long)m) - 1; }
BSLS_HD int bit_hi(uint32_t m) { return 31 - __clz((int)m); }
BSLS_HD int bit_hi(uint64_t m) { return 63 - __clzll((long long)m); }
#else
BSLS_HD int bit_lo(uint32_t m) { return __builtin_ctz(m); }
BSLS_HD int bit_lo(uint64_t m) { return __builtin_ctzll(m); }
BSLS_HD int bit_hi(uint32_t m) { return 31 - __builtin_clz(m); }
BSLS_HD int bit_hi(uint64_t m) { return 63 - __builtin_clzll(m); }
#endif

// num / den for a small positive integer den, correctly rounded (== the reference's
// `numerator / denominator`, isotonic_regression.h:40) without the generic division sequence:
// with y = RN(1/den) from a table, q0 = RN(num*y), two residual corrections
// q <- fma(fma(-den, q, num), y, q) give RN(num/den) (Markstein: a faithful q and a correctly
// rounded reciprocal make the corrected quotient correctly rounded; the residuals are exact
// while nothing under- or overflows).  Zero, subnormal-range, huge and non-finite numerators
// -- where a residual could underflow or the sign of zero would be lost -- take the true
// division.  tools/divtest.c compares 10^9 quotients (random and near-midpoint) bit for bit.
BSLS_HD double div_small(double num, int den, const double *rcp, int rcp_n) {
    const double d = small_int_to(double(0), den);
    const unsigned hi = double_hi(num) & 0x7fffffffu;
    // 2^-900 <= |num| < 2^1001  (biased exponent 123 .. 2023)
    if (hi - 0x07b00000u < 0x76d00000u && (unsigned)den < (unsigned)rcp_n) {
        const double y = rcp[den];
        double q = num * y;
        double r = fma(-d, q, num);
        q = fma(r, y, q);
        r = fma(-d, q, num);
        return fma(r, y, q);
    }
    return num / d;
}
BSLS_HD float div_small(float num, int den, const float *, int) { return num / (float)den; }

// y / w point at one block of K <= 8*sizeof(M) entries in shared memory.
//
// State: `alive` = bit k set when entry k is a pool head; S = bit k set when head k STARTS a
// run of the current sweep, i.e. !(y[k] <= y[previous head]) on the values the sweep began
// with (isotonic_regression.h:23-28 reads only values the sweep has not rewritten yet).  Heads
// that start a run and are followed by another run start are singleton runs: the reference
// leaves them alone, and so they cost nothing here.  Every run with followers is summed
// exactly as the reference does (:33-39: 0 + y*w products left to right, integer weights) and,
// if its first and last value differ, replaced by the quotient at its head (:40-42); followers
// keep their stale value / weight as in the reference.  A merge changes the run-start bit of
// only two heads (the merged head against its predecessor, the successor against the merged
// head): those are re-evaluated into Snext, which becomes S for the next sweep.  The loop ends
// after a sweep that merged nothing (:46).  Lanes do not wait for each other between sweeps.
//
// WMEM = false: cold start (all weights 1 on entry): pool sizes are the gaps between head bits
//               and w[] is only written (when `wout`), never read.
// WMEM = true:  warm start: weights are read from w[] as the reference does.
// Returns the final head mask.
// bst: positions that always start a run -- bit 0, and the first entry of every further block when
// one thread takes several short blocks as one row (runs then never cross a block boundary).
// KC > 0: K == KC is known at compile time (the run-start scan is unrolled).
// STRIDE: distance between consecutive entries of the block in y (1: a row; THREADS: a column of a
// [entry][thread] array, which every lane addresses in its own bank whatever entry it reads).
template <typename T, typename W, typename M, bool WMEM, int KC = 0, int STRIDE = 1>
BSLS_HD M pava_block_runs(T *y, W *w, int K, M alive, M bst, bool wout, const T *rcp, int rcp_n) {
    const M one = 1;
    M S;
    if (KC > 0) K = KC;
    if (!WMEM) {
        T prev = y[(0) * STRIDE];
#ifdef __CUDA_ARCH__
        if (sizeof(M) == 4) {
            // compare bits shifted in from the top (one funnel shift per entry), aligned afterwards
            uint32_t acc = 0x80000000u;  // entry 0
            if (KC > 0) {
#pragma unroll
                for (int r = 1; r < (KC > 0 ? KC : 1); ++r) {
                    const T v = y[(r) * STRIDE];
                    acc = __funnelshift_r(acc, (uint32_t)(!(v <= prev)), 1);
                    prev = v;
                }
            } else {
                for (int r = 1; r < K; ++r) {
                    const T v = y[(r) * STRIDE];
                    acc = __funnelshift_r(acc, (uint32_t)(!(v <= prev)), 1);
                    prev = v;
                }
            }
            S = (M)(acc >> (32 - K)) | bst;
        } else
#endif
        {
            S = bst;
            for (int r = 1; r < K; ++r) {
                const T v = y[(r) * STRIDE];
                S |= (M)(!(v <= prev)) << r;
                prev = v;
            }
        }
    } else {
        M rem = alive;
        int k = bit_lo(rem);
        rem &= rem - 1;
        S = (one << k) | (bst & alive);
        T prev = y[(k) * STRIDE];
        while (rem) {
            k = bit_lo(rem);
            rem &= rem - 1;
            const T v = y[(k) * STRIDE];
            if (!(v <= prev)) S |= one << k;
            prev = v;
        }
    }
    M Snext = S;
    M NS = alive & ~S;  // followers not yet consumed by this sweep
    bool any = false;
    while (NS) {
        const M fb = NS & (~NS + 1);              // lowest follower
        const M low = alive & (fb - 1);           // heads below it: the highest is its run's head
        const int p = bit_hi(low);
        const M pb = one << p;
        const M above = (~one) << p;              // positions above p
        const M Sab = S & alive & above;
        const M eb = Sab & (~Sab + 1);            // next run start (0: the run reaches the end of the block)
        const M fol = alive & (eb - 1) & above;   // the run's followers
        const int e = eb ? bit_lo(eb) : K;
        const M lowp = low ^ pb;                  // heads below p
        // the two neighbours whose run-start bit a merge re-evaluates; fetched early, next to `first`
        const T first = y[(p) * STRIDE];
        const T yprev = y[(lowp ? bit_hi(lowp) : p) * STRIDE];
        const T ynext = y[(eb ? e : p) * STRIDE];
        NS &= ~fol;
        // first follower (always there), then the rare longer tail
        M rem = fol;
        int k = bit_lo(rem);
        rem &= rem - 1;
        int wp = WMEM ? (int)w[p] : k - p;
        int den = wp;
        // the reference starts from 0.0 (isotonic_regression.h:33): 0 + y*w differs from y*w only for a product of
        // -0.0, and a run whose sum could keep that sign (all members -0.0) has first == last and does not pool
        T num = first * small_int_to(T(0), wp);  // -fmad=false: product and sum round separately
        T vprev = y[(k) * STRIDE];
        int kprev = k;
        while (rem) {
            k = bit_lo(rem);
            rem &= rem - 1;
            wp = WMEM ? (int)w[kprev] : k - kprev;
            num += vprev * small_int_to(T(0), wp);
            den += wp;
            kprev = k;
            vprev = y[(k) * STRIDE];
        }
        wp = WMEM ? (int)w[kprev] : e - kprev;
        num += vprev * small_int_to(T(0), wp);
        den += wp;
        if (first != vprev) {
            const T val = div_small(num, den, rcp, rcp_n);
            y[(p) * STRIDE] = val;
            if (WMEM || wout) w[p] = (W)den;
            alive &= ~fol;
            any = true;
            const bool sp = (pb & bst) || !(val <= yprev);
            const bool sq = (eb & bst) || !(ynext <= val);
            Snext = sp ? (Snext | pb) : (Snext & ~pb);
            Snext = sq ? (Snext | eb) : (Snext & ~eb);
        }
        if (NS == 0 && any) {  // this sweep is through and pooled something: the next one starts from Snext
            S = Snext;
            NS = alive & ~S;
            any = false;
        }
    }
    return alive;
}

// heads of a warm-started block: the entries reached by i += weight[i] (isotonic_regression.h:22)
template <typename W, typename M> BSLS_HD M pava_heads_from_weights(const W *w, int K) {
    M alive = 0;
    for (int i = 0; i < K; i += ((int)w[i] > 1 ? (int)w[i] : 1)) alive |= (M)1 << i;
    return alive;
}

// copy each head value over its pool (isotonic_regression.h:50-57), pools given by the head mask
template <typename T, typename M> BSLS_HD void pava_spread(T *y, int K, M alive) {
    M rem = alive;
    while (rem) {
        const int k = bit_lo(rem);
        rem &= rem - 1;
        const int stop = rem ? bit_lo(rem) : K;
        const T v = y[k];
        for (int r = k + 1; r < stop; ++r) y[r] = v;
    }
}

constexpr int kPavaSmallMaxBlock = 64;  // longest block of the thread-per-block kernel

// Uniform layouts with K <= kPavaSmallMaxBlock: a CTA stages THREADS*bpt blocks in shared memory
// (rows padded to an odd pitch so that lanes walking their own rows do not bank-conflict),
// every thread regresses its own block(s) and leaves the head mask, then the tile is written
// back coalesced -- every element fetches the value of the head its mask names.
// A ROW is what one thread regresses: G consecutive blocks of K entries (KR = G*K <= bits of M),
// G > 1 for short blocks so that a lane has enough work to stay in step with its warp.
// A tile is THREADS rows, staged in shared memory with rows padded to an odd pitch (lanes walking
// their own rows then do not bank-conflict).  Tiles are fetched with per-element cp.async
// (LDGSTS: global -> padded shared address, no registers).  CTAs are small (64 threads) and a
// tile is single-buffered: measured on B200, many resident CTAs in different phases hide the
// load latency better than a two-buffer pipeline with half the warps (tools/pava_cfg_sweep.py,
// K = 16: 0.45 ms against 0.50 ms for 10^8 values).
// FAST = the hot configuration (cold start, no weight array, update = 1; main.py:64): the
// loops carry no run-time flags.  CLIP is the [0,1] clamp of python/main.py:65.
// KRC > 0: the row length KR == KRC is a compile-time divisor of THREADS (K = 16, or 4 blocks of 4, ...): every
// thread keeps its column and strides over rows, so the fetch and store loops need no index division, and the
// run-start scan is unrolled.
template <typename T, int THREADS, typename M, bool FAST, bool CLIP, int KRC>
__global__ void __launch_bounds__(THREADS)  // a register cap for 20 CTAs per SM (45 registers) measured 3-6 % slower than 58 registers / 16 CTAs
pava_small_kernel(T *__restrict__ yg, int32_t *__restrict__ wg, long long first, int nb, int K, int G, FastDiv rdiv, PavaFlags fl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(KRC == 0 || (THREADS % KRC == 0 && FAST), "compile-time rows: a length dividing the CTA");
    const int KR = KRC ? KRC : K * G;  // row length (elements)
    const int KS = KR | 1;             // row pitch
    const int pad = KS - KR;           // 1 for even KR: element e of the tile sits at e + row
    constexpr int RCPN = kPavaSmallMaxBlock + 1;
    T *rcp = reinterpret_cast<T *>(smem_raw);
    M *masks = reinterpret_cast<M *>(smem_raw + ((RCPN * sizeof(T) + 15) & ~size_t(15)));
    T *ys = reinterpret_cast<T *>(masks + THREADS);
    uint8_t *wsm = reinterpret_cast<uint8_t *>(ys + (((size_t)THREADS * KS + 1) & ~size_t(1)));  // only with a weight array
    const int tid = threadIdx.x;
    const bool has_weight = !FAST && fl.has_weight;
    const bool update = FAST || fl.update;
    const bool clip = FAST ? CLIP : (fl.clip01 != 0);
    for (int i = tid + 1; i < RCPN; i += THREADS) rcp[i] = T(1) / (T)i;
    M bst = 0;  // first entry of every block of a row
    for (int g = 0; g < G; ++g) bst |= (M)1 << (g * K);
    const int nrows = (nb + G - 1) / G;
    const int ntiles = (nrows + THREADS - 1) / THREADS;
    const long long ntot = (long long)nb * K;
    const int tile_elems = THREADS * KR;
    // two entries per thread and 16-byte stores when rows hold an even number of entries and the span is aligned
    const bool pairs = FAST && sizeof(T) == 8 && pad == 1 && ((reinterpret_cast<uintptr_t>(yg + first) & 15) == 0);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long e0 = (long long)tile * tile_elems;
        const int nel = (int)min((long long)tile_elems, ntot - e0);
        T *gy = yg + first + e0;
        int32_t *gw = wg ? wg + first + e0 : nullptr;
        if (KRC > 0) {
            // element tid + j*THREADS sits in row tid/KRC + j*(THREADS/KRC), column tid % KRC
            constexpr int KRX = KRC > 0 ? KRC : 1;
            constexpr int KSX = KRX | 1;
            constexpr int RPI = THREADS / KRX;
            const T *src = gy + tid;
            T *dst = ys + (tid / KRX) * KSX + (tid % KRX);
            if (nel == tile_elems) {
#pragma unroll
                for (int j = 0; j < KRX; ++j) cp_async_elem<sizeof(T)>(dst + j * RPI * KSX, src + j * THREADS);
            } else {
                for (int j = 0; j < KRX; ++j)
                    if (tid + j * THREADS < nel) cp_async_elem<sizeof(T)>(dst + j * RPI * KSX, src + j * THREADS);
            }
            cp_async_commit();
        } else {
            // thread-relative pointer + constant offsets: no 64-bit address arithmetic per element
            const T *src = gy + tid;
            int i = tid;
            for (; i + 3 * THREADS < nel; i += 4 * THREADS, src += 4 * THREADS) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e = (uint32_t)(i + u * THREADS);
                    cp_async_elem<sizeof(T)>(&ys[e + (pad ? fdiv(e, rdiv) : 0u)], src + u * THREADS);
                }
            }
            for (; i < nel; i += THREADS, src += THREADS) {
                const uint32_t e = (uint32_t)i;
                cp_async_elem<sizeof(T)>(&ys[e + (pad ? fdiv(e, rdiv) : 0u)], src);
            }
            cp_async_commit();
        }
        if (has_weight) {
            for (int e2 = tid; e2 < nel; e2 += THREADS) {
                const uint32_t e = (uint32_t)e2;
                wsm[e + (pad ? fdiv(e, rdiv) : 0u)] = (uint8_t)gw[e2];
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        {
            const int len = min(KR, nel - tid * KR);  // the very last row may hold fewer blocks
            if (len > 0) {
                T *yb = ys + (size_t)tid * KS;
                const M full = len == (int)(8 * sizeof(M)) ? ~M(0) : (((M)1 << len) - 1);
                if (has_weight) {
                    uint8_t *wb = wsm + (size_t)tid * KS;
                    masks[tid] = pava_block_runs<T, uint8_t, M, true>(yb, wb, len, pava_heads_from_weights<uint8_t, M>(wb, len), (M)(bst & full), true, rcp, RCPN);
                } else if (KRC > 0 && len == KRC) {
                    masks[tid] = pava_block_runs<T, uint8_t, M, false, KRC>(yb, nullptr, len, full, bst, false, rcp, RCPN);
                } else {
                    masks[tid] = pava_block_runs<T, uint8_t, M, false>(yb, nullptr, len, full, (M)(bst & full), false, rcp, RCPN);
                }
            }
        }
        __syncthreads();
        // coalesced store; with `update` every element takes the value of the head its mask names
        if (KRC > 0 && (KRC % 2 != 0 || !pairs)) {
            // element tid + j*THREADS: row tid/KRC + j*(THREADS/KRC), column tid % KRC (odd row lengths, unaligned spans)
            constexpr int KRX = KRC > 0 ? KRC : 1;
            constexpr int KSX = KRX | 1;
            constexpr int RPI = THREADS / KRX;
            const int c = tid % KRX, rr0 = tid / KRX;
            const M below = (((M)2) << c) - 1;
            T *dst = gy + tid;
            const M *mrow = masks + rr0;
            const T *yrow = ys + rr0 * KSX;
            if (nel == tile_elems) {
#pragma unroll
                for (int j = 0; j < KRX; ++j) {
                    const uint32_t h = (uint32_t)bit_hi((M)(mrow[j * RPI] & below));
                    T v = yrow[j * RPI * KSX + h];
                    if (clip) v = clip01(v);
                    dst[j * THREADS] = v;
                }
            } else {
                for (int j = 0; j < KRX; ++j) {
                    if (tid + j * THREADS >= nel) break;
                    const uint32_t h = (uint32_t)bit_hi((M)(mrow[j * RPI] & below));
                    T v = yrow[j * RPI * KSX + h];
                    if (clip) v = clip01(v);
                    dst[j * THREADS] = v;
                }
            }
        } else if (KRC > 0) {
            // pair 2*tid + j*2*THREADS: row (2*tid)/KRC + j*(2*THREADS/KRC), columns c2 and c2 + 1 of that row
            constexpr int KRX = KRC > 0 ? KRC : 2;
            constexpr int RP2 = 2 * THREADS / KRX;
            const int c2 = (2 * tid) % KRX, rr0 = (2 * tid) / KRX;
            const M below = (((M)2) << c2) - 1;
            T *dst = gy + 2 * tid;
            const M *mrow = masks + rr0;
            const T *yrow = ys + rr0 * (KRX + 1);
            if (nel == tile_elems) {
#pragma unroll
                for (int j = 0; j < KRX / 2; ++j) {
                    const M m = mrow[j * RP2];
                    const uint32_t h0 = (uint32_t)bit_hi((M)(m & below));
                    const uint32_t h1 = ((m >> (c2 + 1)) & 1) ? (uint32_t)(c2 + 1) : h0;
                    T v0 = yrow[j * RP2 * (KRX + 1) + h0], v1 = yrow[j * RP2 * (KRX + 1) + h1];
                    if (clip) {
                        v0 = clip01(v0);
                        v1 = clip01(v1);
                    }
                    st_stream_v2(reinterpret_cast<double *>(dst + j * 2 * THREADS), (double)v0, (double)v1);
                }
            } else {
                for (int j = 0; j < KRX / 2; ++j) {
                    if (2 * tid + j * 2 * THREADS >= nel) break;
                    const M m = mrow[j * RP2];
                    const uint32_t h0 = (uint32_t)bit_hi((M)(m & below));
                    const uint32_t h1 = ((m >> (c2 + 1)) & 1) ? (uint32_t)(c2 + 1) : h0;
                    T v0 = yrow[j * RP2 * (KRX + 1) + h0], v1 = yrow[j * RP2 * (KRX + 1) + h1];
                    if (clip) {
                        v0 = clip01(v0);
                        v1 = clip01(v1);
                    }
                    st_stream_v2(reinterpret_cast<double *>(dst + j * 2 * THREADS), (double)v0, (double)v1);
                }
            }
        } else if (pairs) {
            T *dst = gy + 2 * tid;
#pragma unroll 2
            for (int e2 = 2 * tid; e2 < nel; e2 += 2 * THREADS, dst += 2 * THREADS) {  // nel is even (KR is)
                const uint32_t e = (uint32_t)e2, r = fdiv(e, rdiv), c = e - r * KR;
                const M m = masks[r];
                const uint32_t h0 = (uint32_t)bit_hi((M)(m & ((((M)2) << c) - 1)));
                const uint32_t h1 = ((m >> (c + 1)) & 1) ? c + 1 : h0;
                T v0 = ys[r * KS + h0], v1 = ys[r * KS + h1];
                if (clip) {
                    v0 = clip01(v0);
                    v1 = clip01(v1);
                }
                st_stream_v2(reinterpret_cast<double *>(dst), (double)v0, (double)v1);
            }
        } else {
#pragma unroll 4
            for (int e2 = tid; e2 < nel; e2 += THREADS) {
                const uint32_t e = (uint32_t)e2, r = fdiv(e, rdiv), c = e - r * KR;
                uint32_t src = c;
                if (update) src = (uint32_t)bit_hi((M)(masks[r] & ((((M)2) << c) - 1)));
                T v = ys[r * KS + src];
                if (clip) v = clip01(v);
                gy[e2] = v;
                if (has_weight) gw[e2] = (int32_t)wsm[r * KS + c];
            }
        }
        __syncthreads();
    }
}

template <typename T, int THREADS, typename M, bool FAST, bool CLIP, int KRC = 0>
int launch_pava_small_cfg(T *y, int32_t *w, long long first, int nb, int K, int G, PavaFlags fl, cudaStream_t stream) {
    auto kern = pava_small_kernel<T, THREADS, M, FAST, CLIP, KRC>;
    const int KR = K * G, KS = KR | 1;
    const size_t buf_elems = ((size_t)THREADS * KS + 1) & ~size_t(1);
    const size_t smem = (((kPavaSmallMaxBlock + 1) * sizeof(T) + 15) & ~size_t(15)) + (size_t)THREADS * sizeof(M) + buf_elems * sizeof(T) +
                        (fl.has_weight ? (size_t)THREADS * KS : 0) + 16;
    static thread_local PerDevice<size_t> cached_smem_pd;
    size_t &cached_smem = cached_smem_pd.get(0);
    static thread_local PerDevice<int> per_sm_pd;
    int &per_sm = per_sm_pd.get(0);
    static thread_local PerDevice<int> num_sm_pd;
    int &num_sm = num_sm_pd.get(0);
    if (cached_smem != smem) {
        BSLS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        BSLS_CUDA_TRY(cudaGetDevice(&dev));
        BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
        BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
        if (per_sm < 1) {
            set_error("pava_small: K=%d does not fit shared memory", K);
            return BSLS_ERR_ARG;
        }
        cached_smem = smem;
    }
    const int nrows = (nb + G - 1) / G;
    const int ntiles = (nrows + THREADS - 1) / THREADS;
    const int grid = ntiles < num_sm * per_sm ? ntiles : num_sm * per_sm;
    kern<<<grid, THREADS, smem, stream>>>(y, w, first, nb, K, G, make_fastdiv((uint32_t)KR), fl);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

template <typename T, int THREADS, typename M>
int launch_pava_small_flags(T *y, int32_t *w, long long first, int nb, int K, int G, PavaFlags fl, cudaStream_t stream) {
    if (!fl.has_weight && fl.update) {  // hot configuration
        if constexpr (sizeof(M) == 4 && (THREADS == 64 || THREADS == 128) && sizeof(T) == 8) {
            const int KR = K * G;
            if (KR == 16 && !getenv("BSLS_PAVA_NO_KRC")) {
                if (fl.clip01) return launch_pava_small_cfg<T, THREADS, M, true, true, 16>(y, w, first, nb, K, G, fl, stream);
                return launch_pava_small_cfg<T, THREADS, M, true, false, 16>(y, w, first, nb, K, G, fl, stream);
            }
            if (KR == 32 && !getenv("BSLS_PAVA_NO_KRC")) {
                if (fl.clip01) return launch_pava_small_cfg<T, THREADS, M, true, true, 32>(y, w, first, nb, K, G, fl, stream);
                return launch_pava_small_cfg<T, THREADS, M, true, false, 32>(y, w, first, nb, K, G, fl, stream);
            }
        }
        if (fl.clip01) return launch_pava_small_cfg<T, THREADS, M, true, true>(y, w, first, nb, K, G, fl, stream);
        return launch_pava_small_cfg<T, THREADS, M, true, false>(y, w, first, nb, K, G, fl, stream);
    }
    return launch_pava_small_cfg<T, THREADS, M, false, false>(y, w, first, nb, K, 1, fl, stream);
}

// row lengths with their own instantiation (CTA size a multiple of the row length): the z-space block sizes of the
// named configurations (K - 1 = 15, 19; also 3 blocks of 5, 5 blocks of 3), besides the 16 / 32 handled with 64
// threads.  (K = 20 with 120 threads measured slower than the generic path: 0.50 ms against 0.44 ms.)
template <typename T, int THREADS, int KRC>
int launch_pava_small_krc(T *y, int32_t *w, long long first, int nb, int K, int G, PavaFlags fl, cudaStream_t stream) {
    if (fl.clip01) return launch_pava_small_cfg<T, THREADS, uint32_t, true, true, KRC>(y, w, first, nb, K, G, fl, stream);
    return launch_pava_small_cfg<T, THREADS, uint32_t, true, false, KRC>(y, w, first, nb, K, G, fl, stream);
}

template <typename T> int launch_pava_small(T *y, int32_t *w, long long first, int nb, int K, PavaFlags fl, cudaStream_t stream) {
    if (nb <= 0) return BSLS_OK;
    if (K <= 32) {
        int G = K <= 8 ? 16 / K : 1;  // short blocks: several to a row (<= 16 entries)
        if constexpr (sizeof(T) == 8) {
            if (!fl.has_weight && fl.update && !getenv("BSLS_PAVA_NO_KRC") && !getenv("BSLS_PAVA_CFG")) {
                const int KR = K * G;
                if (KR == 15) return launch_pava_small_krc<T, 120, 15>(y, w, first, nb, K, G, fl, stream);
                if (KR == 19) return launch_pava_small_krc<T, 114, 19>(y, w, first, nb, K, G, fl, stream);
            }
        }
        int th = 64;
        if (const char *cfg = getenv("BSLS_PAVA_CFG")) {  // tuning experiments: "<threads>,<G>"
            int g = G;
            sscanf(cfg, "%d,%d", &th, &g);
            if (g >= 1 && g * K <= 32) G = g;
        }
        if (th == 128) return launch_pava_small_flags<T, 128, uint32_t>(y, w, first, nb, K, G, fl, stream);
        if (th == 32) return launch_pava_small_flags<T, 32, uint32_t>(y, w, first, nb, K, G, fl, stream);
        return launch_pava_small_flags<T, 64, uint32_t>(y, w, first, nb, K, G, fl, stream);
    }
    return launch_pava_small_flags<T, 64, uint64_t>(y, w, first, nb, K, 1, fl, stream);
}

// ---------------------------------------------------------------------------------------------
// ragged layouts: tiles of whole blocks (the projection's tile grid, plan.cu)
// ---------------------------------------------------------------------------------------------
constexpr int kPavaTileElems = 2048;     // == kPlanTileElems
constexpr int kPavaTileMaxBlock = 512;   // == kPlanTileMaxBlock
constexpr int kPavaTileThreads = 128;
constexpr int kPavaThreadMax = 32;       // longest block one thread takes in a tile (== kPlanMidMin)

// The tile's short blocks are regressed as ROWS -- runs of
// consecutive blocks that one thread takes as a single 32-bit mask problem with forced run starts at
// the block boundaries (pava_block_runs, `bst`), so that a lane has 16..31 entries of work whatever
// the block sizes are.  Rows follow from a bitmap of block starts alone, without lists or scans:
// thread q owns the 16-entry bucket q of the tile; the blocks that START in a bucket are a sequence
// of blocks of <= 16 entries (row A, span <= 31) followed by at most one longer block (a block of
// more than 16 entries that starts in the bucket ends beyond it), which is row B if it has <= 32
// entries and belongs to pava_words_kernel / pava_words_cta_kernel otherwise.
// WMEM: a weight array is given (in / out): heads follow the chain i += weight[i] inside every block of a row, weights
// are read and written as the reference does (isotonic_regression.h:22,37-42).
template <typename T, bool CLIP, bool WMEM>
__global__ void __launch_bounds__(kPavaTileThreads)
pava_tile_rows_kernel(T *__restrict__ yg, int32_t *__restrict__ wg, const int32_t *__restrict__ starts, const int32_t *__restrict__ tile_first,
                      int ntiles, int update) {
    static_assert(kPavaTileThreads * 16 == kPavaTileElems, "one 16-entry bucket per thread");
    // the window holds the rows: blocks that start inside the tile and have at most kPavaThreadMax entries; the body of
    // a longer last block beyond it is neither staged nor written back
    constexpr int WIN = kPavaTileElems + kPavaThreadMax;
    constexpr int NW = WIN / 32 + 4;  // bitmap words (+ look-ahead)
    __shared__ __align__(16) T ybuf[WIN];
    __shared__ uint32_t sb[NW];   // bit i: a block starts at entry i of the window (and one bit at the window's end)
    __shared__ uint32_t cov[NW];  // bit i: entry i belongs to a row (is written back from here)
    __shared__ uint2 rows[2 * kPavaTileThreads];  // {position | length << 16, block-start mask}
    __shared__ int nrows;
    __shared__ uint16_t wbuf[WMEM ? WIN : 2];
    __shared__ T rcp[kPavaThreadMax + 1];
    const int tid = threadIdx.x;
    for (int i = tid + 1; i <= kPavaThreadMax; i += kPavaTileThreads) rcp[i] = T(1) / (T)i;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int fb = tile_first[tile];
        const int nblk = tile_first[tile + 1] - fb;
        if (nblk <= 0) continue;
        const int tile_lo = starts[fb];
        const int nel = min(starts[fb + nblk] - tile_lo, WIN);        // to the end of the last block that starts in the tile
        for (int i = tid; i < NW; i += kPavaTileThreads) {
            sb[i] = 0;
            cov[i] = 0;
        }
        if (tid == 0) nrows = 0;
        {
            const T *src = yg + (size_t)tile_lo + tid;
            for (int i = tid; i < nel; i += kPavaTileThreads, src += kPavaTileThreads) cp_async_elem<sizeof(T)>(&ybuf[i], src);
            cp_async_commit();
            if (WMEM)
                for (int i = tid; i < nel; i += kPavaTileThreads) wbuf[i] = (uint16_t)min(max(wg[(size_t)tile_lo + i], 0), 65535);
        }
        __syncthreads();
        for (int i = tid; i <= nblk; i += kPavaTileThreads) {
            const int s = starts[fb + i] - tile_lo;
            if (s < NW * 32) atomicOr(&sb[s >> 5], 1u << (s & 31));  // a long last block ends beyond the window: no bit needed
        }
        cp_async_wait<0>();
        __syncthreads();
        {
            // 64 entries of the bitmap from the bucket's first position
            const int word = tid >> 1, sh = (tid & 1) * 16;
            const unsigned long long lo64 = ((unsigned long long)sb[word + 1] << 32) | sb[word];
            const unsigned long long W = sh ? ((lo64 >> 16) | ((unsigned long long)sb[word + 2] << 48)) : lo64;
            uint32_t field = (uint32_t)W & 0xffffu;  // blocks that start in this bucket
            if (field) {
                const int s0 = __ffs((int)field) - 1;
                // row A: leading blocks of <= 16 entries
                int e = s0;
                uint32_t bst = 0;
                int longer = -1, longer_len = 0;
                while (field) {
                    const int s = __ffs((int)field) - 1;
                    field &= field - 1;
                    const unsigned long long up = W >> (s + 1);
                    const int len = up ? __ffsll((long long)up) : 64;  // distance to the next block start (64: none in sight)
                    if (len <= 16) {
                        bst |= 1u << (s - s0);
                        e = s + len;
                    } else {
                        longer = s;
                        longer_len = len;
                        break;
                    }
                }
                // queue the rows: the threads then share them evenly (a bucket has 0, 1 or 2 rows)
                const int p0 = tid * 16;
                if (e > s0) {
                    const int slot = atomicAdd(&nrows, 1);
                    rows[slot] = make_uint2((uint32_t)(p0 + s0) | ((uint32_t)(e - s0) << 16), bst);
                }
                if (longer >= 0 && longer_len <= kPavaThreadMax) {
                    const int slot = atomicAdd(&nrows, 1);
                    rows[slot] = make_uint2((uint32_t)(p0 + longer) | ((uint32_t)longer_len << 16), 1u);
                }
            }
        }
        __syncthreads();
        for (int r = tid; r < nrows; r += kPavaTileThreads) {
            const uint2 d = rows[r];
            const int pos = (int)(d.x & 0xffffu), len = (int)(d.x >> 16);
            const uint32_t full = len == 32 ? ~0u : ((1u << len) - 1u);
            T *yb = ybuf + pos;
            uint32_t heads;
            if (WMEM) {
                uint16_t *wb = wbuf + pos;
                uint32_t alive = 0;  // the chain of every block of the row
                uint32_t bs = d.y;
                while (bs) {
                    const int s0 = __ffs((int)bs) - 1;
                    bs &= bs - 1;
                    const int e0 = bs ? __ffs((int)bs) - 1 : len;
                    for (int i = s0; i < e0; i += max(1, (int)wb[i])) alive |= 1u << i;
                }
                heads = pava_block_runs<T, uint16_t, uint32_t, true>(yb, wb, len, alive, d.y, true, rcp, kPavaThreadMax + 1);
            } else {
                heads = pava_block_runs<T, uint16_t, uint32_t, false>(yb, nullptr, len, full, d.y, false, rcp, kPavaThreadMax + 1);
            }
            if (update) pava_spread(yb, len, heads);
            const unsigned long long span = (unsigned long long)full << (pos & 31);
            atomicOr(&cov[pos >> 5], (uint32_t)span);
            if (span >> 32) atomicOr(&cov[(pos >> 5) + 1], (uint32_t)(span >> 32));
        }
        __syncthreads();
        {
            T *dst = yg + (size_t)tile_lo + tid;
            for (int i = tid; i < nel; i += kPavaTileThreads, dst += kPavaTileThreads) {
                if (!((cov[i >> 5] >> (i & 31)) & 1u)) continue;
                T v = ybuf[i];
                if (CLIP) v = clip01(v);
                *dst = v;
                if (WMEM) wg[(size_t)tile_lo + i] = (int32_t)wbuf[i];
            }
        }
        __syncthreads();
    }
}

template <typename T, bool CLIP, bool WMEM>
int launch_pava_tile_rows_cfg(T *y, int32_t *w, const int32_t *starts, const int32_t *tile_first, int ntiles, int update, int cap_per_sm,
                              cudaStream_t stream) {
    auto k = pava_tile_rows_kernel<T, CLIP, WMEM>;
    static thread_local PerDevice<int> grid_full_pd;
    int &grid_full = grid_full_pd.get(0);
    if (!grid_full) {
        int dev = 0, num_sm = num_sms(), per_sm = 1;
        BSLS_CUDA_TRY(cudaGetDevice(&dev));
        BSLS_CUDA_TRY(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev));
        BSLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kPavaTileThreads, 0));
        grid_full = num_sm * (per_sm < 1 ? 1 : per_sm);
    }
    int grid = ntiles < grid_full ? ntiles : grid_full;
    if (cap_per_sm > 0 && grid > cap_per_sm * num_sms()) grid = cap_per_sm * num_sms();
    k<<<grid, kPavaTileThreads, 0, stream>>>(y, w, starts, tile_first, ntiles, update);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

template <typename T>
int launch_pava_tile_rows(T *y, int32_t *w, const int32_t *starts, const int32_t *tile_first, int ntiles, int update, int clip, int cap_per_sm,
                          cudaStream_t stream) {
    if (ntiles <= 0) return BSLS_OK;
    if (w) {
        if (clip) return launch_pava_tile_rows_cfg<T, true, true>(y, w, starts, tile_first, ntiles, update, cap_per_sm, stream);
        return launch_pava_tile_rows_cfg<T, false, true>(y, w, starts, tile_first, ntiles, update, cap_per_sm, stream);
    }
    if (clip) return launch_pava_tile_rows_cfg<T, true, false>(y, w, starts, tile_first, ntiles, update, cap_per_sm, stream);
    return launch_pava_tile_rows_cfg<T, false, false>(y, w, starts, tile_first, ntiles, update, cap_per_sm, stream);
}

}  // namespace bsls
