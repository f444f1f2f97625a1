// simplex_core.cuh -- per-block arithmetic of the simplex projection, in registers.
//
// Replaces python/c_extensions/proj_simplex.h:17-34 of the reference.  The reference
// sorts a copy of the block in descending order, walks it left to right keeping a running
// sum, and keeps the LAST shift (1 - running_k)/k whose element stays positive.  The
// result only depends on the sorted multiset and on the left-to-right order of that sum,
// so we are free to choose how to sort; here a block lives in the registers of G
// cooperating lanes (E values each, -inf padded) and is sorted by
//   * Batcher's merge-exchange network inside a lane  (any E, all comparators same way),
//   * bitonic merges across lanes with warp shuffles   (G a power of two).
// The running sum is then formed strictly left to right (lane after lane), the
// candidates are evaluated with the reference's own expression, and the last passing one
// wins -- which makes the shift, and therefore every projected value, bit-identical to
// the reference for any input.
#pragma once
#include "common.cuh"

namespace bsls {

template <typename T> __device__ __forceinline__ T clip_neg(T v) { return (v < T(0)) ? T(0) : v; }

// compare-exchange, larger value first.  fp64: one DSETP + four 32-bit SELs (sm_100a has no
// native fp64 min/max: fmax() expands to DSETP + SEL + FSEL + NaN fix-up + moves, which is
// slower).  fp32: FMNMX pairs.  Inputs are never NaN.
__device__ __forceinline__ void cmpx_desc(double &hi, double &lo) {
    const double a = hi, b = lo;
    const bool sw = a < b;
    hi = sw ? b : a;
    lo = sw ? a : b;
}
__device__ __forceinline__ void cmpx_desc(float &hi, float &lo) {
    const float a = hi, b = lo;
    hi = fmaxf(a, b);
    lo = fminf(a, b);
}
// keep the larger (take_min == false) or the smaller (take_min == true) of a and b
__device__ __forceinline__ double pick(double a, double b, bool take_min) { return ((a < b) != take_min) ? b : a; }
__device__ __forceinline__ float pick(float a, float b, bool take_min) { return take_min ? fminf(a, b) : fmaxf(a, b); }

#define CE(i, j) cmpx_desc(v[i], v[j]);
#include "sortnet_gen.cuh"
#undef CE

// Merge-exchange sort (Knuth 5.2.2, Algorithm M; comparator lists spelled out by
// gen_sortnet.py so that v[] provably stays in registers).
template <typename T, int N> __device__ __forceinline__ void sort_desc_regs(T (&v)[N]) { SortNet<T, N>::run(v); }

// In-lane bitonic merge: v[] is bitonic on entry (E a power of two), descending on exit.
template <typename T, int E> __device__ __forceinline__ void bitonic_merge_regs(T (&v)[E]) {
#pragma unroll
    for (int d = E / 2; d >= 1; d >>= 1) {
#pragma unroll
        for (int i = 0; i < E; ++i)
            if ((i & d) == 0) cmpx_desc(v[i], v[i + d]);
    }
}

// Sort E*G values held by G consecutive lanes (lane-major: lane 0 of the group ends up
// with the E largest).  G = 1 needs no shuffles.  All lanes of the warp must call this.
template <typename T, int E, int G> __device__ __forceinline__ void sort_desc_group(T (&v)[E], int lane) {
    sort_desc_regs<T, E>(v);
    if constexpr (G > 1) {
        static_assert((E & (E - 1)) == 0 && (G & (G - 1)) == 0 && G <= 32, "cross-lane merge needs powers of two");
#pragma unroll
        for (int L = 1; L < G; L <<= 1) {
            // two sorted runs of L lanes each -> one of 2L lanes.  First stage mirrors the
            // second run (element i against element 2LE-1-i) so that every comparator
            // keeps the same direction and -inf padding stays at the tail.
            {
                const bool upper = (lane & L) != 0;
#pragma unroll
                for (int e = 0; e < E / 2; ++e) {
                    const T a = v[e], b = v[E - 1 - e];
                    const T pa = __shfl_xor_sync(0xffffffffu, b, 2 * L - 1);  // partner's element E-1-e
                    const T pb = __shfl_xor_sync(0xffffffffu, a, 2 * L - 1);  // partner's element e
                    v[e] = pick(a, pa, upper);  // lower lane keeps the max, upper the min
                    v[E - 1 - e] = pick(b, pb, upper);
                }
            }
#pragma unroll
            for (int D = L / 2; D >= 1; D >>= 1) {  // half-cleaners between lanes
                const bool upper = (lane & D) != 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const T a = v[e];
                    const T b = __shfl_xor_sync(0xffffffffu, a, D);
                    v[e] = pick(a, b, upper);
                }
            }
            bitonic_merge_regs<T, E>(v);
        }
    }
}

// ---- the candidate test without a division ---------------------------------------------
// The reference accepts position k (0-based, d = k + 1) when  u_k + fl(w / d) > 0  with
// w = fl(1 - running_k)  (proj_simplex.h:29-31).  A sum of two doubles is positive iff it is
// positive in exact arithmetic, and fl() is monotone, so with X = w + u_k * d (exact):
//     X <  0                      =>  w/d <  -u_k        =>  fl(w/d) <= -u_k  =>  rejected
//     X >= d * ulp_above(-u_k)    =>  w/d >= succ(-u_k)  =>  fl(w/d) >  -u_k  =>  accepted
// X is evaluated with ONE fused multiply-add (its sign is exact) and the acceptance margin
// is over-estimated by |u_k| * d * 2^-51 (+ a tiny floor that covers subnormal u_k).  Only
// when X falls in the sliver between the two tests is the reference's own expression
// evaluated, so the decision is always the reference's, at ~5 instructions instead of an
// IEEE division per element.
template <typename T> struct Margin;
template <> struct Margin<double> {
    __device__ __forceinline__ static double scale() { return 4.440892098500626e-16; }   // 2^-51
    __device__ __forceinline__ static double floor_() { return 1.0e-290; }
};
template <> struct Margin<float> {
    __device__ __forceinline__ static float scale() { return 2.384185791015625e-07f; }   // 2^-22
    __device__ __forceinline__ static float floor_() { return 1.0e-30f; }
};

// classification of one candidate: +1 accepted, -1 rejected, 0 undecided (near-tie)
template <typename T> __device__ __forceinline__ int candidate_class(T u, T w, T d) {
    const T x = fma(u, d, w);
    const T m = fma(fabs(u), d * Margin<T>::scale(), Margin<T>::floor_());
    return (x > m) ? 1 : ((x < T(0)) ? -1 : 0);
}

// Shift (the reference's `lambda`) of one block whose K values sit, sorted descending, in
// v[] of G lanes (position = sub*E + e; positions >= K hold -inf and are ignored).
// Returns the shift to every lane of the group.
template <typename T, int E, int G>
__device__ __forceinline__ T simplex_shift_sorted(const T (&v)[E], int K, int lane) {
    if constexpr (G == 1) {
        // one lane owns the whole block: the reference loop (proj_simplex.h:24-32); the
        // winning candidate is remembered as (numerator, position) and divided once.
        T run = v[0];
        T num = T(1) - run;
        int last = 0;
        bool undecided = false;
#pragma unroll
        for (int e = 1; e < E; ++e) {
            if (e < K) {
                run += v[e];
                const T w = T(1) - run;
                const int c = candidate_class<T>(v[e], w, T(e) + T(1));
                undecided |= (c == 0);
                if (c > 0) {
                    num = w;
                    last = e;
                }
            }
        }
        if (undecided) {  // some candidate was a near-tie: redo the block with the reference's division
            run = v[0];
            num = T(1) - run;
            last = 0;
#pragma unroll
            for (int e = 1; e < E; ++e) {
                if (e < K) {
                    run += v[e];
                    const T w = T(1) - run;
                    if (v[e] + w / (T(e) + T(1)) > T(0)) {
                        num = w;
                        last = e;
                    }
                }
            }
        }
        return num / (T(last) + T(1));  // /1 reproduces `lambda = 1 - sum` of position 0 exactly
    } else {
        const int sub = lane & (G - 1);
        T run = T(0);
        T pre[E];  // pre[e] = running sum up to and including this lane's element e
#pragma unroll
        for (int r = 0; r < G; ++r) {
            if (sub == r) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int pos = r * E + e;
                    run = (pos == 0) ? v[0] : ((pos < K) ? run + v[e] : run);
                    pre[e] = run;
                }
            }
            run = __shfl_sync(0xffffffffu, run, (lane & ~(G - 1)) + r);
        }
        // position 0 is accepted unconditionally (proj_simplex.h:25), later ones by the test
        T num = T(0);
        int last = -1;
        bool undecided = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int pos = sub * E + e;
            if (pos < K) {
                const T w = T(1) - pre[e];
                const int c = (pos == 0) ? 1 : candidate_class<T>(v[e], w, T(pos) + T(1));
                undecided |= (c == 0);
                if (c > 0) {
                    num = w;
                    last = pos;
                }
            }
        }
        if (undecided) {
            num = T(0);
            last = -1;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = sub * E + e;
                if (pos < K) {
                    const T w = T(1) - pre[e];
                    if (pos == 0 || v[e] + w / (T(pos) + T(1)) > T(0)) {
                        num = w;
                        last = pos;
                    }
                }
            }
        }
#pragma unroll
        for (int D = 1; D < G; D <<= 1) {
            const int o_last = __shfl_xor_sync(0xffffffffu, last, D);
            const T o_num = __shfl_xor_sync(0xffffffffu, num, D);
            if (o_last > last) {
                last = o_last;
                num = o_num;
            }
        }
        return num / (T(last) + T(1));
    }
}

}  // namespace bsls
