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

template <typename T> __device__ __forceinline__ void cmpx_desc(T &hi, T &lo) {
    const T a = hi, b = lo;
    const bool sw = a < b;
    hi = sw ? b : a;
    lo = sw ? a : b;
}

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
                T w[E];
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = __shfl_xor_sync(0xffffffffu, v[E - 1 - e], 2 * L - 1);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const T a = v[e], b = w[e];
                    const bool a_lt_b = a < b;
                    v[e] = (a_lt_b != upper) ? b : a;  // lower lane keeps the max, upper the min
                }
            }
#pragma unroll
            for (int D = L / 2; D >= 1; D >>= 1) {  // half-cleaners between lanes
                const bool upper = (lane & D) != 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const T a = v[e];
                    const T b = __shfl_xor_sync(0xffffffffu, a, D);
                    const bool a_lt_b = a < b;
                    v[e] = (a_lt_b != upper) ? b : a;
                }
            }
            bitonic_merge_regs<T, E>(v);
        }
    }
}

// Shift (the reference's `lambda`) of one block whose K values sit, sorted descending, in
// v[] of G lanes (position = sub*E + e; positions >= K hold -inf and are ignored).
// Returns the shift to every lane of the group.
template <typename T, int E, int G>
__device__ __forceinline__ T simplex_shift_sorted(const T (&v)[E], int K, int lane) {
    if constexpr (G == 1) {
        // one lane owns the whole block: the reference loop, verbatim (proj_simplex.h:24-32)
        T run = v[0];
        T shift = T(1) - run;
#pragma unroll
        for (int e = 1; e < E; ++e) {
            if (e < K) {
                run += v[e];
                const T cand = (T(1) - run) / (T(e) + T(1));
                if (v[e] + cand > T(0)) shift = cand;
            }
        }
        return shift;
    }
    const int sub = lane & (G - 1);
    T run = T(0);
    T pre[E];  // pre[e] = running sum up to and including this lane's element e
    {
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
    }
    // candidates; position 0 is accepted unconditionally (proj_simplex.h:25), later ones
    // when sorted[k] + (1 - running_k)/(k + 1) > 0 (proj_simplex.h:29-31).
    T shift = T(0);
    int last = -1;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = sub * E + e;
        if (pos < K) {
            const T cand = (pos == 0) ? (T(1) - pre[e]) : (T(1) - pre[e]) / (T(pos) + T(1));
            if (pos == 0 || v[e] + cand > T(0)) {
                shift = cand;
                last = pos;
            }
        }
    }
    if constexpr (G > 1) {
#pragma unroll
        for (int D = 1; D < G; D <<= 1) {
            const int o_last = __shfl_xor_sync(0xffffffffu, last, D);
            const T o_shift = __shfl_xor_sync(0xffffffffu, shift, D);
            if (o_last > last) {
                last = o_last;
                shift = o_shift;
            }
        }
    }
    return shift;
}

}  // namespace bsls
