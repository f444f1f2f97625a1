// pava_seq.cuh -- the reference's three isotonic-regression routines run VERBATIM in their own order of operations,
// one thread per block, on global memory.
//
//   variant 1  isotonic_regression    python/c_extensions/isotonic_regression.h:13-58   (weighted pool-skipping sweeps)
//   variant 2  isotonic_regression_2  python/c_extensions/isotonic_regression.h:61-82   (weight-free sweeps, every entry rewritten)
//   variant 3  isotonic_regression_3  python/c_extensions/isotonic_regression.h:105-155 (one pass with back-tracking, w[k-1] tail markers)
//
// Why these exist next to the parallel kernels of pava.cuh / pava_words.cuh:
//   * variants 2 and 3 form their pool means in a different order (plain sums; nested two-pool means with ">=" ties) and
//     variant 3 leaves a different weight array (tail back-pointers, tie-fused pools).  Serving them with the variant-1
//     kernel gave values equal only to ~1e-15 and variant 1's weights; these kernels return the reference's bits,
//     weights included.
//   * variant 1 on blocks longer than the 8192-entry shared-memory window of pava_words_cta_kernel: the same sequential
//     routine is the fall-back ("served, not fast": one thread per such block).
// The hot path (z-space projection of the solvers) never comes here: it uses variant 1 through the parallel kernels.
#pragma once
#include "common.cuh"

namespace bsls {

template <typename T> __device__ __forceinline__ T seq_clip01(T v) { return v < T(0) ? T(0) : (v > T(1) ? T(1) : v); }

// isotonic_regression.h:13-58, statement for statement
template <typename T> __device__ void pava_seq_v1(T *y, int start, int end, int32_t *weight, int update) {
    for (;;) {
        int i = start, pooled = 0;
        while (i < end) {
            int k = i + weight[i];
            int j = i;
            while (k < end && y[k] <= y[j]) {
                j = k;
                k += weight[k];
            }
            if (y[i] != y[j]) {
                T numerator = T(0);
                int denominator = 0;
                j = i;
                while (j < k) {
                    numerator += y[j] * T(weight[j]);
                    denominator += weight[j];
                    j += weight[j];
                }
                y[i] = numerator / T(denominator);
                weight[i] = denominator;
                pooled = 1;
            }
            i = k;
        }
        if (!pooled) break;
    }
    if (update) {
        int i = start;
        while (i < end) {
            const int k = i + weight[i];
            for (int j = i + 1; j < k; ++j) y[j] = y[i];
            i += weight[i];
        }
    }
}

// isotonic_regression.h:61-82
template <typename T> __device__ void pava_seq_v2(T *y, int start, int end) {
    end -= 1;
    for (;;) {
        int i = start, pooled = 0;
        while (i < end) {
            int k = i;
            while (k < end && y[k] >= y[k + 1]) k += 1;
            if (y[i] != y[k]) {
                T numerator = T(0);
                for (int j = i; j < k + 1; ++j) numerator += y[j];
                const T average = numerator / T(k + 1 - i);
                for (int j = i; j < k + 1; ++j) y[j] = average;
                pooled = 1;
            }
            i = k + 1;
        }
        if (!pooled) break;
    }
}

// isotonic_regression.h:105-155
template <typename T> __device__ void pava_seq_v3(T *y, int start, int end, int32_t *w, int update) {
    int i = start;
    while (i < end) {
        int k = i + w[i];
        int j = i;
        while (k < end && y[k] <= y[j]) {
            j = k;
            k += w[k];
        }
        if (y[i] != y[j]) {
            T numerator = T(0);
            int denominator = 0;
            j = i;
            while (j < k) {
                numerator += y[j] * T(w[j]);
                denominator += w[j];
                j += w[j];
            }
            y[i] = numerator / T(denominator);
            w[i] = denominator;
            w[k - 1] = denominator;
            if (i > start) {
                // back-tracking step
                j = i - w[i - 1];
                while (j >= start && y[j] >= y[i]) {
                    y[j] = (T(w[i]) * y[i] + T(w[j]) * y[j]) / T(w[i] + w[j]);
                    w[j] = w[i] + w[j];
                    i = j;
                    if (j == start) break;
                    j -= w[j - 1];
                }
                w[k - 1] = w[i];
            }
        } else {
            i = k;
        }
    }
    if (update) {
        i = start;
        while (i < end) {
            const int k = i + w[i];
            for (int j = i + 1; j < k; ++j) y[j] = y[i];
            i += w[i];
        }
    }
}

// One thread per block.  ids == nullptr: blocks 0..count-1; else the listed blocks.  `cold` != 0: the weight array is
// scratch and is set to ones first (weight=None of the Python layer, c_extensions.pyx:70-71).
template <typename T, int VARIANT>
__global__ void __launch_bounds__(128) pava_seq_kernel(T *__restrict__ y, int32_t *__restrict__ w, const int32_t *__restrict__ starts,
                                                        const int32_t *__restrict__ ids, int count, int min_size, int update, int cold,
                                                        int clip) {
    for (long long t = blockIdx.x * 128ll + threadIdx.x; t < count; t += 128ll * gridDim.x) {
        const int b = ids ? ids[t] : (int)t;
        const int s = starts[b], e = starts[b + 1];
        if (e - s <= min_size) continue;
        if (VARIANT != 2 && cold)
            for (int i = s; i < e; ++i) w[i] = 1;
        if (VARIANT == 1) pava_seq_v1<T>(y, s, e, w, update);
        if (VARIANT == 2) pava_seq_v2<T>(y, s, e);
        if (VARIANT == 3) pava_seq_v3<T>(y, s, e, w, update);
        if (clip)
            for (int i = s; i < e; ++i) y[i] = seq_clip01(y[i]);
    }
}

template <typename T>
int launch_pava_seq(int variant, T *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int min_size, int update, int cold,
                    int clip, cudaStream_t stream) {
    if (count <= 0) return BSLS_OK;
    long long want = ((long long)count + 127) / 128;
    const int grid = (int)(want < 65535 ? want : 65535);
    if (variant == 1)
        pava_seq_kernel<T, 1><<<grid, 128, 0, stream>>>(y, w, starts, ids, count, min_size, update, cold, clip);
    else if (variant == 2)
        pava_seq_kernel<T, 2><<<grid, 128, 0, stream>>>(y, w, starts, ids, count, min_size, update, cold, clip);
    else
        pava_seq_kernel<T, 3><<<grid, 128, 0, stream>>>(y, w, starts, ids, count, min_size, update, cold, clip);
    BSLS_LAUNCH_CHECK();
    return BSLS_OK;
}

}  // namespace bsls
