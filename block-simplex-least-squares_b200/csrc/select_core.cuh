// select_core.cuh -- the shift of one simplex projection WITHOUT sorting the whole block.
//
// The reference (python/c_extensions/proj_simplex.h:17-34) sorts a copy of the block in
// descending order u_0 >= u_1 >= ..., accumulates S_i = u_0 + ... + u_i left to right and keeps
// tmp_i = (1 - S_i)/(i + 1) of the LAST i with u_i + tmp_i > 0.  Write T_i = (S_i - 1)/(i + 1)
// and theta = max_i T_i (the exact threshold).  Two facts of exact arithmetic:
//   (a) every T_i is a lower bound of theta, and so is (sum_C u - 1)/|C| for ANY set C that
//       contains the support {u > theta} (Michelot's bound);  T_0 = u_max - 1.
//   (b) for i beyond the support, T_i - u_i >= (r + 1)/(i + 1) * (T_r - u_i) for every r < i:
//       an element that lies Delta below some bound T_r fails the reference's test with margin
//       Delta / K.
// The reference's floating-point T_i differs from the exact one by at most ~3 eps (1 + K M)
// (M = largest magnitude involved), so an element more than
//       delta = 32 eps K (1 + K (|u_max| + 2))
// below a lower bound of theta is rejected by the reference's own floating-point test, sits
// after every accepted element in the sorted order and therefore influences nothing.  What is
// left -- the CANDIDATES {u >= bound - delta} -- is processed exactly as the reference does:
// descending order, left-to-right sum, the same test, last passing index wins.  The result is
// bit-identical; the work drops from a K log^2 K sort to two or three scans of the block plus
// a selection sort of the handful of candidates.  Blocks whose candidate set stays large
// (dense supports) are reported back to the caller, which runs them through the full-sort
// path (simplex_core.cuh).
#pragma once
#include "simplex_core.cuh"

namespace bsls {

template <typename T> struct Eps;
template <> struct Eps<double> {
    __device__ __forceinline__ static double v() { return 1.1102230246251565e-16; }
};
template <> struct Eps<float> {
    __device__ __forceinline__ static float v() { return 5.9604645e-08f; }
};

template <typename T> __device__ __forceinline__ T select_delta(int K, T umax) {
    const T k = (T)K;
    return T(32) * Eps<T>::v() * k * (T(1) + k * (fabs(umax) + T(2)));
}

constexpr int kSelMaxCand = 16;    // candidate slots per block in the thread-per-block path
constexpr int kSelRounds = 6;      // Michelot refinements before a block is handed to the sorter

// Visits every value of a block once, four 16-byte granules (or four scalars) in flight per
// thread.  The visiting order is rotated by `rot` so that the lanes of a warp, which walk rows
// of equal pitch, hit different banks; the scans that use this are order-independent.
template <typename T, bool CLIP, int KC, class F>
__device__ __forceinline__ void scan_block(const T *blk, int K, int rot, bool vec, F &&f) {
    if constexpr (KC > 0 && (KC & (KC - 1)) == 0 && KC * sizeof(T) >= 64) {
        // power-of-two block size known at compile time: fully unrolled, rotation by masking
        constexpr int VN = 16 / (int)sizeof(T);
        constexpr int KV = KC / VN;
        const int q0 = rot & (KV - 1);
#pragma unroll
        for (int i = 0; i < KV; i += 4) {
            T x[4][VN];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int qu = (q0 + i + u) & (KV - 1);
                if constexpr (VN == 2) {
                    const double2 g = *reinterpret_cast<const double2 *>(blk + qu * 2);
                    x[u][0] = g.x;
                    x[u][1] = g.y;
                } else {
                    const float4 g = *reinterpret_cast<const float4 *>(blk + qu * 4);
                    x[u][0] = g.x;
                    x[u][1] = g.y;
                    x[u][2] = g.z;
                    x[u][3] = g.w;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int j = 0; j < VN; ++j) f(CLIP ? clip_neg(x[u][j]) : x[u][j]);
            }
        }
    } else if (vec) {
        constexpr int VN = 16 / (int)sizeof(T);
        const int KV = K / VN;
        int q = rot % KV;
        for (int i = 0; i < KV; i += 4) {
            T x[4][VN];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int qu = q + u;
                if (qu >= KV) qu -= KV;
                if (i + u < KV) {
                    if constexpr (VN == 2) {
                        const double2 g = *reinterpret_cast<const double2 *>(blk + qu * 2);
                        x[u][0] = g.x;
                        x[u][1] = g.y;
                    } else {
                        const float4 g = *reinterpret_cast<const float4 *>(blk + qu * 4);
                        x[u][0] = g.x;
                        x[u][1] = g.y;
                        x[u][2] = g.z;
                        x[u][3] = g.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + u < KV) {
#pragma unroll
                    for (int j = 0; j < VN; ++j) f(CLIP ? clip_neg(x[u][j]) : x[u][j]);
                }
            }
            q += 4;
            if (q >= KV) q -= KV;
        }
    } else {
        int q = rot % K;
        for (int i = 0; i < K; i += 4) {
            T x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int qu = q + u;
                if (qu >= K) qu -= K;
                if (i + u < K) x[u] = blk[qu];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < K) f(CLIP ? clip_neg(x[u]) : x[u]);
            q += 4;
            if (q >= K) q -= K;
        }
    }
}

// One THREAD computes the shift of one block of K values at blk[] (shared memory).
// cand[i * cstride] are this thread's candidate slots; `vec`: blk is 16-byte aligned and K a
// multiple of the granule.  CLIP: l1-ball mode, negatives count as zero (proj_simplex.h:54-58).
// Returns false when the block keeps more than kSelMaxCand candidates (dense support).
template <typename T, bool CLIP, int KC = 0>
__device__ __forceinline__ bool select_shift_thread(const T *blk, int K, int rot, bool vec, T *cand, int cstride, T &shift_out) {
    // ---- scan 1: the two largest values -> T_0 = u_0 - 1, T_1 = (u_0 + u_1 - 1)/2 ------------------
    T u0 = Num<T>::neg_inf(), u1 = Num<T>::neg_inf();
    scan_block<T, CLIP, KC>(blk, K, rot, vec, [&](T x) {
        const bool top = x > u0;
        const T lo = top ? u0 : x;   // the smaller of (x, u0)
        u0 = top ? x : u0;
        u1 = (lo > u1) ? lo : u1;
    });
    const T delta = select_delta<T>(K, u0);
    T tau = u0 - T(1);
    if (K > 1) {
        const T t1 = ((u0 + u1) - T(1)) * T(0.5);
        tau = (t1 > tau) ? t1 : tau;
    }
    tau -= delta;
    int c = 0, c_prev = K + 1;
    // ---- scans 2..: count, sum and collect the candidates; tighten the bound (Michelot) ---------------
#pragma unroll 1
    for (int round = 0; round < kSelRounds; ++round) {
        T s = T(0);
        c = 0;
        scan_block<T, CLIP, KC>(blk, K, rot, vec, [&](T x) {
            if (x >= tau) {
                s += x;
                if (c < kSelMaxCand) cand[c * cstride] = x;
                ++c;
            }
        });
        if (c <= kSelMaxCand) break;
        if (round >= 1 && 4 * c > 3 * c_prev && c > 2 * kSelMaxCand) break;  // shrinking too slowly: dense support
        c_prev = c;
        const T t2 = (s - T(1)) / (T)c - delta;
        if (!(t2 > tau)) break;
        tau = t2;
    }
    if (c > kSelMaxCand) return false;
    // ---- the reference's loop over the candidates, sorted by a fixed network in registers -----------
    // (branch-free for the common case of at most 8 candidates; -inf pads the tail)
    const T ninf = Num<T>::neg_inf();
    if (c <= 8) {
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (k < c) ? cand[k * cstride] : ninf;
        sort_desc_regs<T, 8>(v);
        shift_out = simplex_shift_sorted<T, 8, 1>(v, c, 0);
    } else {
        T v[kSelMaxCand];
#pragma unroll
        for (int k = 0; k < kSelMaxCand; ++k) v[k] = (k < c) ? cand[k * cstride] : ninf;
        sort_desc_regs<T, kSelMaxCand>(v);
        shift_out = simplex_shift_sorted<T, kSelMaxCand, 1>(v, c, 0);
    }
    return true;
}

// One WARP computes the shift of one block of K values at blk[] (shared memory, lanes stride the
// block).  cand[] is a per-warp list with room for 32 * E_W candidates; they end up in the
// registers of the warp (E_W per lane) and are consumed in descending order with shuffles.
constexpr int kSelWarpRegs = 4;                      // candidate registers per lane
constexpr int kSelWarpCand = 32 * kSelWarpRegs;      // 128 candidates per block at most
constexpr int kSelWarpRounds = 12;

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = (u > v) ? u : v;
    }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, bool CLIP>
__device__ __forceinline__ bool select_shift_warp(const T *blk, int K, int lane, T *cand, T &shift_out) {
    const T ninf = Num<T>::neg_inf();
    T umax = ninf;
    for (int j = lane; j < K; j += 32) {
        T x = blk[j];
        if (CLIP) x = clip_neg(x);
        umax = (x > umax) ? x : umax;
    }
    umax = warp_max(umax);
    const T delta = select_delta<T>(K, umax);
    T tau = (umax - T(1)) - delta;
    int c = 0, c_prev = K + 1;
#pragma unroll 1
    for (int round = 0; round < kSelWarpRounds; ++round) {
        T s = T(0);
        int cl = 0;
        for (int j = lane; j < K; j += 32) {
            T x = blk[j];
            if (CLIP) x = clip_neg(x);
            if (x >= tau) {
                s += x;
                ++cl;
            }
        }
        s = warp_sum(s);
        c = __reduce_add_sync(0xffffffffu, cl);
        // keep tightening while it pays: every candidate less saves a round of shuffles below
        if (c <= 8 || c >= c_prev) break;
        c_prev = c;
        const T t2 = (s - T(1)) / (T)c - delta;
        if (!(t2 > tau)) break;
        tau = t2;
    }
    if (c > kSelWarpCand) return false;
    // ---- compact the candidates into cand[0..c) ---------------------------------------------------
    {
        int base = 0;
        const unsigned lt = (1u << lane) - 1u;
        for (int j0 = 0; j0 < K; j0 += 32) {
            const int j = j0 + lane;
            T x = ninf;
            if (j < K) {
                x = blk[j];
                if (CLIP) x = clip_neg(x);
            }
            const bool in = (j < K) && (x >= tau);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) cand[base + __popc(m & lt)] = x;
            base += __popc(m);
        }
        __syncwarp();
    }
    T r[kSelWarpRegs];
#pragma unroll
    for (int e = 0; e < kSelWarpRegs; ++e) {
        const int i = e * 32 + lane;
        r[e] = (i < c) ? cand[i] : ninf;
    }
    __syncwarp();
    // ---- consume in descending order: every lane tracks the same (sum, num, last) --------------------
    T sum = T(0), num = T(0);
    int last = 0;
    for (int i = 0; i < c; ++i) {
        T lm = r[0];
#pragma unroll
        for (int e = 1; e < kSelWarpRegs; ++e) lm = (r[e] > lm) ? r[e] : lm;
        const T v = warp_max(lm);
        // (S_{i-1} - 1)/i bounds the threshold from below: a value more than delta under it, and
        // every later one, is rejected by the reference -- stop (multiplied out: no division)
        if (i > 0 && fma(v, (T)i, delta * (T)i) < sum - T(1)) break;
        // exactly one holder gives its copy up (lowest lane, lowest register)
        const unsigned holders = __ballot_sync(0xffffffffu, lm == v);
        if (lane == __ffs(holders) - 1) {
            bool done = false;
#pragma unroll
            for (int e = 0; e < kSelWarpRegs; ++e) {
                if (!done && r[e] == v) {
                    r[e] = ninf;
                    done = true;
                }
            }
        }
        if (i == 0) {
            sum = v;
            num = T(1) - sum;
        } else {
            sum += v;
            const T w = T(1) - sum;
            const int cls = candidate_class<T>(v, w, T(i) + T(1));
            bool pass = cls > 0;
            if (cls == 0) pass = (v + w / (T(i) + T(1))) > T(0);
            if (pass) {
                num = w;
                last = i;
            }
        }
    }
    shift_out = num / (T(last) + T(1));
    return true;
}

}  // namespace bsls
