// solver_cluster.cuh -- the device-resident BATCH loop of solver_tiny.cuh on a thread-block CLUSTER of 8 CTAs.
//
// One CTA runs config 1 (n = 5000 routes, m = 2000 links, 50 000 entries) at ~45 000 iterations/s: a thread owns a
// whole OD block in the projection (a dependent chain of ~600 fp64 instructions), the index arrays of A and A^T stream from
// L2 every iteration (the vectors fill the shared memory, so L1 holds next to nothing) and one thread adds the 32 warp
// partials of every sum.  A cluster spreads that chain over 8 SMs and, with a share of an eighth of the matrix per CTA,
// has room for the index arrays ON CHIP:
//
//   * every CTA keeps the full x (2 buffers) and r (2 buffers) in its shared memory, plus -- for ITS rows of A, ITS
//     columns of A^T and ITS OD blocks only -- the sliced-ELL index arrays, b, the block starts and the two gradient
//     buffers.  After the prologue the loop loads nothing from global memory (value arrays of a general A excepted).
//   * projection: G lanes per OD block (register sorting network across lanes, simplex_core.cuh) for the CTA's blocks;
//     the CTA's slice of the new x then goes to the x buffer of the 7 other CTAs.
//   * r = A x - b for the CTA's link rows; the slice and the CTA's partial sums go to the 7 other CTAs.
//   * g = A^T r for the CTA's routes (kept local: only its own blocks step along it) with the BB sums; partial sums to
//     the other CTAs; thread 0 of EVERY CTA adds the partials in rank order and takes the decision on its own copy of
//     the solver state -- identical inputs, identical code, identical result: no further exchange.
//
// The exchange: a slice is ONE bulk copy per peer, shared memory to the peer's shared memory (cp.async.bulk
// shared::cluster <- shared::cta, issued by 7 threads), that completes on an mbarrier of the RECEIVER: a CTA waits until
// the bytes it expects for a phase have landed, nobody waits for a cluster-wide barrier (whose release is a
// MEMBAR.ALL.GPU: a third of the first version's time, profiles/r02_c1_cluster_solver_stalls.txt) and no thread issues
// remote stores one value at a time (the other third).  Odd ends of a slice (16-byte alignment) travel as single
// st.async values on the same mbarrier.  No receiver-to-sender handshake is needed: a CTA can only send its slice of
// phase p+1 after it has received every slice of phase p, and a peer sends that only after it has finished reading what
// phase p+1 overwrites (x and r are double-buffered; the one exception, the pull-back of a back-tracked point, ends in
// a full cluster barrier).
//
// Column shares are aligned to OD blocks, link shares to groups of 32 rows.  A column share starts at the first block at or
// after column 32 * (groups * q / 8), so the host can bound the shared-memory need of a share from the group offsets
// alone (cluster_share_end: at most max_k - 1 columns of halo), whatever the block layout is.
#pragma once
#include <cooperative_groups.h>

#include <type_traits>

#include "solver_tiny.cuh"

namespace bsls {

constexpr int kClusterCtas = 8;  // the portable maximum

struct ClusterLayout {
    int groups_a, groups_t;  // 32-row groups of A (links) and A^T (routes)
    int max_k;               // longest OD block (bounds the halo of a column share)
    // arena offsets, the same in every CTA (sized for the largest share)
    int x_off, r_off, g_off, g_cap, b_off, sc_off, part_off;  // in doubles
    int ia_off, it_off, ga_off, gt_off, st_off;               // in 32-bit words
};

__host__ __device__ inline int cluster_share(int groups, int q) { return (int)((long long)groups * q / kClusterCtas); }
// one past the last group a column share can touch: its blocks start at or after 32 * share(q) and the first block of the
// next share starts before 32 * share(q + 1) + max_k
__host__ __device__ inline int cluster_share_end(int groups, int q, int max_k) {
    if (q == kClusterCtas - 1) return groups;
    const int e = cluster_share(groups, q + 1) + (max_k >= 2 ? (max_k - 2) / 32 : 0) + 1;
    return e < groups ? e : groups;
}

__device__ __forceinline__ uint32_t cluster_map(uint32_t shared_addr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(shared_addr), "r"(rank));
    return r;
}
// one value into a peer's shared memory, counted on the peer's mbarrier when it lands
__device__ __forceinline__ void cluster_store_tx(uint32_t addr, double v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr), "l"(__double_as_longlong(v)), "r"(mbar)
                 : "memory");
}
// `bytes` (a multiple of 16, both addresses 16-byte aligned) from my shared memory into a peer's, by the bulk-copy
// engine; counted on the peer's mbarrier when they have landed
__device__ __forceinline__ void cluster_copy_tx(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(src), "r"(bytes),
                 "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// (the default .acquire.cta form: what lands in SHARED memory needs no L1 invalidation; the .acquire.cluster form emits a
// CCTL.IVALL per waiting warp, which was a quarter of the kernel's stall samples)
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
// my generic-proxy writes to shared memory, made visible to the bulk-copy engine that reads them next
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// HV: A carries values (read from global memory; the index arrays are on chip either way).  E x G: registers per lane x
// lanes per OD block of the projection (E * G >= max_k).
template <bool HV, int E, int G>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kTinyThreads, 1) solver_cluster_kernel(TinyArgs a, DevOpts o, ClusterLayout L) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int C = kClusterCtas, T = kTinyThreads;
    const int me = (int)cluster.block_rank();
    extern __shared__ __align__(16) double tiny_sm[];
    int32_t *smi = reinterpret_cast<int32_t *>(tiny_sm);
    const int n = a.n, m = a.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n1 = n + 1, m1 = m + 1;
    double *scal = tiny_sm + L.sc_off;
    __shared__ double s_red2[2][32 * 5];  // per slot: warps of a CTA that is through with the residual's sums may already write the gradient's
    __shared__ __align__(8) uint64_t s_mbar[3];  // bytes landed in my shared memory: x slices / r slices + sums / gradient sums
    __shared__ int s_blk[2];
    __shared__ DevState s_state;
    __shared__ long long prof_acc[12];
    DevState *st = &s_state;

    // ---- my share ------------------------------------------------------------------------------------------
    const int ga0 = cluster_share(L.groups_a, me), ga1 = cluster_share(L.groups_a, me + 1);
    const int gt0 = cluster_share(L.groups_t, me), gt1 = cluster_share_end(L.groups_t, me, L.max_k);
    const int row0 = 32 * ga0, col0 = 32 * gt0;
    if (tid < 2) {  // first block at or after the share's first column (the next share's: my end)
        const int target = 32 * cluster_share(L.groups_t, me + tid);
        int lo = 0, hi = a.nb;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a.starts[mid] >= target)
                hi = mid;
            else
                lo = mid + 1;
        }
        s_blk[tid] = lo;
    }
    if (tid == 0) s_state = DevState{};
    if (tid < 12) prof_acc[tid] = 0;
    const uint32_t mb_x = (uint32_t)__cvta_generic_to_shared(&s_mbar[0]), mb_r = mb_x + 8, mb_g = mb_x + 16;
    if (tid == 0) {
        mbar_init(mb_x, 1);
        mbar_init(mb_r, 1);
        mbar_init(mb_g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int b_lo = s_blk[0], nbl = s_blk[1] - s_blk[0];
    const int c_lo = a.starts[b_lo], c_hi = a.starts[b_lo + nbl];
    const int oa = a.A.goff[ga0], ot = a.AT.goff[gt0];
    {
        const int na = a.A.goff[ga1] - oa, nt = a.AT.goff[gt1] - ot;
        for (int i = tid; i < na; i += T) smi[L.ia_off + i] = a.A.idx[oa + i];
        for (int i = tid; i < nt; i += T) smi[L.it_off + i] = a.AT.idx[ot + i];
        for (int i = tid; i <= ga1 - ga0; i += T) smi[L.ga_off + i] = a.A.goff[ga0 + i] - oa;
        for (int i = tid; i <= gt1 - gt0; i += T) smi[L.gt_off + i] = a.AT.goff[gt0 + i] - ot;
        for (int i = tid; i <= nbl; i += T) smi[L.st_off + i] = a.starts[b_lo + i];
        for (int i = tid; i < n; i += T) tiny_sm[L.x_off + i] = a.x[i];
        for (int i = tid; i < 32 * (ga1 - ga0); i += T) tiny_sm[L.b_off + i] = row0 + i < m ? a.b[row0 + i] : 0.0;
        if (tid < 2) {  // the slot padding entries point at
            tiny_sm[L.x_off + tid * n1 + n] = 0.0;
            tiny_sm[L.r_off + tid * m1 + m] = 0.0;
        }
        if (tid < kScalCount) scal[tid] = 0.0;
    }
    const uint32_t arena = (uint32_t)__cvta_generic_to_shared(tiny_sm);
    // bytes a phase delivers into my shared memory (the other CTAs' slices; their 3 or 5 partial sums)
    const int my_rows = max(0, min(m, 32 * ga1) - row0);
    const uint32_t x_bytes = 8u * (uint32_t)(n - (c_hi - c_lo)), r_bytes = 8u * (uint32_t)(m - my_rows) + 24u * (C - 1), g_bytes = 40u * (C - 1);
    uint32_t ph_x = 0, ph_r = 0, ph_g = 0;
    if (tid == 0) {  // the expectation of a phase is posted as soon as the previous one is through
        mbar_expect(mb_x, x_bytes);
        mbar_expect(mb_r, r_bytes);
        mbar_expect(mb_g, g_bytes);
    }
    // my entries [lo, hi) of the vector at arena offset `vec` (doubles) to the same place in CTA q, counted on its
    // mbarrier `mb`: the 16-byte aligned middle as one bulk copy, an odd first / last entry as single values
    auto send_slice = [&](int q, int vec, int lo, int hi, uint32_t mb) {
        const uint32_t peer = cluster_map(arena, q), pmb = cluster_map(mb, q);
        const int a0 = lo + ((vec + lo) & 1), a1 = hi - ((vec + hi) & 1);  // the aligned middle [a0, a1)
        if (a1 > a0) {
            cluster_copy_tx(peer + 8u * (uint32_t)(vec + a0), arena + 8u * (uint32_t)(vec + a0), 8u * (uint32_t)(a1 - a0), pmb);
            if (a0 > lo) cluster_store_tx(peer + 8u * (uint32_t)(vec + lo), tiny_sm[vec + lo], pmb);
            if (a1 < hi) cluster_store_tx(peer + 8u * (uint32_t)(vec + hi - 1), tiny_sm[vec + hi - 1], pmb);
        } else {  // at most two entries
            for (int i = lo; i < hi; ++i) cluster_store_tx(peer + 8u * (uint32_t)(vec + i), tiny_sm[vec + i], pmb);
        }
    };
    // phase clock of thread 0 of CTA 0 (development aid; a.prof is null in production)
    const bool prof = a.prof != nullptr && me == 0 && tid == 0;
    long long prof_last = 0;
    auto stamp = [&](int k) {
        if (prof) {
            const long long c = clock64();
            prof_acc[k] += c - prof_last;
            prof_last = c;
        }
    };
    cluster.sync();

    // one value to the same place in every other CTA, counted on their mbarrier `mb`
    auto to_peers = [&](int off /* doubles */, double v, uint32_t mb) {
#pragma unroll
        for (int q = 0; q < C; ++q)
            if (q != me) cluster_store_tx(cluster_map(arena, q) + 8u * (uint32_t)off, v, cluster_map(mb, q));
    };
    // CTA sums (fixed tree: lanes, then warps) -> row (slot, me) of the partial-sum table of every CTA; then wait until
    // everything the other CTAs send in this phase (their sums, and their rows of r in the residual's phase) is here
    auto publish = [&](auto &acc, auto ns_tag, int slot, uint32_t mb, uint32_t &phase, uint32_t bytes) {
        constexpr int NS = decltype(ns_tag)::value;
        constexpr int N = sizeof(acc) / sizeof(double);
        double *s_red = s_red2[slot];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double v = acc[k];
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double u = __shfl_xor_sync(0xffffffffu, v, d);
                v = (k < NS) ? v + u : fmax(v, u);
            }
            if (lane == 0) s_red[wid * N + k] = v;
        }
        __syncthreads();
        if (wid == 0) {
            const int row = L.part_off + (slot * C + me) * 8;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                double v = s_red[lane * N + k];  // T / 32 == 32 warps
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const double u = __shfl_xor_sync(0xffffffffu, v, d);
                    v = (k < NS) ? v + u : fmax(v, u);
                }
                if (lane == 0) tiny_sm[row + k] = v;
                if (lane < C && lane != me)  // the butterfly left the total in every lane: lane q sends to CTA q
                    cluster_store_tx(cluster_map(arena, lane) + 8u * (uint32_t)(row + k), v, cluster_map(mb, lane));
            }
        }
        mbar_wait(mb, phase);
        phase ^= 1u;
        if (tid == 0) mbar_expect(mb, bytes);  // the next use of this barrier
    };
    // thread 0: the cluster's totals in rank order (the same bits in every CTA)
    auto total = [&](int slot, int k, bool is_max) {
        const double *p = tiny_sm + L.part_off + slot * C * 8 + k;
        double v = p[0];
#pragma unroll
        for (int q = 1; q < C; ++q) v = is_max ? fmax(v, p[q * 8]) : v + p[q * 8];
        return v;
    };

    // r = A x - b for my link rows -> every CTA; x, r, r_old: arena offsets (r_old < 0: none)
    auto residual = [&](int x, int r, int r_old) {
        double acc[3] = {0, 0, 0};
        const char *base = reinterpret_cast<const char *>(tiny_sm + x);
        for (int lr = tid; lr < 32 * (ga1 - ga0); lr += T) {
            const int row = row0 + lr;
            if (row < m) {
                const int g = lr >> 5, off = smi[L.ga_off + g], width = (smi[L.ga_off + g + 1] - off) >> 5;
                const int32_t *ip = smi + L.ia_off + off + lane;
                const double *vp = HV ? a.A.val + oa + off + lane : nullptr;
                double sum = 0.0;
                for (int k = 0; k < width; k += 4) {  // widths are multiples of 4, padding entries point at the 0.0 slot
                    int j[4];
                    double av[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) j[u] = ip[32 * (k + u)];
                    if (HV) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) av[u] = vp[32 * (k + u)];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double xv = *reinterpret_cast<const double *>(base + j[u]);
                        sum += HV ? av[u] * xv : xv;
                    }
                }
                const double v = sum - tiny_sm[L.b_off + lr];
                tiny_sm[r + row] = v;
                to_peers(r + row, v, mb_r);
                residual_sums(v, r_old >= 0 ? tiny_sm[r_old + row] : 0.0, r_old >= 0, acc);
            }
        }
        stamp(2);
        publish(acc, std::integral_constant<int, 3>{}, 0, mb_r, ph_r, r_bytes);
        stamp(3);
    };
    // g_new = A^T r for my routes (local) and the sums of EpiGradBB (g < 0: only <g_new, g_new>)
    auto gradient = [&](int r, int g_new, int g_old, int x, int x_new) {
        double acc[5] = {0, 0, 0, 0, 0};
        const char *base = reinterpret_cast<const char *>(tiny_sm + r);
        for (int lc = tid; lc < 32 * (gt1 - gt0); lc += T) {
            const int col = col0 + lc;
            if (col >= c_lo && col < c_hi) {
                const int g = lc >> 5, off = smi[L.gt_off + g], width = (smi[L.gt_off + g + 1] - off) >> 5;
                const int32_t *ip = smi + L.it_off + off + lane;
                const double *vp = HV ? a.AT.val + ot + off + lane : nullptr;
                double dot = 0.0;
                for (int k = 0; k < width; k += 4) {
                    int j[4];
                    double av[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) j[u] = ip[32 * (k + u)];
                    if (HV) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) av[u] = vp[32 * (k + u)];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double rv = *reinterpret_cast<const double *>(base + j[u]);
                        dot += HV ? av[u] * rv : rv;
                    }
                }
                tiny_sm[g_new + lc] = dot;
                acc[3] += dot * dot;
                if (g_old >= 0) {
                    const double go = tiny_sm[g_old + lc], dx = tiny_sm[x_new + col] - tiny_sm[x + col], dg = dot - go;
                    acc[0] += dx * dg;
                    acc[1] += dg * dg;
                    acc[2] += go * dx;
                    acc[4] = fmax(acc[4], fabs(dx));
                }
            }
        }
        stamp(4);
        publish(acc, std::integral_constant<int, 4>{}, 1, mb_g, ph_g, g_bytes);
        stamp(5);
    };
    auto decide = [&](int first) {
        if (tid == 0) {
            const double rr = total(0, 0, false);
            scal[kScalF] = 0.5 * rr;
            scal[kScalRR] = rr;
            scal[kScalRdr] = total(0, 1, false);
            scal[kScalDrdr] = total(0, 2, false);
            scal[kScalSxy] = total(1, 0, false);
            scal[kScalSyy] = total(1, 1, false);
            scal[kScalGd] = total(1, 2, false);
            scal[kScalGnn] = total(1, 3, false);
            scal[kScalStep] = total(1, 4, true);
            decide_step(st, scal, nullptr, o, me == 0 ? a.progress_f : nullptr, me == 0 ? a.progress_t : nullptr, first);
        }
        stamp(6);
        __syncthreads();
        stamp(7);
    };

    const int X = L.x_off, R = L.r_off, GR = L.g_off;
    residual(X, R, -1);
    gradient(R, GR, -1, -1, -1);
    decide(1);

    constexpr int GROUPS = T / G;
    const int sub = tid & (G - 1), grp = tid / G;
    const double ninf = Num<double>::neg_inf();
    int cur = 0;
    if (prof) {
        for (int k = 0; k < 12; ++k) prof_acc[k] = 0;
        prof_last = clock64();
    }
    while (!st->done) {
        const int nxt = cur ^ 1;
        const int xc = X + cur * n1, xn = X + nxt * n1, gc = GR + cur * L.g_cap, gn = GR + nxt * L.g_cap, rc = R + cur * m1, rn = R + nxt * m1;
        const double nt = -st->t;
        // ---- x_new = proj(x - t g) for my OD blocks, G lanes per block, written to every CTA --------------------------
        for (int lb0 = 0; lb0 < nbl; lb0 += GROUPS) {  // uniform trip count: the shuffles below need whole warps
            const int lb = lb0 + grp;
            const bool live = lb < nbl;
            int s = 0, K = 0;
            if (live) {
                s = smi[L.st_off + lb];
                K = smi[L.st_off + lb + 1] - s;
            }
            double raw[E], v[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = sub * E + e;
                double w = ninf;
                if (pos < K) {
                    const double u = nt * tiny_sm[gc + (s + pos - col0)];  // np.add(x, -t*g, x_new): product and sum rounded separately
                    w = tiny_sm[xc + s + pos] + u;
                    if (a.proj_mode == 1) w = clip_neg(w);                 // proj_multi_ball (proj_simplex.h:54-62)
                }
                raw[e] = v[e] = w;
            }
            bool project = true;
            if (a.proj_mode == 1) {  // only blocks whose clipped values sum to more than 1, summed in the reference's order
                double run = 0.0;
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    if (sub == q) {
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            if (sub * E + e < K) run += raw[e];
                    }
                    run = __shfl_sync(0xffffffffu, run, (lane & ~(G - 1)) + q);
                }
                project = run > 1.0;
            }
            sort_desc_group<double, E, G>(v, lane);
            const double shift = simplex_shift_sorted<double, E, G>(v, K, lane);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = sub * E + e;
                if (pos < K) {
                    double w = raw[e];
                    if (project) {
                        w = shift + w;
                        w = (w < 0.0) ? 0.0 : w;
                    }
                    tiny_sm[xn + s + pos] = w;
                }
            }
        }
        stamp(0);
        fence_async_smem();
        __syncthreads();
        if (tid < C && tid != me && c_hi > c_lo) send_slice(tid, xn, c_lo, c_hi, mb_x);
        mbar_wait(mb_x, ph_x);
        ph_x ^= 1u;
        if (tid == 0) mbar_expect(mb_x, x_bytes);
        stamp(1);
        residual(xn, rn, rc);
        gradient(rn, gn, gc, xc, xn);
        decide(0);
        const double tau = st->tau;
        if (tau != 1.0) {  // commit_kernel's pull-back on every CTA's copies (the same bits everywhere) and on my gradient
            const double c = 1.0 - tau;
            for (int i = tid; i < n; i += T) {
                if (tau == 0.0) {
                    tiny_sm[xn + i] = tiny_sm[xc + i];
                } else {
                    const double u = c * tiny_sm[xc + i];
                    tiny_sm[xn + i] = u + tau * tiny_sm[xn + i];
                }
            }
            for (int i = tid; i < 32 * (gt1 - gt0); i += T) {
                if (tau == 0.0) {
                    tiny_sm[gn + i] = tiny_sm[gc + i];
                } else {
                    const double u = c * tiny_sm[gc + i];
                    tiny_sm[gn + i] = u + tau * tiny_sm[gn + i];
                }
            }
            for (int i = tid; i < m; i += T) {
                if (tau == 0.0) {
                    tiny_sm[rn + i] = tiny_sm[rc + i];
                } else {
                    const double u = c * tiny_sm[rc + i];
                    tiny_sm[rn + i] = u + tau * tiny_sm[rn + i];
                }
            }
            // the pull-back read buffer `cur` of x and r, which the peers overwrite with their slices of the next trial
            // point: they must not start before every CTA is through (tau is the same everywhere: a cluster-uniform branch)
            fence_async_smem();
            cluster.sync();
            stamp(8);
        }
        cur = nxt;
    }
    cluster.sync();  // nobody leaves while a peer may still write into its shared memory
    if (me == 0) {
        for (int i = tid; i < n; i += T) a.x[i] = tiny_sm[X + cur * n1 + i];
        if (tid == 0) *a.st = s_state;
        if (prof)
            for (int k = 0; k < 12; ++k) a.prof[k] = prof_acc[k];
    }
}

}  // namespace bsls
