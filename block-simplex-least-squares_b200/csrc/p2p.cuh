// p2p.cuh -- the link-vector reduction of a sharded solve over NVLink peer memory, fused with the kernels around it.
//
// Every objective evaluation of a sharded solve ends its product r_p = A_p x_p with a sum over ranks, r = sum_p r_p - b,
// followed by |r|^2 and the two line-search sums (lsq.cuh, residual_sums); the gradient kernel then leaves this rank's
// share of five step scalars that every rank needs in full.  With NCCL that is: panel-reduce kernel, ncclAllReduce (8 MB),
// finish kernel, ..., ncclAllGather (48 B) -- about 0.1 ms of a 1.8 ms iteration at 8 GPUs, none of it overlapped.
//
// Here every rank owns an exchange region that all its peers map (cudaIpc), and the reduction is done by the kernels
// themselves with loads / stores over NVLink:
//
//   p2p_reduce_kernel    rank p owns the rows [m p / P, m (p+1) / P).  It announces its partial vector (flag A), waits for
//                        the flags of the others, pulls ITS rows from every rank's partial vector (P2P loads, added in
//                        rank order: every rank obtains the same bits whatever the timing), subtracts b, accumulates the
//                        three sums, and pushes the finished rows into the residual buffer of EVERY rank (P2P stores):
//                        reduce-scatter + subtract + norms + all-gather in one pass over m / P rows.  Its last CTA
//                        publishes the rank's three partial sums and raises flag B everywhere.
//   p2p_wait_kernel      one warp: waits for flag B of all ranks (the whole residual has arrived), adds the P partial
//                        sums in rank order into the scalar block.
//   p2p_post_kernel      one warp: pushes this rank's step scalars to every rank and raises flag C; decide_kernel waits for
//                        the P flags and adds the shares in rank order (the all-gather of the NCCL build).
//
// Flags are 64-bit epoch counters (one evaluation = one epoch, counted identically on every rank), written with a
// system-scope release after a system-scope fence and polled with system-scope acquires; data a flag guards is read
// with ld.volatile / __ldcv so that no stale L1 line of an earlier epoch can be seen (peer memory is cached in L1 only).
// A rank never waits for something that depends on its own later progress: flag A is raised before anything is waited
// for, flag B needs only flags A, flag C only local work -- no cycle, whatever the relative speed of the ranks.
// Buffer reuse is safe by the same chain: a peer writes into my residual buffer of parity k only after my flag A of
// evaluation k, which my stream raises after everything that read the old content.
#pragma once
#include "lsq.cuh"

namespace bsls {

constexpr int kP2pMaxRanks = 8;  // one NVSwitch domain

struct P2pView {
    int nranks, rank;
    double *partial[kP2pMaxRanks];             // m: the rank's partial product (read by the owners of the rows)
    double *rfull[kP2pMaxRanks][2];            // m each: the two residual buffers of the rank (written by the row owners)
    double *slots[kP2pMaxRanks];               // nranks x 8: residual sums of every source rank
    double *gath[kP2pMaxRanks];                // nranks x kStepScalars: step scalars of every source rank
    unsigned long long *flags[kP2pMaxRanks];   // 3 x nranks: flag A / B / C of every source rank
    unsigned long long *prof;                  // development (BSLS_P2P_PROF=1): ns waited for flags A / spent reducing / waited for flags B, calls; else null
};

__device__ __forceinline__ unsigned long long p2p_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// wait until the first `count` flags of `f` have reached `epoch` (threads 0..count-1 of the CTA poll one flag each)
__device__ __forceinline__ void p2p_wait_flags(const unsigned long long *f, int count, unsigned long long epoch) {
    if ((int)threadIdx.x < count)
        while (ld_acquire_sys(f + threadIdx.x) < epoch) {
        }
    __syncthreads();
}

constexpr int kP2pThreads = 256;

__global__ void __launch_bounds__(kP2pThreads)
p2p_reduce_kernel(P2pView v, const double *__restrict__ b, int64_t m, int nxt, unsigned long long epoch, unsigned *ticket,
                  double *cta_partials /* gridDim.x * 3 */, const int *__restrict__ skip) {
    if (skip && *skip) return;
    const int P = v.nranks, me = v.rank, tid = threadIdx.x;
    __shared__ double s_w[kP2pThreads / 32][3];
    __shared__ bool s_last;
    // flag A: my partial vector is complete (the kernels before this one in the stream wrote it)
    if (blockIdx.x == 0 && tid < P) {
        __threadfence_system();
        st_release_sys(v.flags[tid] + 0 * P + me, epoch);
    }
    const unsigned long long t0 = (v.prof && blockIdx.x == 0 && tid == 0) ? p2p_now() : 0ull;
    p2p_wait_flags(v.flags[me] + 0 * P, P, epoch);
    if (v.prof && blockIdx.x == 0 && tid == 0) {
        const unsigned long long t1 = p2p_now();
        v.prof[0] += t1 - t0;
        v.prof[6] = t1;
    }
    const int64_t lo = m * me / P, hi = m * (me + 1) / P;
    const double *r_old = v.rfull[me][nxt ^ 1];
    double acc[3] = {0, 0, 0};
    // two rows per thread and trip: 2 x P loads over NVLink in flight before the first sum
    const int64_t stride = (int64_t)gridDim.x * kP2pThreads;
    for (int64_t i = lo + (int64_t)blockIdx.x * kP2pThreads + tid; i < hi; i += 2 * stride) {
        const int64_t j = i + stride;
        const bool two = j < hi;
        double s0 = __ldcv(v.partial[0] + i), s1 = two ? __ldcv(v.partial[0] + j) : 0.0;
        double t0v[kP2pMaxRanks], t1v[kP2pMaxRanks];
#pragma unroll
        for (int q = 1; q < kP2pMaxRanks; ++q) {
            t0v[q] = q < P ? __ldcv(v.partial[q] + i) : 0.0;
            t1v[q] = (q < P && two) ? __ldcv(v.partial[q] + j) : 0.0;
        }
#pragma unroll
        for (int q = 1; q < kP2pMaxRanks; ++q)
            if (q < P) {  // rank order: the same bits on every rank
                s0 += t0v[q];
                s1 += t1v[q];
            }
        const double r0 = s0 - b[i];
        residual_sums(r0, r_old[i], true, acc);
        for (int q = 0; q < P; ++q) v.rfull[q][nxt][i] = r0;
        if (two) {
            const double r1 = s1 - b[j];
            residual_sums(r1, r_old[j], true, acc);
            for (int q = 0; q < P; ++q) v.rfull[q][nxt][j] = r1;
        }
    }
    // deterministic sum over the CTA, then over the grid by the last CTA (fixed order).  One system-scope fence per CTA,
    // by the thread that takes the ticket after the CTA barrier (fences are cumulative: the barrier orders the other
    // threads' peer stores before it); flag B must not overtake them
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double x = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((tid & 31) == 0) s_w[tid >> 5][k] = x;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        for (int k = 0; k < 3; ++k) {
            double x = s_w[0][k];
            for (int w = 1; w < kP2pThreads / 32; ++w) x += s_w[w][k];
            cta_partials[blockIdx.x * 3 + k] = x;
        }
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {   // the grid's sums: every thread adds a strided share of the CTA partials, then the fixed tree over the CTA
        double g[3] = {0, 0, 0};
        for (unsigned c = tid; c < gridDim.x; c += kP2pThreads)
            for (int k = 0; k < 3; ++k) g[k] += __ldcg(cta_partials + c * 3 + k);
        __syncthreads();  // s_w is reused
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double x = g[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((tid & 31) == 0) s_w[tid >> 5][k] = x;
        }
        __syncthreads();
        if (tid < 3) {
            double x = s_w[0][tid];
            for (int w = 1; w < kP2pThreads / 32; ++w) x += s_w[w][tid];
            for (int q = 0; q < P; ++q) v.slots[q][me * 8 + tid] = x;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (tid < P) st_release_sys(v.flags[tid] + 1 * P + me, epoch);  // flag B: my rows and my sums are everywhere
    if (tid == 0) *ticket = 0u;
    if (v.prof && tid == 0) v.prof[1] += p2p_now() - v.prof[6];
}

// The residual is complete on this rank once every rank has raised flag B; the scalar block gets the sums in rank order.
__global__ void p2p_wait_kernel(P2pView v, unsigned long long epoch, double *scal, const int *__restrict__ skip) {
    if (skip && *skip) return;
    const int P = v.nranks, me = v.rank;
    const unsigned long long t0 = (v.prof && threadIdx.x == 0) ? p2p_now() : 0ull;
    p2p_wait_flags(v.flags[me] + 1 * P, P, epoch);
    if (v.prof && threadIdx.x == 0) {
        v.prof[2] += p2p_now() - t0;
        v.prof[3] += 1;
    }
    if (threadIdx.x == 0) {
        double s[3] = {0, 0, 0};
        for (int q = 0; q < P; ++q)
            for (int k = 0; k < 3; ++k) s[k] += __ldcv(v.slots[me] + q * 8 + k);
        scal[kScalF] = 0.5 * s[0];
        scal[kScalRR] = s[0];
        scal[kScalRdr] = s[1];
        scal[kScalDrdr] = s[2];
    }
}

// this rank's step scalars (slots 1..kStepScalars of the scalar block) to every rank; flag C
__global__ void p2p_post_kernel(P2pView v, unsigned long long epoch, const double *scal, const int *__restrict__ skip) {
    if (skip && *skip) return;
    const int P = v.nranks, me = v.rank, tid = threadIdx.x;
    if (tid < kStepScalars) {
        const double x = scal[kScalSxy + tid];
        for (int q = 0; q < P; ++q) v.gath[q][me * kStepScalars + tid] = x;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < P) st_release_sys(v.flags[tid] + 2 * P + me, epoch);
}

}  // namespace bsls
