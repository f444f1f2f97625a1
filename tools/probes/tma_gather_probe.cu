// tma_gather_probe.cu -- development probe (not part of the library): how many random 16-byte gathers per second can the
// bulk-copy engine (cp.async.bulk global -> shared, SASS UBLKCP) sustain next to / instead of LSU gathers (LDG)?
// The SpMV kernels are bound by the L1TEX wavefront rate of their LDG gathers (~1 per clock per SM); if bulk copies
// travel another path, a hybrid kernel could exceed that.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather_probe tma_gather_probe.cu && ./tma_gather_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: LDG gathers only; mode 1: bulk-copy gathers only; mode 2: half and half
template <int PER, int MODE>
__global__ void __launch_bounds__(256) probe(const double *__restrict__ v, const int32_t *__restrict__ idx, int64_t rows, double *out) {
    __shared__ __align__(16) double stage[256 * PER * 2];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(256));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc = 0.0;
    uint32_t parity = 0;
    for (int64_t row = (int64_t)blockIdx.x * 256 + tid; row < rows; row += (int64_t)gridDim.x * 256) {
        int32_t j[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) j[k] = __ldcs(idx + row * PER + k);
        constexpr int NB = MODE == 0 ? 0 : (MODE == 1 ? PER : PER / 2);
        if (NB > 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(16 * NB) : "memory");
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const double *src = v + (j[k] & ~1);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_addr(&stage[(tid * PER + k) * 2])),
                             "l"(src), "r"(16), "r"(smem_addr(&bar))
                             : "memory");
            }
        }
        double w[PER];
#pragma unroll
        for (int k = NB; k < PER; ++k) w[k] = __ldcg(v + j[k]);
        if (NB > 0) {
            asm volatile(
                "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_addr(&bar)),
                "r"(parity)
                : "memory");
            parity ^= 1;
#pragma unroll
            for (int k = 0; k < NB; ++k) w[k] = stage[(tid * PER + k) * 2 + (j[k] & 1)];
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) acc += w[k];
        if (NB > 0) __syncthreads();  // the stage is reused by the next round
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int MODE> float run(const double *v, const int32_t *idx, int64_t rows, double *out, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) probe<8, MODE><<<grid, 256>>>(v, idx, rows, out);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) probe<8, MODE><<<grid, 256>>>(v, idx, rows, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return ms / 5;
}

int main() {
    const int64_t m = 1000000, rows = 20000000;
    double *v, *out;
    int32_t *idx;
    cudaMalloc(&v, (m + 2) * sizeof(double));
    cudaMalloc(&out, 8);
    cudaMalloc(&idx, rows * 8 * sizeof(int32_t));
    int32_t *h = (int32_t *)malloc(rows * 8 * sizeof(int32_t));
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < rows * 8; ++i) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        h[i] = (int32_t)(s % m);
    }
    cudaMemcpy(idx, h, rows * 8 * sizeof(int32_t), cudaMemcpyHostToDevice);
    cudaMemset(v, 0, (m + 2) * sizeof(double));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int per = 2; per <= 8; per *= 2) {
        const int grid = sms * per;
        const float t0 = run<0>(v, idx, rows, out, grid), t1 = run<1>(v, idx, rows, out, grid), t2 = run<2>(v, idx, rows, out, grid);
        printf("CTAs/SM %d: LDG %.3f ms (%.1f G gathers/s) | bulk %.3f ms (%.1f G/s) | half/half %.3f ms (%.1f G/s)\n", per, t0,
               rows * 8 / t0 * 1e-6, t1, rows * 8 / t1 * 1e-6, t2, rows * 8 / t2 * 1e-6);
    }
    return 0;
}
