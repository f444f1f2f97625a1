// red_scatter_probe.cu -- development probe (not part of the library): `A x` from the CSC side.
// The round-1 review asked to try r += A[:, j] x_j with x read coalesced and the updates done by RED.ADD.F64 into the
// L2-resident 8 MB link vector, instead of gathering x from L2 row by row (one launch per column panel).  This probe
// measures what that scatter can sustain on config-5 shapes (8 link ids per route, 10^6 links) next to the gather it
// would replace: thread per route, two 16-byte index loads, 8 atomics (scatter) or 8 loads (gather) in flight.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_scatter_probe red_scatter_probe.cu && ./red_scatter_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int L = 8;

// mode 0: gather r[idx] (what A^T r does); mode 1: RED.ADD.F64 x_j into r[idx] (what A x from the CSC side would do)
template <int MODE>
__global__ void __launch_bounds__(256) probe(double *__restrict__ r, const double *__restrict__ x, const int4 *__restrict__ idx, int64_t cols,
                                             double *out) {
    double acc = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < cols; c += (int64_t)gridDim.x * 256) {
        const int4 a = __ldcs(idx + 2 * c), b = __ldcs(idx + 2 * c + 1);
        const int j[L] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < L; ++k) acc += __ldcg(r + j[k]);
        } else {
            const double v = __ldcs(x + c);
#pragma unroll
            for (int k = 0; k < L; ++k) atomicAdd(r + j[k], v);  // result unused: compiles to RED.E.ADD.F64
        }
    }
    if (MODE == 0 && acc == 12345.678) out[0] = acc;
}

int main(int argc, char **argv) {
    const int64_t cols = argc > 1 ? atoll(argv[1]) : 160000000;  // config 5: 1.6e8 routes
    const int64_t m = argc > 2 ? atoll(argv[2]) : 1000000;
    int4 *idx;
    double *r, *x, *out;
    cudaMalloc(&idx, sizeof(int4) * 2 * cols);
    cudaMalloc(&r, sizeof(double) * m);
    cudaMalloc(&x, sizeof(double) * cols);
    cudaMalloc(&out, 8);
    cudaMemset(r, 0, sizeof(double) * m);
    cudaMemset(x, 0, sizeof(double) * cols);
    {
        const int64_t n = cols * L;
        int32_t *h = (int32_t *)malloc(sizeof(int32_t) * (1 << 24));
        uint64_t s = 88172645463325252ull;
        for (int64_t o = 0; o < n; o += (1 << 24)) {
            const int64_t cnt = n - o < (1 << 24) ? n - o : (1 << 24);
            for (int64_t i = 0; i < cnt; ++i) {
                s ^= s << 13;
                s ^= s >> 7;
                s ^= s << 17;
                h[i] = (int32_t)(s % (uint64_t)m);
            }
            cudaMemcpy((int32_t *)idx + o, h, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice);
        }
        free(h);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int per_sm = 4; per_sm <= 8; per_sm += 4) {
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0)
                    probe<0><<<sms * per_sm, 256>>>(r, x, idx, cols, out);
                else
                    probe<1><<<sms * per_sm, 256>>>(r, x, idx, cols, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"G_per_s\": %.1f, \"err\": \"%s\"}\n", mode ? "red_add_f64_scatter" : "ldg_gather",
                   per_sm, best, cols * L / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
