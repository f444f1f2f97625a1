// common.cuh -- shared device helpers for libbsls_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bsls_b200.h"

namespace bsls {

// ---- per-device host state ---------------------------------------------------------------
// Launch parameters (resident grids, shared-memory opt-ins) are properties of a DEVICE: cudaFuncSetAttribute and the
// occupancy queries apply to the current device only.  Caches are therefore indexed by the current device.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < kMaxDevices) ? d : 0;
}
int num_sms();  // SM count of the current device (B200: 148 = 2 dies x 74); grids are sized in multiples of it (capi.cu)
template <typename V> struct PerDevice {
    V v[kMaxDevices];
    bool seen[kMaxDevices] = {};
    V &get(V initial = V()) {
        const int d = current_device();
        if (!seen[d]) {
            v[d] = initial;
            seen[d] = true;
        }
        return v[d];
    }
};

// ---- error plumbing (host) ---------------------------------------------------------
void set_error(const char *fmt, ...);
#define BSLS_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::bsls::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return BSLS_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)
#define BSLS_LAUNCH_CHECK() BSLS_CUDA_TRY(cudaGetLastError())

// device-resident solver loop: where a kernel finds the step of the trial point and the "solver has stopped" flag
struct StepCtl {
    const double *t;
    const int *done;
};

// ---- small numeric traits -------------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<double> {
    __device__ __forceinline__ static double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }
};
template <> struct Num<float> {
    __device__ __forceinline__ static float neg_inf() { return __int_as_float(0xff800000); }
};

// ---- mbarrier / bulk-copy (TMA 1-D) PTX ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() {
    // make the initialised barrier visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; src/dst 16-byte aligned, bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- per-element asynchronous copies global -> shared (LDGSTS) ---------------------------------
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr(dst_smem)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- streaming global accesses --------------------------------------------------------------
__device__ __forceinline__ void st_stream_v2(double *p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream_v4(float *p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- exact unsigned division by a runtime constant (tile-local indices < 2^20) -------------
struct FastDiv {
    uint32_t d, magic;  // q = (n * magic) >> 32 is exact for n < 2^20, d < 2^12... see make_fastdiv
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.magic = (uint32_t)(((1ull << 32) + d - 1) / d);  // ceil(2^32/d)
    return f;
}
// exact whenever n * d < 2^32 (error term n*(d*magic-2^32) < 2^32); callers keep n <= 2^16, d <= 2^15
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) { return f.d == 1 ? n : __umulhi(n, f.magic); }

}  // namespace bsls
