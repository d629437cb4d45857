// Shared helpers for the metalens_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/metalens_b200.h"

namespace mlb {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MLB_ERR_CUDA;
    }
    return MLB_OK;
}

#define MLB_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            mlb::set_error(__VA_ARGS__);       \
            return MLB_ERR_ARG;                \
        }                                      \
    } while (0)

#define MLB_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            mlb::set_error("%s failed: %s", #call, cudaGetErrorString(e_));    \
            return MLB_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

// select one of four kernel-parameter pointers without forcing the parameter struct onto the stack
template <typename T>
__device__ __forceinline__ T pick4(T const (&p)[4], unsigned i) {
    return i == 0 ? p[0] : (i == 1 ? p[1] : (i == 2 ? p[2] : p[3]));
}

// Sum of n doubles in ONE fixed order by a 256-thread block (thread t takes in[t], in[t + 256], ...; xor-shuffle tree per
// warp; the 8 warp sums added in order).  Shared by mlb_sum_f64 and by the kernels that finish total_P themselves, so a
// total does not depend on which of them formed it.  ws: 8 doubles of shared memory.  Result valid in thread 0.
__device__ __forceinline__ double ordered_sum_256(const double *__restrict__ in, int n, double *ws) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v += __ldcg(in + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) s += ws[w];
    }
    return s;
}

// one bit per CUDA device: the "function attribute already set" flags are kept per device, not per process (a process
// that uses several GPUs must raise the dynamic shared-memory limit of a kernel on each of them)
inline unsigned long long device_bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ULL << (dev & 63);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- mbarrier / bulk-copy (1-D TMA) PTX wrappers ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk async copy (SASS: UBLKCP); bytes % 16 == 0, both 16B aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// same with an L2 evict-first policy: for data that is streamed exactly once (the aperture), so that it does not
// push the L2-resident intermediates of concurrently running kernels out of the cache
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace mlb
