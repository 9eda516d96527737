// Shared helpers for the eagle_b200 CUDA kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eagle_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "eagle_b200 kernels are written for sm_100a (B200); compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace egl {

constexpr int kLandmarks = EGL_NUM_LANDMARKS;  // 57
constexpr int kMaxPts = 64;                     // per-frame correspondence capacity (>= 53 on-plane)
constexpr unsigned kFull = 0xffffffffu;

// Channels off the ground plane (cross-bar ends), eagle/utils/pitch.py:65.
constexpr uint64_t kOffPlaneMask = (1ull << 0) | (1ull << 1) | (1ull << 24) | (1ull << 25);

// thread-local last error text, exposed by egl_last_error()
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);  // 0 or -(int)e, recording the text

#define EGL_REQUIRE(cond, code, ...)   \
    do {                               \
        if (!(cond)) {                 \
            egl::set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA engine, 1-D) wrappers.  SASS: SYNCS.* / UBLKCP.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// same, with an L2 evict-first policy: streaming data that is read exactly once should not displace
// anything in the 126 MB L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, m);
    hi = __shfl_xor_sync(kFull, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(kFull, lo, src);
    hi = __shfl_sync(kFull, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor_f64(v, m));
    return v;
}

}  // namespace egl
