// Shared device/host helpers of libgk_typing.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gk_typing.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgk_typing is written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

// ---------------------------------------------------------------------------
// error reporting
// ---------------------------------------------------------------------------
void gk_set_error(const char* fmt, ...);

#define GK_CHECK_LAUNCH(name)                                                         \
    do {                                                                              \
        cudaError_t err__ = cudaGetLastError();                                       \
        if (err__ != cudaSuccess) {                                                   \
            gk_set_error("%s: launch failed: %s", name, cudaGetErrorString(err__));   \
            return -2;                                                                \
        }                                                                             \
    } while (0)

#define GK_REQUIRE(cond, ...)                                                         \
    do {                                                                              \
        if (!(cond)) {                                                                \
            gk_set_error(__VA_ARGS__);                                                \
            return -1;                                                                \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA engine, no tensor map: tiles are contiguous
// in the blocked HBM layouts, so a 1D bulk copy moves a whole stage)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gk_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void gk_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gk_smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void gk_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void gk_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void gk_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gk_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void gk_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gk_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool gk_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(gk_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// non-blocking probe of a phase
__device__ __forceinline__ bool gk_mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(gk_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void gk_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!gk_mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy; completion is signalled on `bar` (complete_tx).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void gk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            gk_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(gk_smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------
__device__ __forceinline__ int gk_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int gk_warp() { return threadIdx.x >> 5; }

__host__ __device__ __forceinline__ int gk_ceil_div(int a, int b) { return (a + b - 1) / b; }

// Element offset of row r in block b of a row-blocked array [r_blk][n_blk][GK_RT][width]
// (the layout of L and P, include/gk_typing.h).
__host__ __device__ __forceinline__ int64_t gk_blk_off(int r, int b, int n_blk, int width) {
    return (((int64_t)(r / GK_RT) * n_blk + b) * GK_RT + (r % GK_RT)) * width;
}

// bits needed to store ids 0..n_alleles (one spare code so that a packed key is never all ones)
__host__ __device__ __forceinline__ int gk_id_bits(int n_alleles) {
    int bits = 1;
    while ((1 << bits) <= n_alleles) ++bits;
    return bits;
}

__device__ __forceinline__ uint32_t gk_hash64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return static_cast<uint32_t>(x);
}
