// Microbenchmark: cp.async.bulk (1D, UBLKCP) global -> shared throughput per SM as a function of
// the copy size and the number of copies in flight.  One thread issues, everybody waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES>
__global__ void __launch_bounds__(256) k(const char* src, size_t src_bytes, int copy_bytes, int copies_per_stage, int n_tiles,
                                         unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* buf = smem + 128;
    const int stage_bytes = copy_bytes * copies_per_stage;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * stage_bytes;
    size_t off = (size_t)blockIdx.x * stage_bytes;
    unsigned long long issue_clk = 0, n_issue = 0;
    auto issue = [&](int tile, int s) {
        long long c0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage_bytes) : "memory");
        for (int c = 0; c < copies_per_stage; ++c) {
            size_t o = (off + (size_t)tile * stride + (size_t)c * copy_bytes * 977) % (src_bytes - stage_bytes * 1024);
            o &= ~(size_t)127;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             s32(buf + s * stage_bytes + c * copy_bytes)),
                         "l"(src + o), "r"(copy_bytes), "r"(s32(&full[s]))
                         : "memory");
        }
        issue_clk += clock64() - c0; n_issue += copies_per_stage;
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && s < n_tiles; ++s) issue(s, s);
    unsigned long long acc = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % STAGES;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                         : "=r"(ok) : "r"(s32(&full[s])), "r"((t / STAGES) & 1) : "memory");
        acc += buf[s * stage_bytes + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && t + STAGES < n_tiles) issue(t + STAGES, s);
    }
    if (acc == 12345) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) { sink[1] = issue_clk; sink[2] = n_issue; }
}

template <int STAGES> void run(const char* src, size_t n, int copy_bytes, int cps, unsigned long long* sink) {
    int smem = 128 + STAGES * copy_bytes * cps;
    cudaFuncSetAttribute(k<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int tiles = 2000;
    k<STAGES><<<296, 256, smem>>>(src, n, copy_bytes, cps, 50, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<STAGES><<<296, 256, smem>>>(src, n, copy_bytes, cps, tiles, sink); cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double bytes = (double)tiles * 296 * copy_bytes * cps;
    unsigned long long h[3]; cudaMemcpy(h, sink, 24, cudaMemcpyDeviceToHost);
    printf("stages %d  copy %6d B x %d per stage : %8.3f ms  %8.1f GB/s total  %6.1f GB/s per SM  issue %.0f clk/copy (%s)\n", STAGES,
           copy_bytes, cps, ms, bytes / ms / 1e6, bytes / ms / 1e6 / 148, (double)h[1] / (double)h[2], cudaGetErrorString(cudaGetLastError()));
}

int main() {
    size_t n = (size_t)8 << 30;
    char* src; cudaMalloc(&src, n); cudaMemset(src, 1, n);
    unsigned long long* sink; cudaMalloc(&sink, 64);
    run<4>(src, n, 2048, 6, sink);
    run<4>(src, n, 4096, 6, sink);
    run<4>(src, n, 12288, 1, sink);
    run<4>(src, n, 8192, 2, sink);
    run<4>(src, n, 16384, 1, sink);
    run<8>(src, n, 2048, 6, sink);
    run<2>(src, n, 16384, 2, sink);
    run<4>(src, n, 1024, 6, sink);
    run<4>(src, n, 512, 6, sink);
    run<4>(src, n, 128, 6, sink);
    return 0;
}
