// Microbenchmark of the scoring inner loop: operands come from shared memory every iteration (as in
// gk_score), accumulators in registers.  Cells/clk/SM for
//   (0) FP32 SAD 8x8 per thread: 128 FADD (FMA pipe) per 64 cells
//   (1) packed-u16 min-sum 8x8: k rows paired, VIMNMX.U16x2 + IADD3 (ALU pipe): 48 instr per 64 cells
//   (2) mixed: 2 rows FP32 SAD (32 FADD) + 6 rows packed u16 (24 VIMNMX + 12 IADD3) per 64 cells
//   (3) packed-u16 min-sum, every add an IMAD (x * one + acc, FMA pipe): 32 VIMNMX + 32 IMAD per 64 cells
//   (4..6) packed-u16 min-sum, 1 of every 2 / 3 / 4 row pairs adds with IADD3, the others with IMAD
//   (7) two reads per 16-bit lane pair, 32-bit sums: VIMNMX.U16x2 + IDP.2A per 2 cells
#include <cstdio>
#include <cuda_runtime.h>

constexpr int RT = 16;

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(const float* gin, float* fout, unsigned* uout, int iters, unsigned one) {
    __shared__ __align__(16) float sPf[RT][16 * 2 + 4];      // 2 fp32 rows per tk (16 tk)
    __shared__ __align__(16) unsigned sPu[RT][16 * 4];       // 4 packed row pairs per tk
    __shared__ __align__(16) float sLf[RT][128];
    __shared__ __align__(16) unsigned sLu[RT][128];
    for (int i = threadIdx.x; i < RT * 128; i += 256) {
        (&sLf[0][0])[i] = gin[i % 4096];
        (&sLu[0][0])[i] = (unsigned)i * 2654435761u & 0x00ff00ffu;
    }
    for (int i = threadIdx.x; i < RT * 64; i += 256) (&sPu[0][0])[i] = (unsigned)i * 40503u & 0x00ff00ffu;
    for (int i = threadIdx.x; i < RT * 36; i += 256) (&sPf[0][0])[i] = gin[i % 4096];
    __syncthreads();
    const int tk = threadIdx.x >> 4, ta = threadIdx.x & 15;
    float acc[64];
    unsigned au[32];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) au[i] = 0u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int r = 0; r < RT; ++r) {
            if (MODE == 0) {
                const float4 p0 = *reinterpret_cast<const float4*>(&sLf[r][tk * 4]);          // stand-in P loads
                const float4 p1 = *reinterpret_cast<const float4*>(&sLf[r][64 + tk * 4]);
                const float4 l0 = *reinterpret_cast<const float4*>(&sLf[(r + 1) % RT][ta * 4]);
                const float4 l1 = *reinterpret_cast<const float4*>(&sLf[(r + 1) % RT][64 + ta * 4]);
                const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i * 8 + j] += fabsf(pv[i] - lv[j]);
            } else if (MODE == 7) {
                const uint4 pa = *reinterpret_cast<const uint4*>(&sLu[r][tk * 4]);          // stand-in P loads
                const uint4 pb = *reinterpret_cast<const uint4*>(&sLu[r][64 + tk * 4]);
                const uint4 l0 = *reinterpret_cast<const uint4*>(&sLu[(r + 1) % RT][ta * 4]);
                const uint4 l1 = *reinterpret_cast<const uint4*>(&sLu[(r + 1) % RT][64 + ta * 4]);
                const unsigned pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                const unsigned lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                unsigned* a32 = reinterpret_cast<unsigned*>(acc);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) a32[i * 8 + j] = __dp2a_lo(__vminu2(pv[i], lv[j]), one, a32[i * 8 + j]);
            } else if (MODE >= 3) {
                const uint4 pu4 = *reinterpret_cast<const uint4*>(&sPu[r][tk * 4]);
                const uint4 lu0 = *reinterpret_cast<const uint4*>(&sLu[r][ta * 4]);
                const uint4 lu1 = *reinterpret_cast<const uint4*>(&sLu[r][64 + ta * 4]);
                const unsigned pu[4] = {pu4.x, pu4.y, pu4.z, pu4.w};
                const unsigned lu[8] = {lu0.x, lu0.y, lu0.z, lu0.w, lu1.x, lu1.y, lu1.z, lu1.w};
                constexpr int EVERY = MODE == 3 ? 1000 : MODE - 2;     // every EVERY-th accumulator adds with IADD3
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const int idx = i * 4 + j / 2;
                        const unsigned x = __vminu2(pu[i], lu[j]), y = __vminu2(pu[i], lu[j + 1]);
                        if (idx % EVERY == EVERY - 1) {
                            au[idx] += x + y;
                        } else {
                            asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(au[idx]) : "r"(x), "r"(one));
                            asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(au[idx]) : "r"(y), "r"(one));
                        }
                    }
            } else {
                const uint4 pu4 = *reinterpret_cast<const uint4*>(&sPu[r][tk * 4]);
                const uint4 lu0 = *reinterpret_cast<const uint4*>(&sLu[r][ta * 4]);
                const uint4 lu1 = *reinterpret_cast<const uint4*>(&sLu[r][64 + ta * 4]);
                const unsigned pu[4] = {pu4.x, pu4.y, pu4.z, pu4.w};
                const unsigned lu[8] = {lu0.x, lu0.y, lu0.z, lu0.w, lu1.x, lu1.y, lu1.z, lu1.w};
                constexpr int NU = MODE == 1 ? 4 : 3;        // packed row pairs handled on the ALU pipe
#pragma unroll
                for (int i = 0; i < NU; ++i)
#pragma unroll
                    for (int j = 0; j < 8; j += 2)
                        au[i * 4 + j / 2] += __vminu2(pu[i], lu[j]) + __vminu2(pu[i], lu[j + 1]);
                if (MODE == 2) {
                    const float2 pf = *reinterpret_cast<const float2*>(&sPf[r][tk * 2]);
                    const float4 l0 = *reinterpret_cast<const float4*>(&sLf[r][ta * 4]);
                    const float4 l1 = *reinterpret_cast<const float4*>(&sLf[r][64 + ta * 4]);
                    const float pv[2] = {pf.x, pf.y};
                    const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i * 8 + j] += fabsf(pv[i] - lv[j]);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    unsigned u = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) u += au[i];
    fout[blockIdx.x * blockDim.x + threadIdx.x] = s;
    uout[blockIdx.x * blockDim.x + threadIdx.x] = u;
}

template <int MODE> void run(const char* name) {
    float *gin, *fout; unsigned* uout;
    cudaMalloc(&gin, 4096 * 4); cudaMalloc(&fout, 296 * 256 * 4); cudaMalloc(&uout, 296 * 256 * 4);
    cudaMemset(gin, 0, 4096 * 4);
    int iters = 1600;
    k<MODE><<<296, 256>>>(gin, fout, uout, 10, MODE == 7 ? 0x0101u : 1u);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<296, 256>>>(gin, fout, uout, iters, MODE == 7 ? 0x0101u : 1u); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    double cells = (MODE == 7 ? 128.0 : 64.0) * RT * iters * 296 * 256;
    printf("%-34s %8.3f ms  %7.2f TCells/s  (%.1f cells/clk/SM at 1.965 GHz) %s\n", name, ms, cells / ms / 1e9,
           cells / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(e));
}

int main() {
    run<0>("fp32 SAD 8x8 (FMA pipe)");
    run<1>("u16x2 min+add 8x8 (ALU pipe)");
    run<2>("mixed 2 rows fp32 + 6 rows u16x2");
    run<3>("u16x2 min (ALU) + IMAD adds (FMA)");
    run<4>("u16x2 min, adds 1/2 IADD3 1/2 IMAD");
    run<5>("u16x2 min, adds 1/3 IADD3 2/3 IMAD");
    run<6>("u16x2 min, adds 1/4 IADD3 3/4 IMAD");
    run<7>("u16x2 min + IDP.2A, 32-bit sums");
    return 0;
}
