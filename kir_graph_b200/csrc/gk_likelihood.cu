// Kernel (a): likelihood build.
//
// Replaces AlleleTyping.reads2AlleleProb + np.log10 (reference:
// graphkir/typing_mulit_allele.py:340-381, :263).  For a tile of 128 reads x up to four
// allele blocks (128 alleles), every lane owns up to four alleles (lane, lane+32, ...) and
// walks the read's packed observation entries:
//     m[r, a] += popc((pos & ~mem[word, a]) | (neg & mem[word, a]))
// The membership row mem[word, :] is word-major, so the 32 lanes of a warp read
// 128 consecutive bytes (coalesced; a gene's table is <= 1 MB and stays in L1/L2).
// Outputs, both written with full 128-byte lines:
//     L  4 bytes per cell, row-blocked [r_blk][a_blk][32 reads][32] -> operand of the scoring kernel (TMA bulk tiles):
//        float32(m) for the FP32 scoring path, the 16-bit pair (m, m) for the packed integer path
//     LT uint8,   allele-major [a][r]         -> rescoring / P kernels stream along reads
// and the per-allele column sums (CN=1 scores) via one 64-bit atomic per allele per CTA.
//
// Bound: HBM writes, 5 B per cell (4 B L + 1 B LT); POPC issue is the secondary limit.
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTilePitch = GK_LIK_READS + 16;  // bytes; keeps rows 16-byte aligned
constexpr int kEntCap = 1024;                  // observation entries of the read tile staged in shared memory

// Per-read work for a CTA whose allele span needs NG lane groups of 32 (uniform per CTA).
// STAGED: every entry of the read tile sits in shared memory (one 128-bit load per entry); the
// rare tile with more than kEntCap entries takes the variant that reads them from global memory.
template <int NG, bool HALF, bool STAGED, bool WRITE>
__device__ __forceinline__ void lik_reads(const GkMatrix& M, int r0, int a0, int a_span, int e_lo,
                                          const uint32_t* __restrict__ mem, const int32_t* __restrict__ ent_word,
                                          const uint32_t* __restrict__ ent_pos, const uint32_t* __restrict__ ent_neg,
                                          const int* s_eoff, const int4* s_ent, float* __restrict__ L,
                                          uint8_t* tile, unsigned int (&csum)[4]) {
    constexpr int kReadsPerWarp = GK_LIK_READS / kWarps;            // consecutive reads per warp
    static_assert(GK_RT % kReadsPerWarp == 0 && kReadsPerWarp % 4 == 0, "a warp's reads lie in one row block of L");
    const int lane = gk_lane();
    const int warp = gk_warp();
    bool live[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) live[g] = (lane + 32 * g < a_span) && (a0 + lane + 32 * g < M.n_alleles);
    const uint32_t* mem_lane = mem + a0 + lane;
    // The reads of a warp are consecutive rows of one row block of L, and group g of a lane is
    // allele block a0 / 32 + g, column `lane`: one pointer that advances by a row per read, with
    // the block as a constant offset.  Four byte counts at a time go to the LT tile as one word.
    const int rl0 = warp * kReadsPerWarp;
    float* slot = L + gk_blk_off(r0 + rl0, a0 >> 5, M.n_ablk, 32) + lane;
#pragma unroll 1
    for (int q = 0; q < kReadsPerWarp; q += 4) {
        uint32_t word[NG];                                          // counts of reads q .. q + 3 as bytes
#pragma unroll
        for (int g = 0; g < NG; ++g) word[g] = 0u;
#pragma unroll 1
        for (int j = 0; j < 4; ++j, slot += 32) {
            const int rl = rl0 + q + j;
            const int r = r0 + rl;
            unsigned int cnt[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) cnt[g] = 0u;
            if (r < M.n_reads) {
                const int e0 = s_eoff[rl] - e_lo;
                const int e1 = s_eoff[rl + 1] - e_lo;
                for (int e = e0; e < e1; ++e) {
                    int w;
                    uint32_t p, n;
                    if constexpr (STAGED) {
                        const int4 ent = s_ent[e];
                        w = ent.x;
                        p = (uint32_t)ent.y;
                        n = (uint32_t)ent.z;
                    } else {
                        w = __ldg(ent_word + e_lo + e);
                        p = __ldg(ent_pos + e_lo + e);
                        n = __ldg(ent_neg + e_lo + e);
                    }
                    const uint32_t* row = mem_lane + (int64_t)w * M.n_alleles;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const uint32_t mw = live[g] ? __ldg(row + 32 * g) : 0u;
                        cnt[g] += __popc((p & ~mw) | (n & mw));
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {                          // lane + 32 g < a_span = 32 NG always
                const unsigned int c = live[g] ? cnt[g] : 0u;
                if constexpr (WRITE) {
                    if constexpr (HALF) {
                        reinterpret_cast<uint32_t*>(slot)[g * (GK_RT * 32)] = c * 0x00010001u;   // (m, m) as two 16-bit lanes
                    } else {
                        slot[g * (GK_RT * 32)] = (float)c;
                    }
                    word[g] |= c << (8 * j);
                }
                csum[g] += c;
            }
        }
        if constexpr (WRITE) {
#pragma unroll
            for (int g = 0; g < NG; ++g)
                *reinterpret_cast<uint32_t*>(tile + (lane + 32 * g) * kTilePitch + rl0 + q) = word[g];
        }
    }
}

__global__ void __launch_bounds__(kThreads)
gk_likelihood_kernel(const GkMatrix* __restrict__ matrices, const GkLikItem* __restrict__ items,
                     const uint32_t* __restrict__ mem_pool, const int32_t* __restrict__ entoff_pool,
                     const int32_t* __restrict__ ent_word, const uint32_t* __restrict__ ent_pos,
                     const uint32_t* __restrict__ ent_neg, float* __restrict__ L_pool,
                     uint8_t* __restrict__ LT_pool, unsigned long long* __restrict__ col_pool, int half_mode) {
    __shared__ __align__(16) uint8_t tile[128 * kTilePitch];
    __shared__ unsigned int colpart[kWarps][128];
    __shared__ int s_eoff[GK_LIK_READS + 1];
    __shared__ __align__(16) int4 s_ent[kEntCap];

    const GkLikItem item = items[blockIdx.x];
    const GkMatrix M = matrices[item.matrix];
    const bool colsum_only = (item.flags & GK_LIK_COLSUM_ONLY) != 0;
    const int a_tile = M.a_tile;           // 32
    const int a0 = item.a_blk * a_tile;
    const int r0 = item.r0;
    int n_blk = M.n_ablk - item.a_blk;     // allele blocks covered by this CTA (<= 4)
    n_blk = n_blk > 4 ? 4 : n_blk;
    const int a_span = n_blk * a_tile;     // <= 128
    const int lane = gk_lane();
    const int warp = gk_warp();

    const uint32_t* mem = mem_pool + M.mem_off;
    const int32_t* eoff = entoff_pool + M.entoff_off;
    // group g of a lane is column (lane + 32 g) of the span = block (a_blk + g) when a_tile == 32
    float* L = L_pool + M.L_off;

    // stage the tile's entry offsets and entries with coalesced loads (they are shared by all lanes)
    for (int i = threadIdx.x; i <= GK_LIK_READS; i += kThreads) {
        const int r = r0 + i;
        s_eoff[i] = __ldg(eoff + (r < M.n_reads ? r : M.n_reads));
    }
    __syncthreads();
    const int e_lo = s_eoff[0];
    const int e_n = s_eoff[GK_LIK_READS] - e_lo;
    const bool staged = e_n <= kEntCap;
    if (staged) {
        for (int i = threadIdx.x; i < e_n; i += kThreads)
            s_ent[i] = make_int4(__ldg(ent_word + e_lo + i), (int)__ldg(ent_pos + e_lo + i),
                                 (int)__ldg(ent_neg + e_lo + i), 0);
    }
    __syncthreads();

    unsigned int csum[4] = {0u, 0u, 0u, 0u};
#define GK_LIK_ARGS M, r0, a0, a_span, e_lo, mem, ent_word, ent_pos, ent_neg, s_eoff, s_ent, L, tile, csum
#define GK_LIK_CASE(NG)                                                    \
    if (colsum_only) {             /* no L / LT: the layout flag is moot */   \
        if (staged) lik_reads<NG, false, true, false>(GK_LIK_ARGS);        \
        else lik_reads<NG, false, false, false>(GK_LIK_ARGS);              \
    } else if (!staged) {                                                  \
        if (half_mode) lik_reads<NG, true, false, true>(GK_LIK_ARGS);      \
        else lik_reads<NG, false, false, true>(GK_LIK_ARGS);               \
    } else if (half_mode) {                                                \
        lik_reads<NG, true, true, true>(GK_LIK_ARGS);                      \
    } else {                                                               \
        lik_reads<NG, false, true, true>(GK_LIK_ARGS);                     \
    }
    switch ((a_span + 31) / 32) {
        case 1: GK_LIK_CASE(1); break;
        case 2: GK_LIK_CASE(2); break;
        case 3: GK_LIK_CASE(3); break;
        default: GK_LIK_CASE(4); break;
    }
#undef GK_LIK_CASE
#undef GK_LIK_ARGS
#pragma unroll
    for (int g = 0; g < 4; ++g) colpart[warp][lane + 32 * g] = csum[g];
    __syncthreads();

    unsigned long long* col = col_pool + M.col_off;
    for (int a = threadIdx.x; a < a_span; a += kThreads) {
        unsigned int s = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += colpart[w][a];
        if (a0 + a < M.n_alleles && s) atomicAdd(col + a0 + a, (unsigned long long)s);
    }

    if (colsum_only) return;
    uint8_t* LT = LT_pool + M.LT_off;
    for (int idx = threadIdx.x; idx < a_span * (GK_LIK_READS / 16); idx += kThreads) {
        const int a = idx / (GK_LIK_READS / 16);
        const int seg = idx % (GK_LIK_READS / 16);
        if (a0 + a < M.n_alleles) {
            const uint4 v = *reinterpret_cast<const uint4*>(tile + a * kTilePitch + seg * 16);
            *reinterpret_cast<uint4*>(LT + (int64_t)(a0 + a) * M.r_pad + r0 + seg * 16) = v;
        }
    }
}

}  // namespace

extern "C" int gk_likelihood(const GkMatrix* matrices, const GkLikItem* items, int n_items,
                             const uint32_t* mem_pool, const int32_t* entoff_pool,
                             const int32_t* ent_word, const uint32_t* ent_pos, const uint32_t* ent_neg,
                             float* L_pool, uint8_t* LT_pool, unsigned long long* col_pool, int half_mode,
                             void* stream) {
    if (n_items <= 0) return 0;
    gk_likelihood_kernel<<<n_items, kThreads, 0, (cudaStream_t)stream>>>(
        matrices, items, mem_pool, entoff_pool, ent_word, ent_pos, ent_neg, L_pool, LT_pool, col_pool, half_mode);
    GK_CHECK_LAUNCH("gk_likelihood");
    return 0;
}
