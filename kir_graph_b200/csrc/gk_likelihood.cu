// Kernel (a): likelihood build.
//
// Replaces AlleleTyping.reads2AlleleProb + np.log10 (reference:
// graphkir/typing_mulit_allele.py:340-381, :263):
//     m[r, a] = sum over the read's observation entries e of popc(pos_e ^ ((pos_e ^ neg_e) & mem[word_e, a]))
// (= popc((pos & ~mem) | (neg & mem)): a positive observation disagrees with an allele that lacks the
// variant, a negative one with an allele that carries it).
//
// Work decomposition (round 2; the round-1 kernel walked one read at a time and spent ~120 warp
// instructions per read and 32-allele group, 77 % issue-bound at 24 % of HBM):
//   * a CTA takes 128 reads x up to 128 alleles, a warp 16 consecutive reads of them; nothing is staged
//     in shared memory and the warps never wait for each other (the staged version of this kernel
//     issued 40 % fewer instructions than round 1 and was no faster: 49 % issue-active, the time went
//     into the staging phases and their barriers);
//   * a lane owns VW = 1, 2 or 4 CONSECUTIVE alleles and loads their membership words with one 32-,
//     64- or 128-bit load (the row stride of `mem` is padded to a multiple of 32 words for that);
//   * the counts of four consecutive reads (a "quad") live in the four bytes of one register per allele
//     (m <= 255) and the quad's entries are one flat run: an entry is the 16-byte record {byte offset of
//     the membership row, pos, neg, 1 << 8 (r & 3)} (built by gk_expand_reads, or by the host for packs
//     that arrive as entries), so per entry the warp issues one 128-bit load of the record, one
//     vector load of the membership words and per allele LOP3 + POPC + IMAD - against 31 instructions
//     per entry when the loop walked read by read over separate word / pos / neg arrays (ncu: IMAD
//     24 %, LDG 12 %, POPC 6.5 % of the issue slots); the four packed registers of a
//     warp's 16 reads are exactly the 16 bytes of the allele-major LT row, stored with one 128-bit store
//     per allele (no shared-memory transpose); the column sums take one DP4A per quad;
//   * genes with <= 16 (<= 8) alleles put two (four) quads side by side in a warp instead of idling
//     half (three quarters) of the lanes.
// Outputs, written with full lines:
//     L  4 bytes per cell, row-blocked [r_blk][a_blk][32 reads][32] -> operand of the scoring kernel:
//        the 16-bit pair (m, m) for the packed integer path, float32(m) for the FP32 path
//     LT uint8, allele-major [a][r]                                  -> rescoring / P kernels
//     colsum[a] (CN = 1 scores) via one 64-bit atomic per allele per CTA.
//
// Bound: HBM writes, 5 B per cell (4 B L + 1 B LT).
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
#ifndef GK_LIK_UNROLL
#define GK_LIK_UNROLL 2       // entries of a quad in flight per warp (each is a dependent LDS -> LDG -> LOP3 / POPC chain)
#endif
#ifndef GK_LIK_CTAS
#define GK_LIK_CTAS 5         // resident CTAs per SM the register budget is cut for
#endif
constexpr int kEntUnroll = GK_LIK_UNROLL;
constexpr int kReadsPerWarp = GK_LIK_READS / kWarps;   // 16 consecutive reads = 4 quads of four reads
constexpr int kWarpEnt = 96;                           // entries of a warp's 16 reads staged in shared memory
constexpr int kTilePitch = GK_LIK_READS + 16;          // bytes; keeps the rows of the LT tile 16-byte aligned

struct LikShared {
    uint4 ent[kWarps][kWarpEnt];                       // per warp: the entries of its 16 reads (warp-private)
    uint8_t tile[128 * kTilePitch];                    // counts, allele-major, for the coalesced LT stores
    unsigned int col[128];                             // column sums of the tile
};

template <int VW>
__device__ __forceinline__ void load_words(const unsigned char* p, uint32_t (&w)[VW]) {
    if constexpr (VW == 1) {
        w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
    } else if constexpr (VW == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        w[0] = v.x; w[1] = v.y;
    } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
}

// One warp: 16 consecutive reads x (LPQ lanes x VW alleles).  With LPQ < 32 the warp's 32 / LPQ lane
// groups take different quads of the 16 reads.  Nothing is staged and no barrier is needed: an entry
// is one 128-bit load from the same address for all lanes of a group (a broadcast sector), and the
// warps of an SM hide each other's load latency.  The four reads of a quad are ONE flat run of
// entries: entry.w = 1 << 8 (r & 3) routes the popcount into the read's byte of the packed counter.
template <int LPQ, int VW, bool HALF, bool WRITE>
__device__ __forceinline__ void lik_warp(const GkMatrix& M, int r0, int a0, int n_a, const int32_t* __restrict__ eoff,
                                         const unsigned char* __restrict__ mem_bytes,
                                         const uint4* __restrict__ entries, float* __restrict__ L,
                                         LikShared& sh) {
    constexpr int SUBS = 32 / LPQ;                    // lane groups side by side
    constexpr int QPS = 4 / SUBS;                     // quads per lane group
    const int lane = gk_lane();
    const int sub = lane / LPQ;
    const int al = lane % LPQ;
    const int a_rel = al * VW;                        // first allele of this lane, relative to a0
    bool live[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) live[v] = a_rel + v < n_a;
    const bool lane_on = a_rel < n_a;                 // lanes beyond the gene's alleles only keep step
    const unsigned char* mem_lane = mem_bytes + (size_t)(a0 + (lane_on ? a_rel : 0)) * 4;     // (and stay in bounds)
    // keep the lane's base pointer in a register pair: otherwise the compiler rebuilds it from the
    // uniform base and the lane offset for every entry (IMAD.WIDE + IADD3 + IADD3.X per load)
    asm volatile("" : "+l"(mem_lane));

    // entry offsets at the warp's quad boundaries: lane i holds eoff[rw0 + 4 i], clipped at the last read
    const int rw0 = r0 + gk_warp() * kReadsPerWarp;
    int my_off = 0;
    if (lane <= 4) {
        const int r = rw0 + 4 * lane;
        my_off = __ldg(eoff + (r < M.n_reads ? r : M.n_reads));
    }
    // The entries of the warp's 16 reads are one contiguous run: fetch them with coalesced 128-bit loads
    // (all in flight at once: one memory round trip per warp instead of one per entry) into the warp's
    // private staging area; the loop below then reads an entry as one broadcast shared-memory load.
    const int e_first = __shfl_sync(0xffffffffu, my_off, 0);
    const int n_ent = __shfl_sync(0xffffffffu, my_off, 4) - e_first;
    const bool staged = n_ent <= kWarpEnt;
    uint4* s_ent = sh.ent[gk_warp()];
    if (staged) {
#pragma unroll
        for (int i = 0; i < kWarpEnt / 32; ++i)
            if (lane + 32 * i < n_ent) s_ent[lane + 32 * i] = __ldg(entries + e_first + lane + 32 * i);
        __syncwarp();
    }

    uint32_t cnt[QPS][VW];                            // byte j of cnt[q][v] = count of read 4 quad + j
    unsigned int csum[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) csum[v] = 0u;
#pragma unroll
    for (int q = 0; q < QPS; ++q) {
        const int quad = sub * QPS + q;               // quad of the warp's 16 reads taken by this lane group
#pragma unroll
        for (int v = 0; v < VW; ++v) cnt[q][v] = 0u;
        const int e0 = __shfl_sync(0xffffffffu, my_off, quad);
        const int e1 = __shfl_sync(0xffffffffu, my_off, quad + 1);
        auto visit = [&](const uint4 ent) {           // {row byte offset, pos, neg, 1 << 8 (r & 3)}
            uint32_t mw[VW];
            load_words<VW>(mem_lane + ent.x, mw);
#pragma unroll
            for (int v = 0; v < VW; ++v) cnt[q][v] += (uint32_t)__popc((ent.y & ~mw[v]) | (ent.z & mw[v])) * ent.w;
        };
        if (staged) {                                 // shared-memory loads (LDS), not generic ones
            const uint4* ent_ptr = s_ent + (e0 - e_first);
#pragma unroll kEntUnroll
            for (int i = e1 - e0; i > 0; --i, ++ent_ptr) visit(*ent_ptr);
        } else {
            const uint4* ent_ptr = entries + e0;
#pragma unroll 1
            for (int i = e1 - e0; i > 0; --i, ++ent_ptr) visit(__ldg(ent_ptr));
        }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            if (!live[v]) cnt[q][v] = 0u;             // pad columns of the last allele block stay zero
            csum[v] = __dp4a(cnt[q][v], 0x01010101u, csum[v]);
        }
        if constexpr (WRITE) {
            if (lane_on) {
                // L: rows 4 quad .. 4 quad + 3 of the row block, this lane's VW consecutive columns
                const int a = a0 + a_rel;
                uint32_t* slot = reinterpret_cast<uint32_t*>(L) + gk_blk_off(rw0 + 4 * quad, a >> 5, M.n_ablk, 32) + (a & 31);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t out[VW];
#pragma unroll
                    for (int v = 0; v < VW; ++v) {
                        if constexpr (HALF) {
                            out[v] = __byte_perm(cnt[q][v], 0u, 0x4040u | j | (j << 8));    // (m, m) as two 16-bit lanes
                        } else {
                            out[v] = __float_as_uint((float)((cnt[q][v] >> (8 * j)) & 0xffu));
                        }
                    }
                    uint32_t* dst = slot + j * 32;
                    if constexpr (VW == 1) {
                        dst[0] = out[0];
                    } else if constexpr (VW == 2) {
                        *reinterpret_cast<uint2*>(dst) = make_uint2(out[0], out[1]);
                    } else {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
                    }
                }
            }
        }
    }
    if (lane_on) {
        if constexpr (WRITE) {
            // LT: the lane group's 4 QPS consecutive reads of each of the lane's alleles go to the CTA's
            // allele-major byte tile; the CTA stores whole 128-byte rows afterwards
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                uint8_t* dst = sh.tile + (a_rel + v) * kTilePitch + (rw0 - r0) + 4 * QPS * sub;
                if constexpr (QPS == 4) {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(cnt[0][v], cnt[1][v], cnt[2][v], cnt[3][v]);
                } else if constexpr (QPS == 2) {
                    *reinterpret_cast<uint2*>(dst) = make_uint2(cnt[0][v], cnt[1][v]);
                } else {
                    *reinterpret_cast<uint32_t*>(dst) = cnt[0][v];
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VW; ++v)
            if (live[v] && csum[v]) atomicAdd(&sh.col[a_rel + v], csum[v]);
    }
}

__global__ void __launch_bounds__(kThreads, GK_LIK_CTAS)
gk_likelihood_kernel(const GkMatrix* __restrict__ matrices, const GkLikItem* __restrict__ items,
                     const uint32_t* __restrict__ mem_pool, const int32_t* __restrict__ entoff_pool,
                     const uint4* __restrict__ entries, float* __restrict__ L_pool,
                     uint8_t* __restrict__ LT_pool, unsigned long long* __restrict__ col_pool, int half_mode) {
    extern __shared__ __align__(16) unsigned char lik_smem[];
    LikShared& sh = *reinterpret_cast<LikShared*>(lik_smem);

    const GkLikItem item = items[blockIdx.x];
    const GkMatrix M = matrices[item.matrix];
    const bool colsum_only = (item.flags & GK_LIK_COLSUM_ONLY) != 0;
    const int a0 = item.a_blk * M.a_tile;  // a_tile == 32
    const int r0 = item.r0;
    int n_blk = M.n_ablk - item.a_blk;     // allele blocks covered by this CTA (<= 4)
    n_blk = n_blk > 4 ? 4 : n_blk;
    int n_a = M.n_alleles - a0;            // alleles of the gene inside the span
    n_a = n_a > 32 * n_blk ? 32 * n_blk : n_a;
    const unsigned char* mem_bytes = reinterpret_cast<const unsigned char*>(mem_pool + M.mem_off);
    const int32_t* eoff = entoff_pool + M.entoff_off;
    float* L = L_pool + M.L_off;
    uint8_t* LT = LT_pool + M.LT_off;

    if (threadIdx.x < 128) sh.col[threadIdx.x] = 0u;
    __syncthreads();

#define GK_LIK_ARGS M, r0, a0, n_a, eoff, mem_bytes, entries, L, sh
#define GK_LIK_CASE(LPQ, VW)                                             \
    if (colsum_only) lik_warp<LPQ, VW, false, false>(GK_LIK_ARGS);       \
    else if (half_mode) lik_warp<LPQ, VW, true, true>(GK_LIK_ARGS);      \
    else lik_warp<LPQ, VW, false, true>(GK_LIK_ARGS);
    if (n_a <= 8) {
        GK_LIK_CASE(8, 1)
    } else if (n_a <= 16) {
        GK_LIK_CASE(16, 1)
    } else if (n_a <= 32) {
        GK_LIK_CASE(32, 1)
    } else if (n_a <= 64) {
        GK_LIK_CASE(32, 2)
    } else {
        GK_LIK_CASE(32, 4)
    }
#undef GK_LIK_CASE
#undef GK_LIK_ARGS

    if (!colsum_only && n_a < 32 * n_blk) {
        // L columns of the span beyond n_a (pad of the last 32-block) belong to lanes that are off (or
        // already hold zeros): they must read as zero for the scoring kernel
        const int a_lo = a0 + n_a, width = 32 * n_blk - n_a;
        for (int idx = threadIdx.x; idx < width * GK_LIK_READS; idx += kThreads) {
            const int a = a_lo + idx % width;
            const int rl = idx / width;
            reinterpret_cast<uint32_t*>(L)[gk_blk_off(r0 + rl, a >> 5, M.n_ablk, 32) + (a & 31)] = 0u;
        }
    }
    __syncthreads();
    unsigned long long* col = col_pool + M.col_off;
    if (threadIdx.x < n_a && sh.col[threadIdx.x]) atomicAdd(col + a0 + threadIdx.x, (unsigned long long)sh.col[threadIdx.x]);
    if (colsum_only) return;
    for (int idx = threadIdx.x; idx < n_a * (GK_LIK_READS / 16); idx += kThreads) {
        const int a = idx / (GK_LIK_READS / 16);
        const int seg = idx % (GK_LIK_READS / 16);
        const uint4 v = *reinterpret_cast<const uint4*>(sh.tile + a * kTilePitch + seg * 16);
        *reinterpret_cast<uint4*>(LT + (int64_t)(a0 + a) * M.r_pad + r0 + seg * 16) = v;
    }
}

}  // namespace

extern "C" int gk_likelihood(const GkMatrix* matrices, const GkLikItem* items, int n_items,
                             const uint32_t* mem_pool, const int32_t* entoff_pool, const void* entries,
                             float* L_pool, uint8_t* LT_pool, unsigned long long* col_pool, int half_mode,
                             void* stream) {
    if (n_items <= 0) return 0;
    GK_REQUIRE(((uintptr_t)entries & 15) == 0, "gk_likelihood: the entry pool must be 16-byte aligned");
    cudaError_t err = cudaFuncSetAttribute(gk_likelihood_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(LikShared));
    GK_REQUIRE(err == cudaSuccess, "gk_likelihood: cannot reserve %d bytes of shared memory: %s",
               (int)sizeof(LikShared), cudaGetErrorString(err));
    gk_likelihood_kernel<<<n_items, kThreads, sizeof(LikShared), (cudaStream_t)stream>>>(
        matrices, items, mem_pool, entoff_pool, reinterpret_cast<const uint4*>(entries), L_pool, LT_pool, col_pool,
        half_mode);
    GK_CHECK_LAUNCH("gk_likelihood");
    return 0;
}
