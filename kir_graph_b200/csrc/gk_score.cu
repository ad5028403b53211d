// Kernel (b): max-then-sum candidate scoring, tiled like a GEMM.
//
// Replaces   np.maximum(log_probs[:, idx], prev.T[:, :, None]).sum(axis=1)
// (reference: graphkir/typing_mulit_allele.py:540-542) without materialising the K x R x A
// temporary.  In mismatch-count form (max of log-probs == min of mismatch counts) one work item
// accumulates, for a tile of kept sets x candidate alleles and a chunk of reads, sum_r min(L, P).
// Reads are the reduction dimension.  No tensor cores: max-then-sum is not a multiply-accumulate.
//
// Two kernels share the tiling and the pipeline:
//   gk_score_packed_kernel (default)  16-bit integer lanes: 2 VIMNMX.U16x2 on the ALU pipe + 2 IMAD (the
//       adds) on the FMA pipe per 4 cells, the min-sum itself goes to S (see score_item_h / score_item_w);
//   gk_score_kernel                   FP32: D[k, a] += sum_r |L[r, a] - P[r, k]| with two FADDs per
//       cell; the consumers recover sum_r min(L, P) = (colsum_L[a] + colsum_P[k] - D[k, a]) / 2
//       (colsum_L = the CN=1 column sums, colsum_P[k] = the previous step's score of set k).  The
//       direct form FMNMX + FADD computes the same thing but FMNMX issues on the half-rate ALU pipe
//       (profiles/r01_score_v0_ncu_summary.txt: ALU 78 %, FMA 39 %).  Operands are float32 holding
//       small integers, exact while a partial sum stays below 2^24 (a chunk is <= 8192 reads x 255).
// Partial sums of the read chunks are merged with 32-bit integer atomics: order independent and
// bit-reproducible.
//
// Data movement: L and P are row-blocked ([r_blk][a_blk][GK_RT][32], [r_blk][k_blk][GK_RT][64]), so the
// GK_RT = 32 reads of the adjacent blocks of a tile are one contiguous span each: a stage is two
// cp.async.bulk copies (TMA engine) signalled on an mbarrier, 3 stages in flight, refilled by
// whichever warp comes by (see score_item).  A full CTA tile is 128 kept sets x 128 alleles, 16 x 16
// threads with an 8 x 8 register tile each; a thread's rows / columns are groups of four
// (k = 4 tk + i, 64 + 4 tk + i) so that a half-warp's 128-bit shared loads hit consecutive banks.
// Ragged K (top_n = 300) and ragged A are covered by smaller tiles instead of computing padding:
// the remainder modes of the FP32 path (granularity 16), the warp-split tiles of the packed path
// (granularity 8).
//
// Bound: non-tensor instruction issue.  FP32 path: 2 FADD per cell on the 128 lanes/clk/SM FMA pipe
// = 148 SM x 64 cells/clk x f_clk; packed path: one issue slot per cell (half an ALU-pipe and half an
// FMA-pipe instruction) = 148 SM x 128 cells/clk x f_clk, 106 cells/clk/SM measured for the mix.
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
#ifndef GK_PACKED_CTAS
#define GK_PACKED_CTAS 3      // resident CTAs per SM of the packed kernel (80 registers per thread)
#endif
constexpr int kStagesF = 3;   // FP32 path
constexpr int kStagesP = GK_PACKED_CTAS == 3 ? 3 : 4;   // packed path
constexpr int kMaxStages = 4;

#ifdef GK_SCORE_PROBE
// Pipeline probe (debug builds only): cycles summed over CTAs of the packed kernel.
// [0] thread 0 waiting for a free stage, [1] thread 0 issuing the bulk copies, [2] thread 0 and
// [3] thread 32 waiting for a full stage, [4] thread 32 computing, [5] stages, [6] CTAs.
__device__ unsigned long long gk_probe[8];
#define GK_PROBE_T(var) const long long var = clock64()
#define GK_PROBE_ADD(slot, val) probe[slot] += (unsigned long long)(val)
#else
#define GK_PROBE_T(var)
#define GK_PROBE_ADD(slot, val)
#endif

// Tile modes per dimension.  F8/F4: 128/64 rows (columns), each thread owns groups of four
// consecutive ones (float4 shared loads).  S1..S3: the ragged remainder, 16/32/48 rows
// (columns): thread t owns row t of each 16-group (scalar shared loads).
enum Mode { F8 = 0, F4 = 1, S1 = 2, S2 = 3, S3 = 4 };

template <int M> struct ModeInfo {
    static constexpr bool kVec = (M == F8 || M == F4);
    static constexpr int kPerThread = M == F8 ? 8 : M == F4 ? 4 : (M - S1 + 1);   // rows (columns) per thread
    static constexpr int kSpan = 16 * kPerThread;                                  // rows (columns) per CTA
};

template <int KM, int AM>
__device__ __forceinline__ void score_item(const GkScoreItem& item, const GkMatrix& M, const GkSearch& X,
                                           const float* __restrict__ L_pool, const float* __restrict__ P_pool,
                                           uint32_t* __restrict__ S_pool, float* smem, uint64_t* full,
                                           uint64_t* empty, int* next) {
    constexpr int TK = ModeInfo<KM>::kPerThread;
    constexpr int TA = ModeInfo<AM>::kPerThread;
    constexpr int AT = 32;                                           // layout block width of L
    constexpr int KW = (ModeInfo<KM>::kSpan + GK_KB - 1) / GK_KB;    // k-blocks staged (1 or 2)
    constexpr int AW = (ModeInfo<AM>::kSpan + AT - 1) / AT;          // a-blocks staged (1, 2 or 4)
    constexpr uint32_t kBytesPBlk = GK_RT * GK_KB * sizeof(float);
    constexpr uint32_t kBytesLBlk = GK_RT * AT * sizeof(float);
    constexpr uint32_t kStageBytes = KW * kBytesPBlk + AW * kBytesLBlk;
    constexpr int kStageFloats = GK_RT * (KW * GK_KB + AW * AT);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tk = tid >> 4;
    const int ta = tid & 15;

    // L and P are row-blocked ([r_blk][block][GK_RT][width]): the GK_RT rows of the KW (AW)
    // adjacent blocks of a tile are contiguous, so a stage is two bulk copies.  Issuing one
    // costs the issuing thread ~500 cycles whatever its size (tools/micro/bulkcopy.cu), which
    // is why the layout keeps their number down.
    const int rb0 = item.r0 / GK_RT;
    const float* gP = P_pool + X.P_off + ((int64_t)rb0 * X.n_kblk + item.k_blk) * (GK_RT * GK_KB);
    const float* gL = L_pool + M.L_off + ((int64_t)rb0 * M.n_ablk + item.a_blk) * (GK_RT * AT);
    const int64_t rb_stride_p = (int64_t)X.n_kblk * (GK_RT * GK_KB);
    const int64_t rb_stride_l = (int64_t)M.n_ablk * (GK_RT * AT);
    const int n_tiles = (item.r1 - item.r0) / GK_RT;

    auto issue = [&](int tile, int s) {
        float* dst = smem + s * kStageFloats;
        gk_mbar_arrive_expect_tx(&full[s], kStageBytes);
        gk_bulk_g2s(dst, gP + tile * rb_stride_p, KW * kBytesPBlk, &full[s]);
        gk_bulk_g2s(dst + KW * GK_RT * GK_KB, gL + tile * rb_stride_l, AW * kBytesLBlk, &full[s]);
    };

    // Refills are not tied to one thread: `next` is the next tile to fetch, and lane 0 of any warp
    // that comes by (at the top of a stage, or each time it wakes up while waiting for data) claims
    // it once every warp has released the stage it goes to.
    if (tid == 0) {
        const int pre = n_tiles < kStagesF ? n_tiles : kStagesF;
        for (int s = 0; s < pre; ++s) issue(s, s);
        *next = pre;
    }
    __syncthreads();
    auto try_refill = [&]() {
        const int n = *reinterpret_cast<volatile int*>(next);
        if (n >= n_tiles) return;
        const int sp = n % kStagesF;
        if (!gk_mbar_test(&empty[sp], (n / kStagesF - 1) & 1)) return;
        if (atomicCAS(next, n, n + 1) == n) issue(n, sp);
    };

    float acc[TK][TA];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < TA; ++j) acc[i][j] = 0.f;

    // Row / column owned by slot i of thread t, relative to the tile origin.
    auto k_of = [&](int i) { return ModeInfo<KM>::kVec ? (i < 4 ? tk * 4 + i : 64 + tk * 4 + (i - 4)) : i * 16 + tk; };
    auto a_of = [&](int j) { return ModeInfo<AM>::kVec ? (j < 4 ? ta * 4 + j : 64 + ta * 4 + (j - 4)) : j * 16 + ta; };
    // shared-memory offset of column c of the staged L blocks ([a_blk][r][AT])
    auto l_off = [&](int c) { return (c / AT) * (GK_RT * AT) + (c % AT); };

#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kStagesF;
        // Warp-uniform wait (the exit is voted on).  A polling loop run by one lane only leaves that
        // lane diverged from the other 31 through the compute loop, which doubles the instructions
        // the warp issues.  try_wait suspends the thread for a bounded time, so a waiting warp costs
        // no issue slots and looks for a refill to claim every time it wakes up.
        if (lane == 0) try_refill();
        __syncwarp();
        while (true) {
            const bool ok = gk_mbar_try_wait(&full[s], (t / kStagesF) & 1);
            if (__all_sync(0xffffffffu, ok)) break;
            if (lane == 0) try_refill();
            __syncwarp();
        }

        const float* p = smem + s * kStageFloats;                       // [k_blk][r][GK_KB]
        const float* l = smem + s * kStageFloats + KW * GK_RT * GK_KB;  // [a_blk][r][AT]
#pragma unroll 4
        for (int r = 0; r < GK_RT; ++r) {
            float pv[TK];
            float lv[TA];
            if constexpr (ModeInfo<KM>::kVec) {
                const float4 x = *reinterpret_cast<const float4*>(p + r * GK_KB + tk * 4);
                pv[0] = x.x; pv[1] = x.y; pv[2] = x.z; pv[3] = x.w;
                if constexpr (TK == 8) {
                    const float4 y = *reinterpret_cast<const float4*>(p + GK_RT * GK_KB + r * GK_KB + tk * 4);
                    pv[4] = y.x; pv[5] = y.y; pv[6] = y.z; pv[7] = y.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TK; ++i) pv[i] = p[r * GK_KB + i * 16 + tk];
            }
            if constexpr (ModeInfo<AM>::kVec) {
                const float4 x = *reinterpret_cast<const float4*>(l + l_off(ta * 4) + r * AT);
                lv[0] = x.x; lv[1] = x.y; lv[2] = x.z; lv[3] = x.w;
                if constexpr (TA == 8) {
                    const float4 y = *reinterpret_cast<const float4*>(l + l_off(64 + ta * 4) + r * AT);
                    lv[4] = y.x; lv[5] = y.y; lv[6] = y.z; lv[7] = y.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TA; ++j) lv[j] = l[l_off(j * 16 + ta) + r * AT];
            }
#pragma unroll
            for (int i = 0; i < TK; ++i)
#pragma unroll
                for (int j = 0; j < TA; ++j) acc[i][j] += fabsf(pv[i] - lv[j]);
        }
        __syncwarp();
        if (lane == 0) gk_mbar_arrive(&empty[s]);
    }

    uint32_t* S = S_pool + X.S_off;
    const int k_base = item.k_blk * GK_KB;
    const int a_base = item.a_blk * AT;
#pragma unroll
    for (int i = 0; i < TK; ++i) {
        const int k = k_base + k_of(i);
#pragma unroll
        for (int j = 0; j < TA; ++j) {
            const uint32_t v = (uint32_t)acc[i][j];
            if (v) atomicAdd(S + (int64_t)k * X.s_stride + a_base + a_of(j), v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Packed 16-bit integer path: one instruction per cell instead of 2, split over the ALU and FMA pipes.
//
// P is stored as uint16 [k_blk][r][64], L as the pair (m, m) in one 32-bit word [a_blk][r][32].
// A thread owns row pairs (2 tk, 2 tk + 1) of every 32-row group g < G of the tile, so one 32-bit
// word of P holds two rows and pairs with the duplicated L value:
//     acc2 += min.u16x2(p2(r), l2(r))
// = 1 VIMNMX.U16x2 + 1 add for 2 cells.  (Until the middle of round 2 two reads were folded into one
// 3-input IADD3: 0.75 instructions per cell, but all of them on the half-rate ALU pipe -
// tools/micro/mixpipe.cu measures 83 cells/clk/SM for that, 55 for the two FADDs of the FP32 path; see
// add_fma below for what replaced it.)  Everything is integer, so
// it is exact for every supported input (counts <= 255).  A 16-bit lane holds 65535 / m_max reads
// (m_max = the largest count of the batch, <= 255), so every `flush` stages (the host derives it
// from m_max) the packed sums are added to S with integer atomics; there are no 32-bit
// accumulator registers.  S receives the min-sum itself (not the sum of absolute differences of
// the FP32 path).
// Tile rows: G in {1..4} groups of 32 kept sets; columns: the same five modes as above.
//
// The adds do not go to the ALU pipe: acc = x * one + acc with `one` a kernel argument (so that ptxas
// cannot fold it back into an IADD3) is an IMAD, which issues on the FMA pipe.  The minima (ALU pipe,
// 2 clk per warp instruction) and the adds (FMA pipe) then overlap: 2 VIMNMX.U16x2 + 2 IMAD per 4 cells
// is one issue slot per cell and 0.5 ALU-pipe instructions per cell.  tools/micro/mixpipe.cu measures
// 106 cells/clk/SM for this mix against 83 for 2 VIMNMX + 1 IADD3 (profiles/r02_mixpipe.txt); mixes of
// IADD3 and IMAD adds land in between.

// N consecutive 32-bit words of shared memory with the widest load their alignment allows (the callers
// guarantee 8-byte alignment for N = 2 and 16-byte alignment for N = 4).
template <int N>
__device__ __forceinline__ void lds_words(const uint32_t* src, uint32_t* out) {
    if constexpr (N == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(src);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else if constexpr (N == 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(src);
        out[0] = v.x; out[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = src[i];
    }
}

__device__ __forceinline__ uint32_t add_fma(uint32_t acc, uint32_t x, uint32_t one) {
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(one));
    return acc;
}
template <int G, int AM>
__device__ __forceinline__ void score_item_h(const GkScoreItem& item, const GkMatrix& M, const GkSearch& X,
                                             const float* __restrict__ L_pool, const uint16_t* __restrict__ P_pool,
                                             uint32_t* __restrict__ S_pool, unsigned char* smem_bytes,
                                             uint64_t* full, uint64_t* empty, int* next, int flush, uint32_t one) {
    constexpr int TA = ModeInfo<AM>::kPerThread;
    constexpr int AT = 32;
    constexpr int KW = (32 * G + GK_KB - 1) / GK_KB;                 // k-blocks staged (1 or 2)
    constexpr int AW = (ModeInfo<AM>::kSpan + AT - 1) / AT;          // a-blocks staged (1, 2 or 4)
    constexpr uint32_t kBytesPBlk = GK_RT * GK_KB * sizeof(uint16_t);
    constexpr uint32_t kBytesLBlk = GK_RT * AT * sizeof(uint32_t);
    constexpr uint32_t kStageBytes = KW * kBytesPBlk + AW * kBytesLBlk;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tk = tid >> 4;
    const int ta = tid & 15;

    // row-blocked L and P: two bulk copies per stage (see score_item)
    const int rb0 = item.r0 / GK_RT;
    const uint16_t* gP = P_pool + X.P_off + ((int64_t)rb0 * X.n_kblk + item.k_blk) * (GK_RT * GK_KB);
    const uint32_t* gL = reinterpret_cast<const uint32_t*>(L_pool) + M.L_off +
                         ((int64_t)rb0 * M.n_ablk + item.a_blk) * (GK_RT * AT);
    const int64_t rb_stride_p = (int64_t)X.n_kblk * (GK_RT * GK_KB);
    const int64_t rb_stride_l = (int64_t)M.n_ablk * (GK_RT * AT);
    const int n_tiles = (item.r1 - item.r0) / GK_RT;

    auto issue = [&](int tile, int s) {
        unsigned char* dst = smem_bytes + (size_t)s * kStageBytes;
        gk_mbar_arrive_expect_tx(&full[s], kStageBytes);
        gk_bulk_g2s(dst, gP + tile * rb_stride_p, KW * kBytesPBlk, &full[s]);
        gk_bulk_g2s(dst + KW * kBytesPBlk, gL + tile * rb_stride_l, AW * kBytesLBlk, &full[s]);
    };

    // opportunistic refill, as in score_item
    if (tid == 0) {
        const int pre = n_tiles < kStagesP ? n_tiles : kStagesP;
        for (int s = 0; s < pre; ++s) issue(s, s);
        *next = pre;
    }
    __syncthreads();
    auto try_refill = [&]() {
        const int n = *reinterpret_cast<volatile int*>(next);
        if (n >= n_tiles) return;
        const int sp = n % kStagesP;
        if (!gk_mbar_test(&empty[sp], (n / kStagesP - 1) & 1)) return;
        if (atomicCAS(next, n, n + 1) == n) issue(n, sp);
    };

    uint32_t acc2[G][TA];      // two 16-bit sums per word: rows 32 g + 2 tk and 32 g + 2 tk + 1
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < TA; ++j) acc2[g][j] = 0u;

    auto a_of = [&](int j) { return ModeInfo<AM>::kVec ? (j < 4 ? ta * 4 + j : 64 + ta * 4 + (j - 4)) : j * 16 + ta; };
    auto l_off = [&](int c) { return (c / AT) * (GK_RT * AT) + (c % AT); };
    // rows 32 g + 2 tk + {0, 1}: offset (in uint16) inside the staged P blocks [k_blk][r][64]
    auto p_off = [&](int g) { return ((32 * g) / GK_KB) * (GK_RT * GK_KB) + ((32 * g) % GK_KB) + 2 * tk; };

    // The packed sums go straight to S (no 32-bit accumulator registers: the 32 words of acc2 are
    // the only accumulators, which leaves the scheduler room to keep the ALU pipe fed).
    uint32_t* S = S_pool + X.S_off + (int64_t)(item.k_blk * GK_KB + 2 * tk) * X.s_stride + item.a_blk * AT;
    auto flush_acc = [&]() {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < TA; ++j) {
                const uint32_t v = acc2[g][j];
                uint32_t* cell = S + (int64_t)(32 * g) * X.s_stride + a_of(j);
                if (v & 0xffffu) atomicAdd(cell, v & 0xffffu);
                if (v >> 16) atomicAdd(cell + X.s_stride, v >> 16);
                acc2[g][j] = 0u;
            }
    };

    auto load_row = [&](const uint16_t* p, const uint32_t* l, int r, uint32_t (&pv)[G], uint32_t (&lv)[TA]) {
#pragma unroll
        for (int g = 0; g < G; ++g) pv[g] = *reinterpret_cast<const uint32_t*>(p + p_off(g) + r * GK_KB);
        if constexpr (ModeInfo<AM>::kVec) {
            const uint4 x = *reinterpret_cast<const uint4*>(l + l_off(ta * 4) + r * AT);
            lv[0] = x.x; lv[1] = x.y; lv[2] = x.z; lv[3] = x.w;
            if constexpr (TA == 8) {
                const uint4 y = *reinterpret_cast<const uint4*>(l + l_off(64 + ta * 4) + r * AT);
                lv[4] = y.x; lv[5] = y.y; lv[6] = y.z; lv[7] = y.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < TA; ++j) lv[j] = l[l_off(j * 16 + ta) + r * AT];
        }
    };

    int since_flush = 0;
#ifdef GK_SCORE_PROBE
    unsigned long long probe[6] = {0, 0, 0, 0, 0, 0};
#endif
#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kStagesP;
        // Warp-uniform wait (the exit is voted on): a lane-0-only polling loop leaves lane 0 diverged
        // from the other 31 lanes through the compute loop, doubling the instructions issued.
        GK_PROBE_T(c0);
        if (lane == 0) try_refill();
        __syncwarp();
        while (true) {
            // try_wait suspends the thread for a bounded time: a waiting warp costs no issue slots
            // and looks for a refill to claim every time it wakes up
            const bool ok = gk_mbar_try_wait(&full[s], (t / kStagesP) & 1);
            if (__all_sync(0xffffffffu, ok)) break;
            if (lane == 0) try_refill();
            __syncwarp();
        }
        GK_PROBE_T(c1);
        GK_PROBE_ADD(0, c1 - c0);
        GK_PROBE_T(c3);
        GK_PROBE_T(c4);
        GK_PROBE_ADD(tid == 0 ? 2 : 3, c4 - c3);

        const uint16_t* p = reinterpret_cast<const uint16_t*>(smem_bytes + (size_t)s * kStageBytes);
        const uint32_t* l = reinterpret_cast<const uint32_t*>(smem_bytes + (size_t)s * kStageBytes + KW * kBytesPBlk);
#pragma unroll 2
        for (int r = 0; r < GK_RT; r += 2) {
            uint32_t pv0[G], pv1[G], lv0[TA], lv1[TA];
            load_row(p, l, r, pv0, lv0);
            load_row(p, l, r + 1, pv1, lv1);
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int j = 0; j < TA; ++j)
                    acc2[g][j] = add_fma(add_fma(acc2[g][j], __vminu2(pv0[g], lv0[j]), one), __vminu2(pv1[g], lv1[j]), one);
        }
        if (++since_flush >= flush) {
            flush_acc();
            since_flush = 0;
        }
        __syncwarp();
        if (lane == 0) gk_mbar_arrive(&empty[s]);
        GK_PROBE_T(c5);
        GK_PROBE_ADD(4, c5 - c4);
    }
    flush_acc();
#ifdef GK_SCORE_PROBE
    if (tid == 0) {
        atomicAdd(&gk_probe[0], probe[0]);
        atomicAdd(&gk_probe[1], probe[1]);
        atomicAdd(&gk_probe[2], probe[2]);
        atomicAdd(&gk_probe[5], (unsigned long long)n_tiles);
        atomicAdd(&gk_probe[6], 1ull);
    }
    if (tid == 32) {
        atomicAdd(&gk_probe[3], probe[3]);
        atomicAdd(&gk_probe[4], probe[4]);
    }
#endif

}

// ---------------------------------------------------------------------------------------
// Warp-split tiles of the packed path (item.shape & GK_SHAPE_WARP_SPLIT): small problems.
//
// A gene with few alleles (or the ragged right edge of a wide one) would leave most of a
// 128 x 128 CTA tile empty, and a CTA tile cut down to 32 x 32 leaves every thread 2 x 2 cells,
// i.e. more shared loads than minima.  Here the tile is 8 TA' columns (TA' = 1..8) by
// WK x 8 G' rows (G' = 1..4 groups of 8, WK = 1, 2 or 4 warps side by side), a lane still owns
// 2 G' rows x TA' columns (lane = 4 row pairs x 8 columns), and the 8 / WK warps that share
// the rows split the 32 reads of every stage between them.  Granularity 8 in both dimensions,
// same instruction mix as the full tile.  Partial sums meet in S through the atomics anyway.
// The tile may start at any multiple of 8 rows inside its k-block (shape bits 20-22), so that a
// ragged kept-set count is covered exactly (300 = 128 + 128 + 32 + 8 + ... instead of padding the
// last tile to 32- or 64-row groups).
//
// Two lane layouts.  VEC = false: rows row0 + 8 g + 2 tk, columns 8 j + ta, scalar shared loads; the eight
// lanes that share a row pair flush eight consecutive words of S (one 32-byte sector per atomic
// instruction and row).  VEC = true (tiles of WK = 4, where a warp has 16 reads of every stage): a lane owns
// 2 G' consecutive rows and its TA' columns in at most two runs of consecutive ones, so its operands of a
// read arrive with vector loads (G' = 4: one LDS.128 for P; TA' = 5: LDS.128 + LDS.32 for L) - the kernel
// is bound by instruction issue, and this takes 18 loads per 80 arithmetic instructions down to 6.  Its
// flush touches up to four sectors per instruction and row, which is why the tiles of fewer warps per
// row block (fewer cells per flush) keep the first layout: measured on the shapes of a cfg3 sample, the
// vector layout is 6-10 % faster on WK = 4 tiles and 7-18 % slower on WK <= 2 tiles.
template <int GP, int TAP, bool VEC>
__device__ __forceinline__ void score_item_w(const GkScoreItem& item, const GkMatrix& M, const GkSearch& X,
                                             const float* __restrict__ L_pool, const uint16_t* __restrict__ P_pool,
                                             uint32_t* __restrict__ S_pool, unsigned char* smem_bytes,
                                             uint64_t* full, uint64_t* empty, int* next, int flush, int wk_log2, uint32_t one) {
    constexpr int AT = 32;
    constexpr int AW = (8 * TAP + AT - 1) / AT;                      // a-blocks staged (1 or 2)
    constexpr uint32_t kBytesPBlk = GK_RT * GK_KB * sizeof(uint16_t);
    constexpr uint32_t kBytesLBlk = GK_RT * AT * sizeof(uint32_t);
    constexpr uint32_t kStageStride = 2 * kBytesPBlk + 2 * kBytesLBlk;
    const int row_off = ((item.shape >> 20) & 7) * 8;                // first row of the tile inside its k-block
    const int KW = (row_off + ((8 * GP) << wk_log2)) > GK_KB ? 2 : 1;            // k-blocks staged
    const uint32_t stage_bytes = KW * kBytesPBlk + AW * kBytesLBlk;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tk = lane >> 3;                                        // row pair inside a group of 8 rows
    const int ta = lane & 7;                                         // column inside a group of 8 columns
    if constexpr (VEC) wk_log2 = 2;                                  // compile-time trip counts
    const int row0 = row_off + (warp & ((1 << wk_log2) - 1)) * (8 * GP);       // first row of this warp
    const int reads_per_warp = 4 << wk_log2;                         // 32 reads / (8 >> wk_log2) warps
    const int rd0 = (warp >> wk_log2) * reads_per_warp;

    const int rb0 = item.r0 / GK_RT;
    const uint16_t* gP = P_pool + X.P_off + ((int64_t)rb0 * X.n_kblk + item.k_blk) * (GK_RT * GK_KB);
    const uint32_t* gL = reinterpret_cast<const uint32_t*>(L_pool) + M.L_off +
                         ((int64_t)rb0 * M.n_ablk + item.a_blk) * (GK_RT * AT);
    const int64_t rb_stride_p = (int64_t)X.n_kblk * (GK_RT * GK_KB);
    const int64_t rb_stride_l = (int64_t)M.n_ablk * (GK_RT * AT);
    const int n_tiles = (item.r1 - item.r0) / GK_RT;

    auto issue = [&](int tile, int s) {
        unsigned char* dst = smem_bytes + (size_t)s * kStageStride;
        gk_mbar_arrive_expect_tx(&full[s], stage_bytes);
        gk_bulk_g2s(dst, gP + tile * rb_stride_p, KW * kBytesPBlk, &full[s]);
        gk_bulk_g2s(dst + 2 * kBytesPBlk, gL + tile * rb_stride_l, AW * kBytesLBlk, &full[s]);
    };
    if (tid == 0) {
        const int pre = n_tiles < kStagesP ? n_tiles : kStagesP;
        for (int s = 0; s < pre; ++s) issue(s, s);
        *next = pre;
    }
    __syncthreads();
    auto try_refill = [&]() {
        const int n = *reinterpret_cast<volatile int*>(next);
        if (n >= n_tiles) return;
        const int sp = n % kStagesP;
        if (!gk_mbar_test(&empty[sp], (n / kStagesP - 1) & 1)) return;
        if (atomicCAS(next, n, n + 1) == n) issue(n, sp);
    };

    uint32_t acc2[GP][TAP];
#pragma unroll
    for (int g = 0; g < GP; ++g)
#pragma unroll
        for (int j = 0; j < TAP; ++j) acc2[g][j] = 0u;

    // row (relative to the k-block) and column (relative to the first a-block) of slot (g, j) of this lane
    constexpr int TA0 = TAP < 4 ? TAP : 4;        // VEC: columns TA0 ta + j (first a-block) ...
    constexpr int TA1 = TAP - TA0;                //      ... and 32 + TA1 ta + (j - TA0) (second a-block)
    auto row_of = [&](int g) { return VEC ? row0 + 2 * GP * tk + 2 * g : row0 + 8 * g + 2 * tk; };
    auto col_of = [&](int j) { return VEC ? (j < TA0 ? TA0 * ta + j : AT + TA1 * ta + (j - TA0)) : 8 * j + ta; };
    // offset (in uint16) inside the staged P blocks [k_blk][r][64]; a lane's rows never straddle a k-block
    // when 2 G' divides 64
    auto p_off = [&](int g) { return (row_of(g) / GK_KB) * (GK_RT * GK_KB) + (row_of(g) % GK_KB); };
    // offset (in words) inside the staged L blocks [a_blk][r][32]
    auto l_off = [&](int j) { return (col_of(j) / AT) * (GK_RT * AT) + (col_of(j) % AT); };

    // slot (0, 0) of this lane and, for the vector layout, the first slot of its second run of columns;
    // the other slots are compile-time offsets from these
    uint32_t* S = S_pool + X.S_off + (int64_t)(item.k_blk * GK_KB + row_of(0)) * X.s_stride + item.a_blk * AT + col_of(0);
    uint32_t* S1 = S + (col_of(TA0 < TAP ? TA0 : 0) - col_of(0));
    auto flush_acc = [&]() {
#pragma unroll
        for (int g = 0; g < GP; ++g)
#pragma unroll
            for (int j = 0; j < TAP; ++j) {
                const uint32_t v = acc2[g][j];
                uint32_t* cell = (VEC ? (j < TA0 ? S + j : S1 + (j - TA0)) : S + 8 * j) +
                                 (int64_t)(VEC ? 2 * g : 8 * g) * X.s_stride;
                if (v & 0xffffu) atomicAdd(cell, v & 0xffffu);
                if (v >> 16) atomicAdd(cell + X.s_stride, v >> 16);
                acc2[g][j] = 0u;
            }
    };

    int since_flush = 0;
#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kStagesP;
        if (lane == 0) try_refill();
        __syncwarp();
        while (true) {                       // warp-uniform wait, see score_item
            const bool ok = gk_mbar_try_wait(&full[s], (t / kStagesP) & 1);
            if (__all_sync(0xffffffffu, ok)) break;
            if (lane == 0) try_refill();
            __syncwarp();
        }
        const uint16_t* p = reinterpret_cast<const uint16_t*>(smem_bytes + (size_t)s * kStageStride);
        const uint32_t* l = reinterpret_cast<const uint32_t*>(smem_bytes + (size_t)s * kStageStride + 2 * kBytesPBlk);
#pragma unroll 2
        for (int i = 0; i < reads_per_warp; i += 2) {
            const int r = rd0 + i;
            uint32_t pv0[GP], pv1[GP], lv0[TAP], lv1[TAP];
            if constexpr (VEC && (GP == 4 || GP == 2)) {
                lds_words<GP>(reinterpret_cast<const uint32_t*>(p + p_off(0) + r * GK_KB), pv0);
                lds_words<GP>(reinterpret_cast<const uint32_t*>(p + p_off(0) + (r + 1) * GK_KB), pv1);
            } else {
#pragma unroll
                for (int g = 0; g < GP; ++g) {
                    pv0[g] = *reinterpret_cast<const uint32_t*>(p + p_off(g) + r * GK_KB);
                    pv1[g] = *reinterpret_cast<const uint32_t*>(p + p_off(g) + (r + 1) * GK_KB);
                }
            }
            if constexpr (VEC) {
                lds_words<TA0>(l + l_off(0) + r * AT, lv0);
                lds_words<TA0>(l + l_off(0) + (r + 1) * AT, lv1);
                if constexpr (TA1 > 0) {
                    lds_words<TA1>(l + l_off(TA0) + r * AT, lv0 + TA0);
                    lds_words<TA1>(l + l_off(TA0) + (r + 1) * AT, lv1 + TA0);
                }
            } else {
#pragma unroll
                for (int j = 0; j < TAP; ++j) {
                    lv0[j] = l[l_off(j) + r * AT];
                    lv1[j] = l[l_off(j) + (r + 1) * AT];
                }
            }
#pragma unroll
            for (int g = 0; g < GP; ++g)
#pragma unroll
                for (int j = 0; j < TAP; ++j)
                    acc2[g][j] = add_fma(add_fma(acc2[g][j], __vminu2(pv0[g], lv0[j]), one), __vminu2(pv1[g], lv1[j]), one);
        }
        if (++since_flush >= flush) {
            flush_acc();
            since_flush = 0;
        }
        __syncwarp();
        if (lane == 0) gk_mbar_arrive(&empty[s]);
    }
    flush_acc();
}

template <int GP>
__device__ __forceinline__ void score_dispatch_w(int tap, const GkScoreItem& item, const GkMatrix& M,
                                                 const GkSearch& X, const float* __restrict__ L_pool,
                                                 const uint16_t* __restrict__ P_pool, uint32_t* __restrict__ S_pool,
                                                 unsigned char* smem, uint64_t* full, uint64_t* empty, int* next,
                                                 int flush, int wk_log2, uint32_t one) {
#define GK_W_CASE(T)                                                                                                        \
    case T:                                                                                                                 \
        if (wk_log2 == 2)                                                                                                   \
            score_item_w<GP, T, true>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, wk_log2, one);   \
        else                                                                                                                \
            score_item_w<GP, T, false>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, wk_log2, one);  \
        break;
    switch (tap) {
        GK_W_CASE(1) GK_W_CASE(2) GK_W_CASE(3) GK_W_CASE(4) GK_W_CASE(5) GK_W_CASE(6) GK_W_CASE(7)
        default: GK_W_CASE(8)
    }
#undef GK_W_CASE
}

__global__ void __launch_bounds__(kThreads, GK_PACKED_CTAS)
gk_score_packed_kernel(const GkMatrix* __restrict__ matrices, const GkSearch* __restrict__ searches,
                     const GkScoreItem* __restrict__ items, const float* __restrict__ L_pool,
                     const uint16_t* __restrict__ P_pool, uint32_t* __restrict__ S_pool, int flush,
                     const int32_t* __restrict__ kept_count, uint32_t one) {
    if (kept_count != nullptr && items[blockIdx.x].k_blk * GK_KB >= kept_count[items[blockIdx.x].search]) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    int* next = reinterpret_cast<int*>(smem_raw + 64);      // next tile to fetch (see score_item_h)
    unsigned char* smem = smem_raw + 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesP; ++s) {
            gk_mbar_init(&full[s], 1);
            gk_mbar_init(&empty[s], kWarps);
        }
        gk_fence_barrier_init();
    }
    __syncthreads();

    const GkScoreItem item = items[blockIdx.x];
    const GkSearch X = searches[item.search];
    const GkMatrix M = matrices[X.matrix];
    if (item.shape & GK_SHAPE_WARP_SPLIT) {      // G' | log2(WK) << 4 | TA' << 8
        const int tap = (item.shape >> 8) & 0xff;
        const int wk_log2 = (item.shape >> 4) & 0xf;
        switch (item.shape & 0xf) {
            case 1: score_dispatch_w<1>(tap, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, wk_log2, one); break;
            case 2: score_dispatch_w<2>(tap, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, wk_log2, one); break;
            case 3: score_dispatch_w<3>(tap, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, wk_log2, one); break;
            default: score_dispatch_w<4>(tap, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, wk_log2, one); break;
        }
        return;
    }
    // full-width tile: 1..4 groups of 32 kept sets (row mode 5..8) x 128 alleles
    switch ((item.shape & 0xff) - 4) {
        case 1: score_item_h<1, F8>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, one); break;
        case 2: score_item_h<2, F8>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, one); break;
        case 3: score_item_h<3, F8>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, one); break;
        default: score_item_h<4, F8>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next, flush, one); break;
    }
}

template <int KM>
__device__ __forceinline__ void score_dispatch_a(int am, const GkScoreItem& item, const GkMatrix& M,
                                                 const GkSearch& X, const float* __restrict__ L_pool,
                                                 const float* __restrict__ P_pool, uint32_t* __restrict__ S_pool,
                                                 float* smem, uint64_t* full, uint64_t* empty, int* next) {
    switch (am) {
        case F8: score_item<KM, F8>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        case F4: score_item<KM, F4>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        case S1: score_item<KM, S1>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        case S2: score_item<KM, S2>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        default: score_item<KM, S3>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
    }
}

__global__ void __launch_bounds__(kThreads, 2)
gk_score_kernel(const GkMatrix* __restrict__ matrices, const GkSearch* __restrict__ searches,
                const GkScoreItem* __restrict__ items, const float* __restrict__ L_pool,
                const float* __restrict__ P_pool, uint32_t* __restrict__ S_pool,
                const int32_t* __restrict__ kept_count) {
    // items may have been sized from an upper bound of the kept-set count (no host round trip)
    if (kept_count != nullptr && items[blockIdx.x].k_blk * GK_KB >= kept_count[items[blockIdx.x].search]) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    int* next = reinterpret_cast<int*>(smem_raw + 64);      // next tile to fetch (see score_item)
    float* smem = reinterpret_cast<float*>(smem_raw + 128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesF; ++s) {
            gk_mbar_init(&full[s], 1);
            gk_mbar_init(&empty[s], kWarps);
        }
        gk_fence_barrier_init();
    }
    __syncthreads();

    const GkScoreItem item = items[blockIdx.x];
    const GkSearch X = searches[item.search];
    const GkMatrix M = matrices[X.matrix];
    const int km = item.shape & 0xff;         // Mode of the kept-set dimension
    const int am = (item.shape >> 8) & 0xff;  // Mode of the allele dimension
    switch (km) {
        case F8: score_dispatch_a<F8>(am, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        case F4: score_dispatch_a<F4>(am, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        case S1: score_dispatch_a<S1>(am, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        case S2: score_dispatch_a<S2>(am, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
        default: score_dispatch_a<S3>(am, item, M, X, L_pool, P_pool, S_pool, smem, full, empty, next); break;
    }
}

constexpr int kSmemBytesF = kStagesF * GK_RT * (2 * GK_KB + 128) * (int)sizeof(float);
constexpr int kSmemBytesP = kStagesP * GK_RT * (2 * GK_KB * 2 + 128 * 4);
constexpr int kSmemF = 128 + kSmemBytesF;
constexpr int kSmemP = 128 + kSmemBytesP;

}  // namespace

#ifdef GK_SCORE_PROBE
// debug builds only: read and clear the pipeline probe
extern "C" int gk_score_probe(unsigned long long* out8) {
    unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out8, gk_probe, sizeof(zero)) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(gk_probe, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int gk_score(const GkMatrix* matrices, const GkSearch* searches, const GkScoreItem* items,
                        int n_items, const float* L_pool, const void* P_pool, uint32_t* S_pool, int half_mode,
                        int flush_stages, const int32_t* kept_count, void* stream) {
    if (n_items <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (half_mode) {
        GK_REQUIRE(flush_stages >= 1, "gk_score: flush interval %d must be >= 1 stage", flush_stages);
        cudaError_t err = cudaFuncSetAttribute(gk_score_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kSmemP);
        GK_REQUIRE(err == cudaSuccess, "gk_score: cannot reserve %d bytes of shared memory: %s", kSmemP,
                   cudaGetErrorString(err));
        gk_score_packed_kernel<<<n_items, kThreads, kSmemP, st>>>(matrices, searches, items, L_pool,
                                                                  reinterpret_cast<const uint16_t*>(P_pool), S_pool,
                                                                  flush_stages, kept_count, 1u);
        GK_CHECK_LAUNCH("gk_score (packed)");
        return 0;
    }
    cudaError_t err = cudaFuncSetAttribute(gk_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF);
    GK_REQUIRE(err == cudaSuccess, "gk_score: cannot reserve %d bytes of shared memory: %s", kSmemF,
               cudaGetErrorString(err));
    gk_score_kernel<<<n_items, kThreads, kSmemF, st>>>(matrices, searches, items, L_pool,
                                                       reinterpret_cast<const float*>(P_pool), S_pool, kept_count);
    GK_CHECK_LAUNCH("gk_score");
    return 0;
}
