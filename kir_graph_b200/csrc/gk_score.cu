// Kernel (b): max-then-sum candidate scoring, tiled like a GEMM.
//
// Replaces   np.maximum(log_probs[:, idx], prev.T[:, :, None]).sum(axis=1)
// (reference: graphkir/typing_mulit_allele.py:540-542) without materialising
// the K x R x A temporary.  In mismatch-count form (max of log-probs == min of
// mismatch counts) one work item computes, for a tile of kept sets x candidate
// alleles and a chunk of reads,
//     D[k, a] += sum_r |L[r, a] - P[r, k]|        (sum of absolute differences)
// from which the consumers recover the min-sum exactly:
//     sum_r min(L, P) = (colsum_L[a] + colsum_P[k] - D[k, a]) / 2
// with colsum_L the CN=1 column sums and colsum_P[k] the previous step's score of set k
// (both already known), so the inner loop is FADD (p - l) + FADD (acc += |d|, the
// absolute value is a free source modifier).  The direct form FMNMX + FADD computes the
// same thing but FMNMX issues on the half-rate ALU pipe, which ncu showed to be the
// limiter (profiles/r01_score_v0_ncu_summary.txt: ALU 78 %, FMA 39 %); both forms are
// 2 FP32 non-tensor instructions per cell.
// Reads are the reduction dimension.  Operands are float32 holding small
// integers, so the two FADDs are exact while a partial sum stays below 2^24
// (the host bounds a chunk to 8192 reads x 255); the partial is converted to
// an integer and merged with a 32-bit integer atomic, which makes the split-R
// reduction order-independent and bit-reproducible.
//
// Data movement: L and P are stored blocked ([a_blk][r][a_tile], [k_blk][r][64]),
// so each block's GK_RT rows of a stage are one contiguous span, moved by the TMA
// engine with cp.async.bulk and signalled on an mbarrier; 4 stages in flight.
// A CTA tile is 1-2 k-blocks x 1-4 a-blocks (64/128 sets x 16..128 alleles) so that
// ragged K (top_n = 300) and ragged A are covered without computing padding.
// Math: 16x16 threads, each a TK x TA register tile (8x8 for the full tile) whose
// rows/columns are interleaved in groups of four (k = 4*tk + i, 64 + 4*tk + i) so a
// half-warp's 128-bit shared loads hit consecutive banks.  No tensor cores:
// max-then-sum is not a multiply-accumulate.
//
// Bound: FP32 non-tensor issue.  One cell = 2 FADD = 2 issue slots of the 4 x 32-lane
// schedulers; peak = 148 SM x 64 cells/clk x f_clk.
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 4;

// TK in {8, 4}: 128 or 64 kept sets per CTA.  TA in {8, 4, 2, 1}: 128/64/32 alleles
// (a_tile 32) or 16 alleles (a_tile 16).
template <int TK, int TA>
__device__ __forceinline__ void score_item(const GkScoreItem& item, const GkMatrix& M, const GkSearch& X,
                                           const float* __restrict__ L_pool, const float* __restrict__ P_pool,
                                           uint32_t* __restrict__ S_pool, float* smem, uint64_t* full,
                                           uint64_t* empty) {
    constexpr int KW = TK / 4;                  // k-blocks of 64
    constexpr int BA = 16 * TA;                 // alleles per CTA tile
    constexpr int AT = TA == 1 ? 16 : 32;       // layout block width of L
    constexpr int AW = BA / AT;                 // a-blocks per CTA tile
    constexpr uint32_t kBytesPBlk = GK_RT * GK_KB * sizeof(float);
    constexpr uint32_t kBytesLBlk = GK_RT * AT * sizeof(float);
    constexpr uint32_t kStageBytes = KW * kBytesPBlk + AW * kBytesLBlk;
    constexpr int kStageFloats = GK_RT * (KW * GK_KB + AW * AT);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tk = tid >> 4;
    const int ta = tid & 15;

    const int64_t blk_stride_p = (int64_t)M.r_pad * GK_KB;
    const int64_t blk_stride_l = (int64_t)M.r_pad * AT;
    const float* gP = P_pool + X.P_off + item.k_blk * blk_stride_p + (int64_t)item.r0 * GK_KB;
    const float* gL = L_pool + M.L_off + item.a_blk * blk_stride_l + (int64_t)item.r0 * AT;
    const int n_tiles = (item.r1 - item.r0) / GK_RT;

    auto issue = [&](int tile, int s) {
        float* dst = smem + s * kStageFloats;
        gk_mbar_arrive_expect_tx(&full[s], kStageBytes);
#pragma unroll
        for (int b = 0; b < KW; ++b)
            gk_bulk_g2s(dst + b * GK_RT * GK_KB, gP + b * blk_stride_p + (int64_t)tile * GK_RT * GK_KB,
                        kBytesPBlk, &full[s]);
        dst += KW * GK_RT * GK_KB;
#pragma unroll
        for (int b = 0; b < AW; ++b)
            gk_bulk_g2s(dst + b * GK_RT * AT, gL + b * blk_stride_l + (int64_t)tile * GK_RT * AT, kBytesLBlk,
                        &full[s]);
    };

    if (tid == 0) {
        const int pre = n_tiles < kStages ? n_tiles : kStages;
        for (int s = 0; s < pre; ++s) issue(s, s);
    }

    float acc[TK][TA];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < TA; ++j) acc[i][j] = 0.f;

    // shared-memory offsets of this thread's operands inside a stage
    // k = 4*tk + i (block 0) and, for TK == 8, the same offsets in block 1
    const int p_off = tk * 4;
    // a: TA == 8 -> cols 4*ta + j and 64 + 4*ta + j ; TA == 4 -> 4*ta + j ; TA == 2 -> 2*ta + j ; TA == 1 -> ta
    int l_off[2];
    {
        const int c0 = (TA >= 4) ? ta * 4 : ta * TA;
        l_off[0] = (c0 / AT) * (GK_RT * AT) + (c0 % AT);
        const int c1 = 64 + ta * 4;
        l_off[1] = (c1 / AT) * (GK_RT * AT) + (c1 % AT);
    }

#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kStages;
        // refill the stage that held tile t-1 once every warp has released it
        if (tid == 0 && t >= 1) {
            const int tp = t - 1;
            const int nt = tp + kStages;
            if (nt < n_tiles) {
                const int sp = tp % kStages;
                gk_mbar_wait(&empty[sp], (tp / kStages) & 1);
                issue(nt, sp);
            }
        }
        __syncwarp();
        gk_mbar_wait(&full[s], (t / kStages) & 1);

        const float* p = smem + s * kStageFloats + p_off;
        const float* l = smem + s * kStageFloats + KW * GK_RT * GK_KB;
#pragma unroll 4
        for (int r = 0; r < GK_RT; ++r) {
            float pv[TK];
            float lv[TA];
            {
                const float4 x = *reinterpret_cast<const float4*>(p + r * GK_KB);
                pv[0] = x.x; pv[1] = x.y; pv[2] = x.z; pv[3] = x.w;
                if constexpr (TK == 8) {
                    const float4 y = *reinterpret_cast<const float4*>(p + GK_RT * GK_KB + r * GK_KB);
                    pv[4] = y.x; pv[5] = y.y; pv[6] = y.z; pv[7] = y.w;
                }
            }
            if constexpr (TA >= 4) {
                const float4 x = *reinterpret_cast<const float4*>(l + l_off[0] + r * AT);
                lv[0] = x.x; lv[1] = x.y; lv[2] = x.z; lv[3] = x.w;
                if constexpr (TA == 8) {
                    const float4 y = *reinterpret_cast<const float4*>(l + l_off[1] + r * AT);
                    lv[4] = y.x; lv[5] = y.y; lv[6] = y.z; lv[7] = y.w;
                }
            } else if constexpr (TA == 2) {
                const float2 x = *reinterpret_cast<const float2*>(l + l_off[0] + r * AT);
                lv[0] = x.x; lv[1] = x.y;
            } else {
                lv[0] = l[l_off[0] + r * AT];
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
        const int k = k_base + (i < 4 ? tk * 4 + i : 64 + tk * 4 + (i - 4));
#pragma unroll
        for (int j = 0; j < TA; ++j) {
            int a;
            if constexpr (TA >= 4) a = a_base + (j < 4 ? ta * 4 + j : 64 + ta * 4 + (j - 4));
            else a = a_base + ta * TA + j;
            const uint32_t v = (uint32_t)acc[i][j];
            if (v) atomicAdd(S + (int64_t)k * X.s_stride + a, v);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 2)
gk_score_kernel(const GkMatrix* __restrict__ matrices, const GkSearch* __restrict__ searches,
                const GkScoreItem* __restrict__ items, const float* __restrict__ L_pool,
                const float* __restrict__ P_pool, uint32_t* __restrict__ S_pool) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kStages;
    float* smem = reinterpret_cast<float*>(smem_raw + 128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            gk_mbar_init(&full[s], 1);
            gk_mbar_init(&empty[s], kWarps);
        }
        gk_fence_barrier_init();
    }
    __syncthreads();

    const GkScoreItem item = items[blockIdx.x];
    const GkSearch X = searches[item.search];
    const GkMatrix M = matrices[X.matrix];
    const int kw = item.shape & 0xff;         // k-blocks of 64: 1 or 2
    const int aw = (item.shape >> 8) & 0xff;  // a-blocks: 1, 2 or 4 (a_tile 32) / 1 (a_tile 16)
#define GK_SCORE_CASE(TK, TA) score_item<TK, TA>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty)
    if (M.a_tile == 16) {
        if (kw == 2) GK_SCORE_CASE(8, 1); else GK_SCORE_CASE(4, 1);
    } else if (aw == 4) {
        if (kw == 2) GK_SCORE_CASE(8, 8); else GK_SCORE_CASE(4, 8);
    } else if (aw == 2) {
        if (kw == 2) GK_SCORE_CASE(8, 4); else GK_SCORE_CASE(4, 4);
    } else {
        if (kw == 2) GK_SCORE_CASE(8, 2); else GK_SCORE_CASE(4, 2);
    }
#undef GK_SCORE_CASE
}

constexpr int kSmemBytes = 128 + kStages * GK_RT * (2 * GK_KB + 128) * (int)sizeof(float);

}  // namespace

extern "C" int gk_score(const GkMatrix* matrices, const GkSearch* searches, const GkScoreItem* items,
                        int n_items, const float* L_pool, const float* P_pool, uint32_t* S_pool,
                        void* stream) {
    if (n_items <= 0) return 0;
    cudaError_t err =
        cudaFuncSetAttribute(gk_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    GK_REQUIRE(err == cudaSuccess, "gk_score: cannot reserve %d bytes of shared memory: %s", kSmemBytes,
               cudaGetErrorString(err));
    gk_score_kernel<<<n_items, kThreads, kSmemBytes, (cudaStream_t)stream>>>(matrices, searches, items, L_pool,
                                                                              P_pool, S_pool);
    GK_CHECK_LAUNCH("gk_score");
    return 0;
}
