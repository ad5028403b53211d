// Kernel (b): max-then-sum candidate scoring, tiled like a GEMM.
//
// Replaces   np.maximum(log_probs[:, idx], prev.T[:, :, None]).sum(axis=1)
// (reference: graphkir/typing_mulit_allele.py:540-542) without materialising
// the K x R x A temporary.  In mismatch-count form (max of log-probs == min of
// mismatch counts) one work item computes, for a 128-set x a_tile-candidate tile
// and a chunk of reads,
//     S[k, a] += sum_r min(L[r, a], P[r, k])
// Reads are the reduction dimension.  Operands are float32 holding small
// integers, so FMNMX + FADD are exact while a partial sum stays below 2^24
// (the host bounds a chunk to 8192 reads x 255); the partial is converted to
// an integer and merged with a 32-bit integer atomic, which makes the split-R
// reduction order-independent and bit-reproducible.
//
// Data movement: L and P are stored blocked ([a_blk][r][a_tile], [k_blk][r][128])
// so a stage (GK_RT reads of both tiles) is two contiguous spans, moved by the
// TMA engine with cp.async.bulk and signalled on an mbarrier; 4 stages.
// Math: 16x16 threads, each an 8 x (a_tile/16) register tile -> per read
// 8+TA shared loads feed 8*TA FMNMX + 8*TA FADD; no tensor cores (max-then-sum
// is not a multiply-accumulate).
//
// Bound: FP32 non-tensor issue.  One cell = 1 FMNMX (ALU pipe) + 1 FADD (FMA
// pipe) = 2 issue slots; peak = 148 SM x 4 schedulers x 32 lanes x f_clk / 2.
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 4;
constexpr int kTK = 8;  // kept sets per thread

template <int TA>
__device__ __forceinline__ void load_cols(const float* src, float (&dst)[TA]) {
    if constexpr (TA == 8) {
        const float4 x = *reinterpret_cast<const float4*>(src);
        const float4 y = *reinterpret_cast<const float4*>(src + 4);
        dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w;
        dst[4] = y.x; dst[5] = y.y; dst[6] = y.z; dst[7] = y.w;
    } else if constexpr (TA == 4) {
        const float4 x = *reinterpret_cast<const float4*>(src);
        dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w;
    } else if constexpr (TA == 2) {
        const float2 x = *reinterpret_cast<const float2*>(src);
        dst[0] = x.x; dst[1] = x.y;
    } else {
        dst[0] = *src;
    }
}

template <int TA>
__device__ __forceinline__ void score_item(const GkScoreItem& item, const GkMatrix& M,
                                           const GkSearch& X, const float* __restrict__ L_pool,
                                           const float* __restrict__ P_pool,
                                           uint32_t* __restrict__ S_pool, float* smem,
                                           uint64_t* full, uint64_t* empty) {
    constexpr int BA = 16 * TA;
    constexpr uint32_t kBytesP = GK_RT * GK_KB * sizeof(float);
    constexpr uint32_t kBytesL = GK_RT * BA * sizeof(float);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tk = tid >> 4;
    const int ta = tid & 15;

    float* sP = smem;
    float* sL = smem + kStages * GK_RT * GK_KB;
    const float* gL = L_pool + M.L_off + ((int64_t)item.a_blk * M.r_pad + item.r0) * BA;
    const float* gP = P_pool + X.P_off + ((int64_t)item.k_blk * M.r_pad + item.r0) * GK_KB;
    const int n_tiles = (item.r1 - item.r0) / GK_RT;

    if (tid == 0) {
        const int pre = n_tiles < kStages ? n_tiles : kStages;
        for (int s = 0; s < pre; ++s) {
            gk_mbar_arrive_expect_tx(&full[s], kBytesP + kBytesL);
            gk_bulk_g2s(sP + s * GK_RT * GK_KB, gP + (int64_t)s * GK_RT * GK_KB, kBytesP, &full[s]);
            gk_bulk_g2s(sL + s * GK_RT * BA, gL + (int64_t)s * GK_RT * BA, kBytesL, &full[s]);
        }
    }

    float acc[kTK][TA];
#pragma unroll
    for (int i = 0; i < kTK; ++i)
#pragma unroll
        for (int j = 0; j < TA; ++j) acc[i][j] = 0.f;

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
                gk_mbar_arrive_expect_tx(&full[sp], kBytesP + kBytesL);
                gk_bulk_g2s(sP + sp * GK_RT * GK_KB, gP + (int64_t)nt * GK_RT * GK_KB, kBytesP, &full[sp]);
                gk_bulk_g2s(sL + sp * GK_RT * BA, gL + (int64_t)nt * GK_RT * BA, kBytesL, &full[sp]);
            }
        }
        __syncwarp();
        gk_mbar_wait(&full[s], (t / kStages) & 1);

        const float* p = sP + s * GK_RT * GK_KB + tk * kTK;
        const float* l = sL + s * GK_RT * BA + ta * TA;
#pragma unroll 4
        for (int r = 0; r < GK_RT; ++r) {
            float pv[kTK];
            float lv[TA];
            load_cols<kTK>(p + r * GK_KB, pv);
            load_cols<TA>(l + r * BA, lv);
#pragma unroll
            for (int i = 0; i < kTK; ++i)
#pragma unroll
                for (int j = 0; j < TA; ++j) acc[i][j] += fminf(pv[i], lv[j]);
        }
        __syncwarp();
        if (lane == 0) gk_mbar_arrive(&empty[s]);
    }

    uint32_t* S = S_pool + X.S_off;
    const int k_base = item.k_blk * GK_KB + tk * kTK;
    const int a_base = item.a_blk * BA + ta * TA;
#pragma unroll
    for (int i = 0; i < kTK; ++i) {
#pragma unroll
        for (int j = 0; j < TA; ++j) {
            const uint32_t v = (uint32_t)acc[i][j];
            if (v) atomicAdd(S + (int64_t)(k_base + i) * X.s_stride + a_base + j, v);
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
    switch (M.a_tile) {
        case 128: score_item<8>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty); break;
        case 64:  score_item<4>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty); break;
        case 32:  score_item<2>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty); break;
        default:  score_item<1>(item, M, X, L_pool, P_pool, S_pool, smem, full, empty); break;
    }
}

constexpr int kSmemBytes = 128 + kStages * GK_RT * (GK_KB + 128) * (int)sizeof(float);

}  // namespace

extern "C" int gk_score(const GkMatrix* matrices, const GkSearch* searches, const GkScoreItem* items,
                        int n_items, const float* L_pool, const float* P_pool, uint32_t* S_pool,
                        void* stream) {
    if (n_items <= 0) return 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(gk_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kSmemBytes);
        GK_REQUIRE(err == cudaSuccess, "gk_score: cannot reserve %d bytes of shared memory: %s", kSmemBytes,
                   cudaGetErrorString(err));
        configured = true;
    }
    gk_score_kernel<<<n_items, kThreads, kSmemBytes, (cudaStream_t)stream>>>(matrices, searches, items, L_pool,
                                                                              P_pool, S_pool);
    GK_CHECK_LAUNCH("gk_score");
    return 0;
}
