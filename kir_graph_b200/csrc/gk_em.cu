// EM path (reference: graphkir/typing_em.py).
//
//   gk_em_compat    per read pair the set of compatible alleles as a bitset over alleles:
//                   per mate  AND of the allele sets of its positive variants, minus the
//                   union of the allele sets of its negative variants (empty when the mate
//                   has no positive variant)  -- getCandidateAllelePerRead (:68-87);
//                   pair = alleles with the highest multiplicity over both mates' lists,
//                   i.e. the intersection when it is non-empty, else the union
//                   -- getMostFreqAllele (:90-104).  One warp per read, lanes over the
//                   32-allele words of the variant-major membership table.
//   gk_em_squarem   SQUAREM-accelerated EM over the distinct compatibility rows (with
//                   multiplicities) of one gene -- hisatEMnp (:107-188).  One CTA per gene,
//                   float64, fixed summation order (bit-reproducible run to run).
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;

// membT[v * n_awords + w]: bit b set <=> allele 32w+b carries variant v
__global__ void __launch_bounds__(kThreads)
gk_em_compat_kernel(const uint32_t* __restrict__ membT, int n_awords, int n_alleles,
                    const int32_t* __restrict__ off_lp, const int32_t* __restrict__ idx_lp,
                    const int32_t* __restrict__ off_ln, const int32_t* __restrict__ idx_ln,
                    const int32_t* __restrict__ off_rp, const int32_t* __restrict__ idx_rp,
                    const int32_t* __restrict__ off_rn, const int32_t* __restrict__ idx_rn, int n_reads,
                    uint32_t* __restrict__ compat) {
    const int r = blockIdx.x * (kThreads / 32) + gk_warp();
    if (r >= n_reads) return;
    const int lane = gk_lane();
    for (int w = lane; w < n_awords; w += 32) {
        const uint32_t tail = (w == n_awords - 1 && (n_alleles & 31)) ? ((1u << (n_alleles & 31)) - 1u) : 0xffffffffu;
        uint32_t mate[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int32_t* op = m ? off_rp : off_lp;
            const int32_t* ip = m ? idx_rp : idx_lp;
            const int32_t* on = m ? off_rn : off_ln;
            const int32_t* in = m ? idx_rn : idx_ln;
            uint32_t acc = 0u;
            if (op[r + 1] > op[r]) {
                acc = tail;
                for (int e = op[r]; e < op[r + 1]; ++e) acc &= membT[(int64_t)ip[e] * n_awords + w];
                for (int e = on[r]; e < on[r + 1]; ++e) acc &= ~membT[(int64_t)in[e] * n_awords + w];
            }
            mate[m] = acc;
        }
        // does any word of the intersection have a bit?  (warp-wide when n_awords <= 32, else looped below)
        compat[(int64_t)r * n_awords + w] = mate[0] & mate[1];
        compat[(int64_t)(n_reads + r) * n_awords + w] = mate[0] | mate[1];
    }
}

// pick intersection rows when non-empty, else union rows
__global__ void gk_em_pick_kernel(int n_awords, int n_reads, uint32_t* __restrict__ compat) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint32_t any = 0u;
    for (int w = 0; w < n_awords; ++w) any |= compat[(int64_t)r * n_awords + w];
    if (!any)
        for (int w = 0; w < n_awords; ++w) compat[(int64_t)r * n_awords + w] = compat[(int64_t)(n_reads + r) * n_awords + w];
}

typedef GkEmProblem EmProblem;

__device__ __forceinline__ double block_sum(double v, double* red) {
    // fixed-order tree: lanes by shuffle, then warps in order by thread 0
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += red[w];
        red[32] = total;
    }
    __syncthreads();
    return red[32];
}

// p_out = getNextProb(p_in)   (typing_em.py:153-161)
__device__ void em_step(const EmProblem& E, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ wgt,
                        const double* __restrict__ len, const double* p_in, double* p_out, double* binv,
                        double* red) {
    const int A = E.n_alleles, W = E.n_awords;
    for (int u = threadIdx.x; u < E.n_rows; u += blockDim.x) {
        double b = 0.0;
        for (int w = 0; w < W; ++w) {
            uint32_t bits = rows[(int64_t)u * W + w];
            while (bits) {
                const int a = 32 * w + __ffs(bits) - 1;
                bits &= bits - 1;
                b += p_in[a];
            }
        }
        binv[u] = b != 0.0 ? (double)wgt[u] / b : 0.0;
    }
    __syncthreads();
    double mine = 0.0;
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const int w = a >> 5;
        const uint32_t bit = 1u << (a & 31);
        double acc = 0.0;
        for (int u = 0; u < E.n_rows; ++u)
            if (rows[(int64_t)u * W + w] & bit) acc += binv[u];
        acc = acc * p_in[a] / len[a];
        p_out[a] = acc;
        mine += acc;
    }
    const double total = block_sum(mine, red);
    for (int a = threadIdx.x; a < A; a += blockDim.x) p_out[a] = p_out[a] / total;
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
gk_em_squarem_kernel(const EmProblem* __restrict__ problems, const uint32_t* __restrict__ row_pool,
                     const uint32_t* __restrict__ wgt_pool, const double* __restrict__ len_pool,
                     double* __restrict__ out_pool, int32_t* __restrict__ iters_out, int iter_max,
                     double diff_threshold) {
    __shared__ double red[40];
    const EmProblem E = problems[blockIdx.x];
    const int A = E.n_alleles;
    const uint32_t* rows = row_pool + E.row_off;
    const uint32_t* wgt = wgt_pool + E.wgt_off;
    double* prob = out_pool + E.out_off;
    const double* len = len_pool + E.len_off;
    double* p1 = prob + A;
    double* p2 = p1 + A;
    double* p3 = p2 + A;
    double* tmp = p3 + A;
    double* binv = tmp + A;

    for (int a = threadIdx.x; a < A; a += blockDim.x) tmp[a] = 1.0;
    __syncthreads();
    em_step(E, rows, wgt, len, tmp, prob, binv, red);               // prob = next(ones)  (:164)
    int iters = 0;
    for (; iters < iter_max; ++iters) {
        em_step(E, rows, wgt, len, prob, p1, binv, red);            // prob_next
        em_step(E, rows, wgt, len, p1, p2, binv, red);              // prob_next2
        double rr = 0.0, vv = 0.0;
        for (int a = threadIdx.x; a < A; a += blockDim.x) {
            const double r = p1[a] - prob[a];
            const double v = p2[a] - p1[a] - r;
            rr += r * r;
            vv += v * v;
        }
        rr = block_sum(rr, red);
        vv = block_sum(vv, red);
        if (vv > 0.0) {
            const double g = -sqrt(rr / vv);
            for (int a = threadIdx.x; a < A; a += blockDim.x) {
                const double r = p1[a] - prob[a];
                const double v = p2[a] - p1[a] - r;
                const double x = prob[a] - r * g * 2.0 + v * (g * g);
                p3[a] = x > 0.0 ? x : 0.0;
            }
            __syncthreads();
            em_step(E, rows, wgt, len, p3, p1, binv, red);          // prob_next = next(prob_next3)
        }
        double diff = 0.0;
        for (int a = threadIdx.x; a < A; a += blockDim.x) diff += fabs(prob[a] - p1[a]);
        diff = block_sum(diff, red);
        if (diff <= diff_threshold) break;                          // returns prob, not prob_next (:181-186)
        for (int a = threadIdx.x; a < A; a += blockDim.x) prob[a] = p1[a];
        __syncthreads();
    }
    if (threadIdx.x == 0) iters_out[blockIdx.x] = iters;
}

}  // namespace

extern "C" int gk_em_compat(const uint32_t* membT, int n_awords, int n_alleles, const int32_t* off_lp,
                            const int32_t* idx_lp, const int32_t* off_ln, const int32_t* idx_ln,
                            const int32_t* off_rp, const int32_t* idx_rp, const int32_t* off_rn,
                            const int32_t* idx_rn, int n_reads, uint32_t* compat /* [2 * n_reads][n_awords] */,
                            void* stream) {
    if (n_reads <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int per = kThreads / 32;
    gk_em_compat_kernel<<<(n_reads + per - 1) / per, kThreads, 0, st>>>(membT, n_awords, n_alleles, off_lp, idx_lp,
                                                                        off_ln, idx_ln, off_rp, idx_rp, off_rn,
                                                                        idx_rn, n_reads, compat);
    GK_CHECK_LAUNCH("gk_em_compat");
    gk_em_pick_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(n_awords, n_reads, compat);
    GK_CHECK_LAUNCH("gk_em_pick");
    return 0;
}

extern "C" int gk_em_squarem(const GkEmProblem* problems, int n_problems, const uint32_t* row_pool,
                             const uint32_t* wgt_pool, const double* len_pool, double* out_pool,
                             int32_t* iters_out, int iter_max, double diff_threshold, void* stream) {
    if (n_problems <= 0) return 0;
    gk_em_squarem_kernel<<<n_problems, 1024, 0, (cudaStream_t)stream>>>(
        problems, row_pool, wgt_pool, len_pool, out_pool, iters_out, iter_max,
        diff_threshold);
    GK_CHECK_LAUNCH("gk_em_squarem");
    return 0;
}

