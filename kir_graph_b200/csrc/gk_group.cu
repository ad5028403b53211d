// Read grouping by called alleles.
//
// Replaces, in graphkir/novel_discover.py:62-64,
//     is_max = np.equal(probs[:, ids], probs[:, ids].max(axis=1)[:, None])
// on the device-resident likelihood: probs[r, a] is strictly decreasing in the mismatch count
// m[r, a] for a fixed read, so "attains the row maximum of probs" is "attains the row minimum of
// m" - exactly, without the reference's sensitivity to the rounding of its ordered float product.
// One thread per read, the allele-major byte matrix LT is read coalesced along reads:
//     pattern[r] bit t = member t of `ids` attains min_t m[r, ids[t]]
// HBM-bound: n_ids bytes read and 4 bytes written per read.
#include "gk_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
gk_group_reads_kernel(const GkMatrix* __restrict__ matrices, int matrix, const int32_t* __restrict__ ids,
                      int n_ids, const uint8_t* __restrict__ LT_pool, uint32_t* __restrict__ pattern) {
    const GkMatrix M = matrices[matrix];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M.n_reads) return;
    const uint8_t* LT = LT_pool + M.LT_off + r;
    unsigned int mn = 256u, pat = 0u;
    for (int t = 0; t < n_ids; ++t) {
        const unsigned int v = __ldg(LT + (int64_t)__ldg(ids + t) * M.r_pad);
        if (v < mn) {
            mn = v;
            pat = 1u << t;
        } else if (v == mn) {
            pat |= 1u << t;
        }
    }
    pattern[r] = pat;
}

}  // namespace

extern "C" int gk_group_reads(const GkMatrix* matrices, int matrix, int n_reads, const int32_t* ids, int n_ids,
                              const uint8_t* LT_pool, uint32_t* pattern, void* stream) {
    GK_REQUIRE(n_ids >= 1 && n_ids <= 32, "gk_group_reads: %d alleles outside 1..32", n_ids);
    if (n_reads <= 0) return 0;
    gk_group_reads_kernel<<<(n_reads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(matrices, matrix, ids, n_ids,
                                                                                 LT_pool, pattern);
    GK_CHECK_LAUNCH("gk_group_reads");
    return 0;
}
