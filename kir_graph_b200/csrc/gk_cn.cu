// CN model: the likelihood curve of CNgroup.fit and the CN-group
// probabilities of calcCNGroupProb (reference: graphkir/cn_model.py:124-204).
//
// For every candidate base b (the mean depth of one copy) the model is max_cn normal densities over
// the depth bins x_i - CN 0 at 0, CN n at n b, deviations growing with n - and
//     likelihood(b) = sum_i log(max_n pdf_n(x_i) * space + 1e-9) * density_i
// with density the histogram of the observed gene depths.  bin_num candidate bases x bin_num bins x
// max_cn densities = 1.75 M exp / log evaluations per fit at the defaults (500 x 500 x 7); depthToCN
// refits in a loop when KIR3DL3 is assumed diploid (kir_cn.py:96-107).  One CTA per base, float64
// throughout (the argmax over the bases must pick the reference's grid point), fixed summation order.
#include <math.h>

#include "gk_common.cuh"

namespace {

constexpr int kCnThreads = 256;

struct CnParams {
    double base_dev, y0_dev, dev_decay, dev_decay_neg, space;
    int bin_num, max_cn, start_base;
};

// scipy.stats.norm.pdf(x, loc, scale): exp(-y^2 / 2) / sqrt(2 pi) / scale, nan unless scale > 0
__device__ __forceinline__ double norm_pdf(double x, double loc, double scale) {
    if (!(scale > 0.0)) return nan("");
    const double y = (x - loc) / scale;
    return exp(-(y * y) / 2.0) / 2.5066282746310002 / scale;
}

__device__ __forceinline__ void row_of(const CnParams& p, double base, int row, double& loc, double& scale) {
    if (p.start_base == 1) {                     // cn_model.py:184-189
        if (row == 0) {
            loc = 0.0;
            scale = p.base_dev * p.y0_dev;
        } else {
            loc = base * (double)row;
            scale = p.base_dev * (p.dev_decay * (double)(row - 1) + 1.0);
        }
    } else {                                     // start_base == 2, :190-199
        loc = base * (double)row;
        scale = row < p.start_base ? p.base_dev * (p.dev_decay_neg * (double)(p.start_base - row) + 1.0)
                                   : p.base_dev * (p.dev_decay * (double)(row - p.start_base) + 1.0);
    }
}

__global__ void __launch_bounds__(kCnThreads)
gk_cn_fit_kernel(const double* __restrict__ x, const double* __restrict__ density, const double* __restrict__ bases,
                 CnParams p, double* __restrict__ likelihood, double* __restrict__ prob_out) {
    __shared__ double partial[kCnThreads];
    const double base = bases[blockIdx.x];
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.bin_num; i += kCnThreads) {
        double best = 0.0;
        bool any_nan = false;
        for (int row = 0; row < p.max_cn; ++row) {
            double loc, scale;
            row_of(p, base, row, loc, scale);
            const double v = norm_pdf(x[i], loc, scale) * p.space;
            if (prob_out != nullptr) prob_out[((size_t)blockIdx.x * p.max_cn + row) * p.bin_num + i] = v;
            any_nan |= isnan(v);                 // numpy's max propagates nan
            best = (row == 0 || v > best) ? v : best;
        }
        if (any_nan) best = nan("");
        acc += log(best + 1e-9) * density[i];
    }
    partial[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kCnThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) partial[threadIdx.x] += partial[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) likelihood[blockIdx.x] = partial[0];
}

}  // namespace

extern "C" int gk_cn_fit(const double* x, const double* density, const double* bases, int n_base, int bin_num, int max_cn,
                         int start_base, double base_dev, double y0_dev, double dev_decay, double dev_decay_neg,
                         double space, double* likelihood, double* prob_out, void* stream) {
    if (n_base <= 0) return 0;
    GK_REQUIRE(start_base == 1 || start_base == 2, "gk_cn_fit: start_base %d is not implemented (1 or 2)", start_base);
    GK_REQUIRE(bin_num >= 1 && max_cn >= 1 && max_cn <= 64, "gk_cn_fit: bin_num %d / max_cn %d out of range", bin_num, max_cn);
    CnParams p{base_dev, y0_dev, dev_decay, dev_decay_neg, space, bin_num, max_cn, start_base};
    gk_cn_fit_kernel<<<n_base, kCnThreads, 0, (cudaStream_t)stream>>>(x, density, bases, p, likelihood, prob_out);
    GK_CHECK_LAUNCH("gk_cn_fit");
    return 0;
}
