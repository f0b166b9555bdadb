// Post-fit statistics on the device: central credible intervals of the Gamma / Beta guide distributions
// over the whole (K, Nt, F, Q) parameter arrays, and signal-to-noise ratio / chi2 per patch.
//
// Replaces, for these arrays, what the reference does on the CPU: cosmos.compute_params (models/cosmos.py:711-784) ->
// stats.torch_to_scipy_dist (utils/stats.py:262-293) -> scipy.stats.gamma / beta `.interval(CI)` (inverse regularised
// incomplete gamma / beta functions, one element at a time inside scipy: 90 M evaluations at 1000 AOIs x 5000 frames),
// and stats.snr_and_chi2 (utils/stats.py:29-86) called per AOI from a Python loop (:193-215).
//
// Inverse CDFs: the textbook scheme (series / continued fraction for the regularised incomplete function, a closed-form
// starting point, Halley steps on  F(x) - p  with the density as derivative), all in double, one thread per element,
// both ends of the interval per thread.  Checked against scipy at 1e-9 over the ranges the guides visit
// (tests/test_stats_gpu.py) and against the reference's own compute_params output (tests/golden/ref_c1_fit.pt).
#include "common.cuh"
#include "stats_math.cuh"

namespace tq {

__global__ void __launch_bounds__(128) gamma_interval_kernel(int64_t n, const double* __restrict__ conc, const double* __restrict__ rate,
                                                             double p_lo, double p_hi, double* __restrict__ lo, double* __restrict__ hi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = conc[i], r = rate[i];
        lo[i] = gamma_p_inv(p_lo, a) / r;
        hi[i] = gamma_p_inv(p_hi, a) / r;
    }
}

__global__ void __launch_bounds__(128) beta_interval_kernel(int64_t n, const double* __restrict__ c1, const double* __restrict__ c0,
                                                            double p_lo, double p_hi, double* __restrict__ lo, double* __restrict__ hi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = c1[i], b = c0[i];
        lo[i] = beta_i_inv(p_lo, a, b);
        hi[i] = beta_i_inv(p_hi, a, b);
    }
}

// ---- SNR and chi2 per patch (utils/stats.py:29-86): one warp per (AOI, frame, channel) patch ------------------------------
//   snr[k]  = sum_ij (D - b - mu_off) N_k(i, j) / sqrt(var_off + b gain),   N_k = gaussian_spots / height
//   chi2    = mean_ij (D - ideal - mu_off)^2 / ideal,                       ideal = b + sum_k h_k N_k
// Spot parameters are (K, U) SoA (U = patches in store order), the separable factors of the K spots go through a
// per-warp table (2 K P exponentials per patch instead of K P P).
constexpr int kStatWarps = 4;
constexpr int kStatK = 2;
constexpr int kStatMaxP = 32;

template <typename PIX>
__global__ void __launch_bounds__(kStatWarps * 32) snr_chi2_kernel(int64_t U, int P, const PIX* __restrict__ pixels,
                                                                   const float* __restrict__ xy, const float* __restrict__ height,
                                                                   const float* __restrict__ width, const float* __restrict__ x,
                                                                   const float* __restrict__ y, const float* __restrict__ background,
                                                                   float gain, float off_mean, float off_var,
                                                                   float* __restrict__ snr, float* __restrict__ chi2) {
    __shared__ float tab[kStatWarps][2 * kStatK * kStatMaxP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* gx = tab[warp];
    float* gy = gx + kStatK * kStatMaxP;
    const int PP = P * P;
    for (int64_t u = (int64_t)blockIdx.x * kStatWarps + warp; u < U; u += (int64_t)gridDim.x * kStatWarps) {
        float h[kStatK], w[kStatK], cx[kStatK], cy[kStatK], norm[kStatK];
        const float tx = xy[u * 2], ty = xy[u * 2 + 1], b = background[u];
#pragma unroll
        for (int k = 0; k < kStatK; ++k) {
            h[k] = height[k * U + u];
            w[k] = width[k * U + u];
            cx[k] = x[k * U + u] + tx;
            cy[k] = y[k * U + u] + ty;
            norm[k] = 1.0f / (6.283185307179586f * w[k] * w[k]);
        }
        __syncwarp();
        for (int idx = lane; idx < 2 * kStatK * P; idx += 32) {
            const int axis = idx / (kStatK * P), rem = idx - axis * (kStatK * P);
            const int k = rem / P, i = rem - k * P;
            const float d = float(i) - (axis == 0 ? cx[k] : cy[k]);
            (axis == 0 ? gx : gy)[k * kStatMaxP + i] = expf(-(d * d) / (2.0f * w[k] * w[k]));
        }
        __syncwarp();
        float sig[kStatK], c2 = 0.0f;
#pragma unroll
        for (int k = 0; k < kStatK; ++k) sig[k] = 0.0f;
        const PIX* pix = pixels + u * PP;
        for (int p = lane; p < PP; p += 32) {
            const int row = p / P, col = p - row * P;
            const float D = (float)pix[p];
            float ideal = b;
#pragma unroll
            for (int k = 0; k < kStatK; ++k) {
                const float nk = gx[k * kStatMaxP + col] * gy[k * kStatMaxP + row] * norm[k];
                sig[k] += (D - b - off_mean) * nk;
                ideal += h[k] * nk;
            }
            const float r = D - ideal - off_mean;
            c2 += r * r / ideal;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
#pragma unroll
            for (int k = 0; k < kStatK; ++k) sig[k] += __shfl_xor_sync(0xffffffffu, sig[k], o);
        }
        if (lane == 0) {
            const float noise = sqrtf(off_var + b * gain);
#pragma unroll
            for (int k = 0; k < kStatK; ++k) snr[k * U + u] = sig[k] / noise;
            chi2[u] = c2 / float(PP);
        }
    }
}

}  // namespace tq

using namespace tq;

static int interval_grid(int64_t n) {
    int64_t g = (n + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// Central credible interval of Gamma(conc, rate): lo / hi = quantiles (1 -+ ci) / 2.  All arrays double, n elements.
extern "C" int tq_gamma_interval(int64_t n, const double* conc, const double* rate, double ci, double* lo, double* hi, void* stream) {
    TQ_CHECK_ARG(n >= 0 && ci > 0.0 && ci < 1.0, "need n >= 0 and 0 < ci < 1");
    if (n == 0) return TQ_OK;
    TQ_CHECK_ARG(conc && rate && lo && hi, "NULL pointer");
    gamma_interval_kernel<<<interval_grid(n), 128, 0, (cudaStream_t)stream>>>(n, conc, rate, 0.5 * (1.0 - ci), 0.5 * (1.0 + ci), lo, hi);
    TQ_LAUNCH_CHECK("gamma_interval_kernel launch");
    return TQ_OK;
}

// Central credible interval of Beta(c1, c0) on [0, 1] (an AffineBeta's interval is low + scale * these).
extern "C" int tq_beta_interval(int64_t n, const double* c1, const double* c0, double ci, double* lo, double* hi, void* stream) {
    TQ_CHECK_ARG(n >= 0 && ci > 0.0 && ci < 1.0, "need n >= 0 and 0 < ci < 1");
    if (n == 0) return TQ_OK;
    TQ_CHECK_ARG(c1 && c0 && lo && hi, "NULL pointer");
    beta_interval_kernel<<<interval_grid(n), 128, 0, (cudaStream_t)stream>>>(n, c1, c0, 0.5 * (1.0 - ci), 0.5 * (1.0 + ci), lo, hi);
    TQ_LAUNCH_CHECK("beta_interval_kernel launch");
    return TQ_OK;
}

// SNR (K = 2 spots, (K, U)) and chi2 (U) of U patches in store order; pixtype as in tq_patch_view; K-major float arrays.
extern "C" int tq_snr_chi2(int64_t U, int P, int pixtype, const void* pixels, const void* xy, const void* height,
                           const void* width, const void* x, const void* y, const void* background, double gain,
                           double offset_mean, double offset_var, void* snr, void* chi2, void* stream) {
    TQ_CHECK_ARG(U >= 0 && P >= 2 && P <= kStatMaxP, "bad shape");
    if (U == 0) return TQ_OK;
    TQ_CHECK_ARG(pixels && xy && height && width && x && y && background && snr && chi2, "NULL pointer");
    int64_t grid = (U + kStatWarps - 1) / kStatWarps;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (grid > cap) grid = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define TQ_SNR_LAUNCH(PIXT)                                                                                                   \
    snr_chi2_kernel<PIXT><<<(int)grid, kStatWarps * 32, 0, st>>>(U, P, (const PIXT*)pixels, (const float*)xy, (const float*)height, \
        (const float*)width, (const float*)x, (const float*)y, (const float*)background, (float)gain, (float)offset_mean,       \
        (float)offset_var, (float*)snr, (float*)chi2)
    if (pixtype == TQ_PIX_U16) TQ_SNR_LAUNCH(uint16_t);
    else if (pixtype == TQ_PIX_F32) TQ_SNR_LAUNCH(float);
    else { set_error("tq_snr_chi2: pixtype %d not supported (uint16 / float32 stores)", pixtype); return TQ_ERR_ARG; }
#undef TQ_SNR_LAUNCH
    TQ_LAUNCH_CHECK("snr_chi2_kernel launch");
    return TQ_OK;
}
