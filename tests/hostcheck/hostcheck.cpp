// CPU harness around the host+device arithmetic headers of tapqir_b200/csrc (TEST INFRASTRUCTURE).
// Compiled with g++ by tests/hostcheck/__init__.py; lets `-m "not gpu"` tests compare the exact
// per-pixel / per-unit formulas the kernels execute against the oracle without a GPU.  Nothing
// in the product imports or links this file.
#include <cstdint>
#include <vector>

#include "ksmogn_core.cuh"

using namespace tq;

template <typename T>
static void ksmogn_host(int64_t U, int P, int O, int NM, const T* height, const T* width, const T* x, const T* y,
                        const T* background, T gain, const T* mcfg, const T* W, const T* value, const T* target,
                        const T* off_s, const T* off_w, T* logp, T* g_h, T* g_w, T* g_x, T* g_y, T* g_b, T* g_rate) {
    const T rate = T(1) / gain, log_rate = Real<T>::log(rate);
    for (int64_t u = 0; u < U; ++u) {
        PatchSpots<T> s;
        for (int k = 0; k < kK; ++k) {
            s.h[k] = height[k * U + u]; s.w[k] = width[k * U + u];
            s.cx[k] = x[k * U + u] + target[u * 2]; s.cy[k] = y[k * U + u] + target[u * 2 + 1];
        }
        s.b = background[u];
        auto run = [&](auto nm_tag) {
            constexpr int NMc = decltype(nm_tag)::value;
            T cfg[NMc][kK], Wm[NMc];
            for (int m = 0; m < NMc; ++m) { Wm[m] = W[m * U + u]; for (int k = 0; k < kK; ++k) cfg[m][k] = mcfg[m * kK + k]; }
            PatchOut<T, NMc> out; out.zero();
            for (int row = 0; row < P; ++row)
                for (int col = 0; col < P; ++col) {
                    T gxk[kK], gyk[kK];
                    for (int k = 0; k < kK; ++k) { gxk[k] = axis_factor<T>(col, s.cx[k], s.w[k]); gyk[k] = axis_factor<T>(row, s.cy[k], s.w[k]); }
                    pixel_accumulate<T, NMc, true>(value[(u * P + row) * P + col], gxk, gyk, col, row, s, cfg, rate, log_rate, O, off_s, off_w, Wm, out);
                }
            for (int m = 0; m < NMc; ++m) logp[m * U + u] = out.logp[m];
            g_b[u] = out.g_b; g_rate[u] = out.g_rate;
            for (int k = 0; k < kK; ++k) { g_h[k * U + u] = out.g_h[k]; g_w[k * U + u] = out.g_w[k]; g_x[k * U + u] = out.g_x[k]; g_y[k * U + u] = out.g_y[k]; }
        };
        if (NM == 1) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, kM>{});
    }
}

extern "C" {
void hc_ksmogn_f64(int64_t U, int P, int O, int NM, const double* height, const double* width, const double* x, const double* y,
                   const double* background, double gain, const double* mcfg, const double* W, const double* value, const double* target,
                   const double* off_s, const double* off_w, double* logp, double* g_h, double* g_w, double* g_x, double* g_y, double* g_b, double* g_rate) {
    ksmogn_host<double>(U, P, O, NM, height, width, x, y, background, gain, mcfg, W, value, target, off_s, off_w, logp, g_h, g_w, g_x, g_y, g_b, g_rate);
}
void hc_ksmogn_f32(int64_t U, int P, int O, int NM, const float* height, const float* width, const float* x, const float* y,
                   const float* background, float gain, const float* mcfg, const float* W, const float* value, const float* target,
                   const float* off_s, const float* off_w, float* logp, float* g_h, float* g_w, float* g_x, float* g_y, float* g_b, float* g_rate) {
    ksmogn_host<float>(U, P, O, NM, height, width, x, y, background, gain, mcfg, W, value, target, off_s, off_w, logp, g_h, g_w, g_x, g_y, g_b, g_rate);
}
double hc_digamma_f64(double x) { return digamma<double>(x); }
float hc_digamma_f32(float x) { return digamma<float>(x); }
double hc_std_gamma_grad_f64(double alpha, double x) { return std_gamma_grad<double>(alpha, x); }
float hc_std_gamma_grad_f32(float alpha, float x) { return std_gamma_grad<float>(alpha, x); }
double hc_beta_grad_f64(double x, double alpha, double total) { return beta_grad<double>(x, alpha, total); }
float hc_beta_grad_f32(float x, float alpha, float total) { return beta_grad<float>(x, alpha, total); }
void hc_philox(uint64_t seed, uint64_t stream, uint64_t offset, int n, uint32_t* out) {
    Philox rng(seed, stream, offset);
    for (int i = 0; i < n; ++i) out[i] = rng.next();
}
void hc_sample_gamma_f64(uint64_t seed, uint64_t stream, double alpha, int n, double* out) {
    for (int i = 0; i < n; ++i) { Philox rng(seed, stream, (uint64_t)i * 64); out[i] = sample_std_gamma<double>(rng, alpha); }
}
}
