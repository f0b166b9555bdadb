// Device math shared by the cosmos kernels (sm_100a).
//
// * Real<T>: thin wrappers so every kernel is written once and instantiated for float (the
//   production path) and double (the CLI's dtype, main.py:428; also the exact-parity check).
// * digamma / lgamma: closed forms used by the likelihood and by the log-densities.
// * std_gamma_grad / beta_grad: the implicit reparameterisation gradients.  The reference gets
//   these from torch (ATen ``_standard_gamma_grad`` / ``_dirichlet_grad``, reached through
//   ``Gamma.rsample`` / ``Beta.rsample`` in models/cosmos.py:342-368,408-462).  ATen's versions are
//   piecewise *approximations* (Taylor / Rice saddle point / fitted rationals), so gradient parity
//   with the reference requires the same published algorithm and coefficient tables
//   (torch 2.x, ATen/native/Distributions.h, BSD-3); they are restated here.
// * Philox4x32-10 + Marsaglia-Tsang: in-kernel sampling for the production (non-replay) path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace tq {

constexpr int kK = 2;          // spots per image (cosmos K, BASELINE configs use 2)
constexpr int kM = 1 << kK;    // enumerated spot-presence configurations
constexpr int kTheta = kK + 1; // theta states
constexpr int kZ = 2;          // z states (S = 1)

#ifdef __CUDACC__
#define TQ_DEV __device__ __forceinline__
#define TQ_HD __host__ __device__ __forceinline__
#define TQ_HD_NOINLINE static __host__ __device__ __noinline__   // static: header-defined, several translation units
#else
#define TQ_DEV inline
#define TQ_HD inline
#define TQ_HD_NOINLINE inline
#endif

// log Gamma(x) for x > 0 in double: recurrence up to x >= 10 then the Stirling series (relative
// error ~1e-15); several times cheaper on the device than the general-purpose ::lgamma.
TQ_HD double lgamma_pos(double x) {
    double prod = 1.0;
    while (x < 10.0) {
        prod *= x;
        x += 1.0;
    }
    const double ix = 1.0 / x, z = ix * ix;
    double r = 6.41025641025641025641e-3;                    // 1/156
    r = 1.91752691752691752692e-3 - z * r;                   // 691/360360
    r = 8.41750841750841750842e-4 - z * r;                   // 1/1188
    r = 5.95238095238095238095e-4 - z * r;                   // 1/1680
    r = 7.93650793650793650794e-4 - z * r;                   // 1/1260
    r = 2.77777777777777777778e-3 - z * r;                   // 1/360
    r = 8.33333333333333333333e-2 - z * r;                   // 1/12
    const double lg = (x - 0.5) * ::log(x) - x + 0.91893853320467274178 + ix * r;
    return prod == 1.0 ? lg : lg - ::log(prod);
}

template <typename T> struct Real;
template <> struct Real<float> {
    static TQ_HD float log(float x) { return logf(x); }
    static TQ_HD float exp(float x) { return expf(x); }
    static TQ_HD float log1p(float x) { return log1pf(x); }
    // MUFU-based forms for log-sum-exp style use (arguments <= 0 resp. sums in [1, n]): absolute error ~1e-7
#ifdef __CUDA_ARCH__
    static __device__ __forceinline__ float exp_fast(float x) { return __expf(x); }
    static __device__ __forceinline__ float log_fast(float x) { return __logf(x); }
#else
    static inline float exp_fast(float x) { return expf(x); }
    static inline float log_fast(float x) { return logf(x); }
#endif
    static TQ_HD float lgamma(float x) { return lgammaf(x); }
    static TQ_HD float sqrt(float x) { return sqrtf(x); }
    static TQ_HD float pow(float x, float y) { return powf(x, y); }
    static TQ_HD float floor(float x) { return floorf(x); }
    static TQ_HD float max(float a, float b) { return fmaxf(a, b); }
    static TQ_HD float min(float a, float b) { return fminf(a, b); }
    static TQ_HD float abs(float a) { return fabsf(a); }
    static TQ_HD float inf() { return INFINITY; }
    static TQ_HD float eps() { return 1.1920928955078125e-07f; }
    static TQ_HD float tiny() { return 1.17549435e-38f; }
};
template <> struct Real<double> {
    static TQ_HD double log(double x) { return ::log(x); }
    static TQ_HD double exp(double x) { return ::exp(x); }
    static TQ_HD double log1p(double x) { return ::log1p(x); }
    static TQ_HD double exp_fast(double x) { return ::exp(x); }
    static TQ_HD double log_fast(double x) { return ::log(x); }
    static TQ_HD double lgamma(double x) { return lgamma_pos(x); }
    static TQ_HD double sqrt(double x) { return ::sqrt(x); }
    static TQ_HD double pow(double x, double y) { return ::pow(x, y); }
    static TQ_HD double floor(double x) { return ::floor(x); }
    static TQ_HD double max(double a, double b) { return fmax(a, b); }
    static TQ_HD double min(double a, double b) { return fmin(a, b); }
    static TQ_HD double abs(double a) { return fabs(a); }
    static TQ_HD double inf() { return (double)INFINITY; }
    static TQ_HD double eps() { return 2.220446049250313e-16; }
    static TQ_HD double tiny() { return 2.2250738585072014e-308; }
};

// ---- warp helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// ---- digamma --------------------------------------------------------------------------------
// psi(x) for x > 0: recurrence up to x >= 10, then the asymptotic series (same construction as
// ATen's digamma_one, Cephes-derived, so the reparameterisation gradients below agree with torch).
template <typename T> TQ_HD_NOINLINE T digamma(T x) {
    if (x == T(0)) return Real<T>::inf();
    // recurrence psi(x) = psi(x + n) - sum_{i<n} 1/(x + i) up to x + n >= 10; written as a fixed-trip
    // predicated loop so that the (independent) reciprocals pipeline instead of forming a serial chain
    T acc = T(0);
    if (x < T(10)) {
        const T x0 = x;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const T xi = x0 + T(i);
            if (xi < T(10)) {
                acc -= T(1) / xi;
                x = xi + T(1);
            }
        }
    }
    if (x == T(10)) return acc + T(2.25175258906672110764);
    T z = T(1) / (x * x);
    // z * (A0 z^6 .. A6) evaluated by Horner
    T poly = T(8.33333333333333333333E-2);
    poly = poly * z + T(-2.10927960927960927961E-2);
    poly = poly * z + T(7.57575757575757575758E-3);
    poly = poly * z + T(-4.16666666666666666667E-3);
    poly = poly * z + T(3.96825396825396825397E-3);
    poly = poly * z + T(-8.33333333333333333333E-3);
    poly = poly * z + T(8.33333333333333333333E-2);
    return acc + Real<T>::log(x) - T(0.5) / x - z * poly;
}

// lgamma(x) and digamma(x) together for x >= 10 from one log and one reciprocal (Stirling series,
// relative error ~1e-15); the guide's Beta/Gamma concentrations are in this range almost always.
TQ_HD void lgamma_digamma_large(double x, double& lg, double& psi) {
    const double lx = ::log(x), ix = 1.0 / x, z = ix * ix;
    double r = 6.41025641025641025641e-3;
    r = 1.91752691752691752692e-3 - z * r;
    r = 8.41750841750841750842e-4 - z * r;
    r = 5.95238095238095238095e-4 - z * r;
    r = 7.93650793650793650794e-4 - z * r;
    r = 2.77777777777777777778e-3 - z * r;
    r = 8.33333333333333333333e-2 - z * r;
    lg = (x - 0.5) * lx - x + 0.91893853320467274178 + ix * r;
    double p = 8.33333333333333333333E-2;
    p = p * z + -2.10927960927960927961E-2;
    p = p * z + 7.57575757575757575758E-3;
    p = p * z + -4.16666666666666666667E-3;
    p = p * z + 3.96825396825396825397E-3;
    p = p * z + -8.33333333333333333333E-3;
    p = p * z + 8.33333333333333333333E-2;
    psi = lx - 0.5 * ix - z * p;
}
// same with log(x) and 1/x supplied by the caller (x >= 10)
TQ_HD void lgamma_digamma_known(double x, double lx, double ix, double& lg, double& psi) {
    const double z = ix * ix;
    double r = 6.41025641025641025641e-3;
    r = 1.91752691752691752692e-3 - z * r;
    r = 8.41750841750841750842e-4 - z * r;
    r = 5.95238095238095238095e-4 - z * r;
    r = 7.93650793650793650794e-4 - z * r;
    r = 2.77777777777777777778e-3 - z * r;
    r = 8.33333333333333333333e-2 - z * r;
    lg = (x - 0.5) * lx - x + 0.91893853320467274178 + ix * r;
    double p = 8.33333333333333333333E-2;
    p = p * z + -2.10927960927960927961E-2;
    p = p * z + 7.57575757575757575758E-3;
    p = p * z + -4.16666666666666666667E-3;
    p = p * z + 3.96825396825396825397E-3;
    p = p * z + -8.33333333333333333333E-3;
    p = p * z + 8.33333333333333333333E-2;
    psi = lx - 0.5 * ix - z * p;
}
TQ_HD void lgamma_digamma(double x, double& lg, double& psi) {
    if (x > 10.0) { lgamma_digamma_large(x, lg, psi); return; }
    lg = lgamma_pos(x);
    psi = digamma<double>(x);
}

// ---- reparameterisation gradient of a standard Gamma(alpha) draw x: d x / d alpha ------------
template <typename T> TQ_HD_NOINLINE T std_gamma_grad(T alpha, T x) {
    using R = Real<T>;
    if (x < T(0.8)) {
        // Taylor series of the lower incomplete gamma function around x = 0
        T numer = T(1), denom = alpha;
        T s1 = numer / denom, s2 = numer / (denom * denom);
#pragma unroll
        for (int i = 1; i <= 5; ++i) {
            numer *= -x / T(i);
            denom += T(1);
            s1 += numer / denom;
            s2 += numer / (denom * denom);
        }
        const T pow_x_alpha = R::pow(x, alpha);
        const T pdf = R::pow(x, alpha - T(1)) * R::exp(-x);
        const T cdf = pow_x_alpha * s1;
        const T cdf_alpha = (R::log(x) - digamma(alpha)) * cdf - pow_x_alpha * s2;
        const T res = -cdf_alpha / pdf;
        return (res != res) ? T(0) : res;
    }
    if (alpha > T(8)) {
        // Rice saddle-point expansion; a Taylor patch removes the singularity at x = alpha
        if (T(0.9) * alpha <= x && x <= T(1.1) * alpha) {
            const T n1 = T(1) + T(24) * alpha * (T(1) + T(12) * alpha);
            const T n2 = T(1440) * (alpha * alpha) + T(6) * x * (T(53) - T(120) * x)
                         - T(65) * x * x / alpha + alpha * (T(107) + T(3600) * x);
            const T den = T(1244160) * (alpha * alpha) * (alpha * alpha);
            return n1 * n2 / den;
        }
        const T den = R::sqrt(T(8) * alpha);
        const T t2 = den / (alpha - x);
        const T t3b = x - alpha - alpha * R::log(x / alpha);
        const T t3 = T(1) / (t3b * R::sqrt(t3b));   // t3b^(-3/2)
        const T t23 = (x < alpha) ? t2 - t3 : t2 + t3;
        const T t1 = R::log(x / alpha) * t23 - R::sqrt(T(2) / alpha) * (alpha + x) / ((alpha - x) * (alpha - x));
        const T stirling = T(1) + T(1) / (T(12) * alpha) * (T(1) + T(1) / (T(24) * alpha));
        return -stirling * (x * t1) / den;
    }
    // bivariate rational fit in (log(x/alpha), log alpha)
    const T u = R::log(x / alpha);
    const T v = R::log(alpha);
    const T c[3][8] = {
        {T(0.16009398), T(-0.094634809), T(0.025146376), T(-0.0030648343), T(1), T(0.32668115), T(0.10406089), T(0.0014179084)},
        {T(0.53487893), T(0.1298071), T(0.065735949), T(-0.0015649758), T(0.16639465), T(0.020070113), T(-0.0035938915), T(-0.00058392623)},
        {T(0.040121004), T(-0.0065914022), T(-0.0026286047), T(-0.0013441777), T(0.017050642), T(-0.0021309326), T(0.00085092367), T(-1.5247877e-07)},
    };
    T cv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cv[i] = c[0][i] + u * (c[1][i] + u * c[2][i]);
    const T p = cv[0] + v * (cv[1] + v * (cv[2] + v * cv[3]));
    const T q = cv[4] + v * (cv[5] + v * (cv[6] + v * cv[7]));
    return R::exp(p / q);
}

// std_gamma_grad for alpha > 8, x >= 0.8 with lxa = log(x / alpha) and ia = 1 / alpha supplied
// (the Rice expansion and its Taylor patch, same formulas as above)
TQ_HD double std_gamma_grad_large(double alpha, double x, double lxa, double ia) {
    if (0.9 * alpha <= x && x <= 1.1 * alpha) {
        const double n1 = 1.0 + 24.0 * alpha * (1.0 + 12.0 * alpha);
        const double n2 = 1440.0 * (alpha * alpha) + 6.0 * x * (53.0 - 120.0 * x) - 65.0 * x * x * ia + alpha * (107.0 + 3600.0 * x);
        const double ia2 = ia * ia;
        return n1 * n2 * (ia2 * ia2) * (1.0 / 1244160.0);
    }
    const double den = sqrt(8.0 * alpha);
    const double iamx = 1.0 / (alpha - x);
    const double t2 = den * iamx;
    const double t3b = x - alpha - alpha * lxa;
    const double t3 = 1.0 / (t3b * sqrt(t3b));
    const double t23 = (x < alpha) ? t2 - t3 : t2 + t3;
    const double t1 = lxa * t23 - sqrt(2.0 * ia) * (alpha + x) * (iamx * iamx);
    const double stirling = 1.0 + ia * (1.0 / 12.0) * (1.0 + ia * (1.0 / 24.0));
    return -stirling * (x * t1) / den;
}

// ---- scaled reparameterisation gradient of a Beta(alpha, total-alpha) draw x wrt alpha --------
//   -(d/dalpha cdf) / pdf / (1 - x); total is passed so that the 2-Dirichlet form
//   dx/dc1 = (1-x) g(x, c1, tot),  dx/dc0 = -x g(1-x, c0, tot)   follows (torch dirichlet.py backward).
template <typename T> TQ_HD_NOINLINE T beta_grad_alpha_small(T x, T alpha, T beta) {
    using R = Real<T>;
    const T factor = digamma(alpha) - digamma(alpha + beta) - R::log(x);
    T numer = T(1);
    T series = numer / alpha * (factor + T(1) / alpha);
#pragma unroll
    for (int i = 1; i <= 10; ++i) {
        numer *= (T(i) - beta) * x / T(i);
        const T denom = alpha + T(i);
        series += numer / denom * (factor + T(1) / denom);
    }
    const T res = x * R::pow(T(1) - x, -beta) * series;
    return (res != res) ? T(0) : res;
}

template <typename T> TQ_HD_NOINLINE T beta_grad_beta_small(T x, T alpha, T beta) {
    using R = Real<T>;
    const T factor = digamma(alpha + beta) - digamma(beta);
    T numer = T(1), betas = T(1), dbetas = T(0), series = factor / alpha;
#pragma unroll
    for (int i = 1; i <= 8; ++i) {
        numer *= -x / T(i);
        dbetas = dbetas * (beta - T(i)) + betas;
        betas = betas * (beta - T(i));
        series += numer / (alpha + T(i)) * (dbetas + factor * betas);
    }
    const T res = -R::pow(T(1) - x, T(1) - beta) * series;
    return (res != res) ? T(0) : res;
}

template <typename T> TQ_HD_NOINLINE T beta_grad_alpha_mid(T x, T alpha, T beta) {
    using R = Real<T>;
    const T total = alpha + beta;
    const T mean = alpha / total;
    const T sd = R::sqrt(alpha * beta / (total + T(1))) / total;
    if (mean - T(0.1) * sd <= x && x <= mean + T(0.1) * sd) {
        const T b2 = beta * beta;
        const T poly = T(47) * x * b2 * b2 + alpha * (
                           (T(43) + T(20) * (T(16) + T(27) * beta) * x) * b2 * beta + alpha * (
                           T(3) * (T(59) + T(180) * beta - T(90) * x) * b2 + alpha * (
                           (T(453) + T(1620) * beta * (T(1) - x) - T(455) * x) * beta + alpha * (
                           T(8) * (T(1) - x) * (T(135) * beta - T(11))))));
        const T pn = (T(1) + T(12) * alpha) * (T(1) + T(12) * beta) / (total * total);
        const T pd = T(12960) * alpha * alpha * alpha * b2 * (T(1) + T(12) * total);
        return pn / (T(1) - x) * poly / pd;
    }
    const T prefactor = -x / R::sqrt(T(2) * alpha * beta / total);
    const T stirling = (T(1) + T(1) / (T(12) * alpha) + T(1) / (T(288) * alpha * alpha))
                     * (T(1) + T(1) / (T(12) * beta) + T(1) / (T(288) * beta * beta))
                     / (T(1) + T(1) / (T(12) * total) + T(1) / (T(288) * total * total));
    const T t1n = T(2) * (alpha * alpha) * (x - T(1)) + alpha * beta * (x - T(1)) - x * (beta * beta);
    const T axbx = alpha * (x - T(1)) + beta * x;
    const T t1d = R::sqrt(T(2) * alpha / beta) * (total * R::sqrt(total)) * axbx * axbx;
    const T t1 = t1n / t1d;
    const T t2 = T(0.5) * R::log(alpha / (total * x));
    const T t3 = R::sqrt(T(8) * alpha * beta / total) / (beta * x + alpha * (x - T(1)));
    const T t4b = beta * R::log(beta / (total * (T(1) - x))) + alpha * R::log(alpha / (total * x));
    const T t4 = T(1) / (t4b * R::sqrt(t4b));   // t4b^(-3/2)
    return stirling * prefactor * (t1 + t2 * (t3 + (x < mean ? t4 : -t4)));
}

template <typename T> TQ_HD_NOINLINE T beta_grad(T x, T alpha, T total) {
    using R = Real<T>;
    const T beta = total - alpha;
    const T boundary = total * x * (T(1) - x);
    if (x <= T(0.5) && boundary < T(2.5)) return beta_grad_alpha_small(x, alpha, beta);
    if (x >= T(0.5) && boundary < T(0.75)) return -beta_grad_beta_small(T(1) - x, beta, alpha);
    if (alpha > T(6) && beta > T(6)) return beta_grad_alpha_mid(x, alpha, beta);
    // rational correction to an analytic approximation (coefficients: torch Distributions.h)
    const T c[2][3][3][4] = {
        {{{T(1.003668233), T(-0.01061107488), T(-0.0657888334), T(0.01201642863)},
          {T(0.6336835991), T(-0.3557432599), T(0.05486251648), T(-0.001465281033)},
          {T(-0.03276231906), T(0.004474107445), T(0.002429354597), T(-0.0001557569013)}},
         {{T(0.221950385), T(-0.3187676331), T(0.01799915743), T(0.01074823814)},
          {T(-0.2951249643), T(0.06219954479), T(0.01535556598), T(0.001550077057)},
          {T(0.02155310298), T(0.004170831599), T(0.001292462449), T(6.976601077e-05)}},
         {{T(-0.05980841433), T(0.008441916499), T(0.01085618172), T(0.002319392565)},
          {T(0.02911413504), T(0.01400243777), T(-0.002721828457), T(0.000751041181)},
          {T(0.005900514878), T(-0.001936558688), T(-9.495446725e-06), T(5.385558597e-05)}}},
        {{{T(1), T(-0.02924021934), T(-0.04438342661), T(0.007285809825)},
          {T(0.6357567472), T(-0.3473456711), T(0.05454656494), T(-0.002407477521)},
          {T(-0.03301322327), T(0.004845219414), T(0.00231480583), T(-0.0002307248149)}},
         {{T(0.5925320577), T(-0.1757678135), T(0.01505928619), T(0.000564515273)},
          {T(0.1014815858), T(-0.06589186703), T(0.01272886114), T(-0.0007316646956)},
          {T(-0.007258481865), T(0.001096195486), T(0.0003934994223), T(-4.12701925e-05)}},
         {{T(0.06469649321), T(-0.0236701437), T(0.002902096474), T(-5.896963079e-05)},
          {T(0.001925008108), T(-0.002869809258), T(0.0008000589141), T(-6.063713228e-05)},
          {T(-0.0003477407336), T(6.959756487e-05), T(1.097287507e-05), T(-1.650964693e-06)}}},
    };
    const T u = R::log(x);
    const T a = R::log(alpha) - u;
    const T b = R::log(total) - a;
    const T pu[3] = {T(1), u, u * u};
    const T pa[3] = {T(1), a, a * a};
    T p = T(0), q = T(0);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const T ua = pu[i] * pa[j];
            p += ua * (c[0][i][j][0] + b * (c[0][i][j][1] + b * (c[0][i][j][2] + b * c[0][i][j][3])));
            q += ua * (c[1][i][j][0] + b * (c[1][i][j][1] + b * (c[1][i][j][2] + b * c[1][i][j][3])));
        }
    const T approx = x * (digamma(total) - digamma(alpha)) / beta;
    return p / q * approx;
}

// Both reparameterisation gradients of a Beta(c1, c0) draw x at once:
//   g1 = beta_grad(x, c1, c1 + c0),  g0 = beta_grad(1 - x, c0, c1 + c0)
// In the "both concentrations large" regime (the usual one for the cosmos guide: sizes of 100-1000)
// the two Rice expansions share their logarithms and square roots.
template <typename T> TQ_HD_NOINLINE void beta_grad_pair(T x, T c1, T c0, T& g1, T& g0) {
    using R = Real<T>;
    const T total = c1 + c0;
    const T boundary = total * x * (T(1) - x);
    const bool mid = (c1 > T(6)) && (c0 > T(6)) && !(boundary < T(2.5)) ;
    const T mean = c1 / total;
    const T sd = R::sqrt(c1 * c0 / (total + T(1))) / total;
    const bool patch = (mean - T(0.1) * sd <= x) && (x <= mean + T(0.1) * sd);
    if (!mid || patch) {
        g1 = beta_grad<T>(x, c1, total);
        g0 = beta_grad<T>(T(1) - x, c0, total);
        return;
    }
    const T y = T(1) - x;
    const T sab = R::sqrt(c1 * c0 / total);                 // sqrt(alpha beta / total)
    const T stirling = (T(1) + T(1) / (T(12) * c1) + T(1) / (T(288) * c1 * c1))
                     * (T(1) + T(1) / (T(12) * c0) + T(1) / (T(288) * c0 * c0))
                     / (T(1) + T(1) / (T(12) * total) + T(1) / (T(288) * total * total));
    const T axbx = c0 * x - c1 * y;                        // alpha (x - 1) + beta x   (call 1); call 2 has the opposite sign
    const T t15 = total * R::sqrt(total) * axbx * axbx;
    const T rab = R::sqrt(c1 / c0);                         // sqrt(alpha / beta)
    const T L1 = R::log(c1 / (total * x)), L0 = R::log(c0 / (total * y));
    const T t4b = c0 * L0 + c1 * L1;
    const T t4 = T(1) / (t4b * R::sqrt(t4b));
    const T s8 = T(2.8284271247461900976) * sab;           // sqrt(8 alpha beta / total)
    const T r2 = T(1.4142135623730950488);
    // call 1: (x, alpha = c1, beta = c0)
    {
        const T t1 = (-T(2) * c1 * c1 * y - c1 * c0 * y - x * c0 * c0) / (r2 * rab * t15);
        const T t3 = s8 / axbx;
        g1 = stirling * (-x / (r2 * sab)) * (t1 + T(0.5) * L1 * (t3 + (x < mean ? t4 : -t4)));
    }
    // call 2: (1 - x, alpha = c0, beta = c1)
    {
        const T t1 = (-T(2) * c0 * c0 * x - c1 * c0 * x - y * c1 * c1) / (r2 / rab * t15);
        const T t3 = -s8 / axbx;
        g0 = stirling * (-y / (r2 * sab)) * (t1 + T(0.5) * L0 * (t3 + (y < c0 / total ? t4 : -t4)));
    }
}

// ---- Philox4x32-10 counter RNG -----------------------------------------------------------------
#ifndef TQ_PHILOX_UNROLL
// Rounds unrolled per trip.  Two, not ten: the guide-site kernel inlines the block generator half a dozen times inside 114 KB of
// SASS, and on a trained model -- where its warps walk most of that code -- instruction fetch is its largest stall reason.
// B200, one 8-GPU rank's C3 shard, site kernel at the initial point / after 1000 / 3000 iterations: 191 / 246 / 499 us
// unrolled by ten, 190 / 236 / 438 by two, 194 / 242 / 444 by one, 190 / 239 / 449 as one out-of-line function
// (profiles/r2s2_sites_icache.md).  Same arithmetic, same bits.
#define TQ_PHILOX_UNROLL 2
#endif
constexpr int kPhiloxUnroll = TQ_PHILOX_UNROLL;
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t out[4];
    int have;
    TQ_HD Philox(uint64_t seed, uint64_t stream, uint64_t offset) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = (uint32_t)offset; ctr[1] = (uint32_t)(offset >> 32);
        ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
        have = 0;
    }
    TQ_HD void round(uint32_t (&c)[4], const uint32_t (&k)[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // the next block of four words; advances the counter
    TQ_HD void block(uint32_t (&c)[4]) {
        c[0] = ctr[0]; c[1] = ctr[1]; c[2] = ctr[2]; c[3] = ctr[3];
        uint32_t k[2] = {key[0], key[1]};
#pragma unroll kPhiloxUnroll
        for (int r = 0; r < 10; ++r) {
            round(c, k);
            k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
        }
        if (++ctr[0] == 0) ++ctr[1];
    }
    TQ_HD void refill() {
        uint32_t c[4];
        block(c);
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
        have = 4;
    }
    TQ_HD uint32_t next() {
        if (have == 0) refill();
        const uint32_t r = out[0];   // shift instead of out[--have]: keeps the state in registers
        out[0] = out[1]; out[1] = out[2]; out[2] = out[3];
        --have;
        return r;
    }
    // uniform in (0, 1]
    static TQ_HD float to_uniform(uint32_t w) { return ((float)(w >> 8) + 1.0f) * (1.0f / 16777216.0f); }
    TQ_HD float uniform() { return to_uniform(next()); }
    TQ_HD double uniform_d() {
        const uint64_t hi = next(), lo = next();
        return ((double)(((hi << 32) | lo) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    }
    TQ_HD float normal() {  // Box-Muller, one value per call (the partner is discarded)
        const float u1 = uniform(), u2 = uniform();
#ifdef __CUDA_ARCH__
        // MUFU forms (lg2 / sqrt / cos): absolute errors ~1e-6, irrelevant for a random draw
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
        return r * __cosf(6.283185307179586f * u2);
#else
        return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
#endif
    }
    // One whole block as TWO (standard normal, uniform in (0, 1]) pairs: both branches of Box-Muller on words 0 and 1,
    // the uniforms from words 2 and 3.  A Marsaglia-Tsang trial consumes exactly one such pair, so a block serves a Beta
    // site's two gamma draws (or a draw and its retry): 9 blocks per unit instead of 15, and none of next()'s bookkeeping.
    // Independent of the word buffer behind next().
    TQ_HD void normal_uniform_pairs(float& n0, float& u0, float& n1, float& u1) {
        uint32_t c[4];
        block(c);
        const float a = to_uniform(c[0]), b = to_uniform(c[1]);
        float r, sn, cs;
#ifdef __CUDA_ARCH__
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(a)));
        __sincosf(6.283185307179586f * b, &sn, &cs);
#else
        r = sqrtf(-2.0f * logf(a));
        sn = sinf(6.283185307179586f * b); cs = cosf(6.283185307179586f * b);
#endif
        n0 = r * cs; n1 = r * sn;
        u0 = to_uniform(c[2]); u1 = to_uniform(c[3]);
    }
};

// (normal, uniform) pairs for Marsaglia-Tsang trials, two per Philox block: the second waits here for the next trial --
// of the same draw after a rejection, or of the next draw of the same site (a Beta variate is two gamma draws)
struct GammaTrials {
    float n0, u0, n1, u1;
    int have = 0;
    // fetch a block ahead of its use (it depends on nothing but the counter: a kernel issues it under the latency of its
    // parameter loads); the pairs are consumed in the same order with or without it
    TQ_HD void preload(Philox& rng) {
        if (have == 0) { rng.normal_uniform_pairs(n0, u0, n1, u1); have = 2; }
    }
    TQ_HD void draw(Philox& rng, float& xn, float& u) {
        preload(rng);
        if (have == 2) { xn = n0; u = u0; } else { xn = n1; u = u1; }
        --have;
    }
};

// fp32 inline variant of the sampler below for the per-site kernels: inlining keeps the Philox state
// in registers (the out-of-line generic version takes it by reference, i.e. through local memory).
// Returns a double: the Marsaglia-Tsang core runs in fp32 (its value d*v is O(alpha)), but the alpha < 1 boost
// u^(1/alpha) spans hundreds of orders of magnitude and is applied in double -- in fp32 it underflows to an exact 0 with
// probability ~(1e-38)^alpha, which puts the guide sample ON its clamp (seen after ~3000 SVI iterations on absent spots,
// whose Beta concentrations drift below 1: x = -7.5 exactly -> log1p(-1) = -inf downstream).
TQ_HD double sample_std_gamma_f32(Philox& rng, GammaTrials& trials, float alpha) {
#ifdef __CUDA_ARCH__
    // MUFU forms: the accept/reject comparison tolerates their ~1e-6 absolute error (a borderline trial
    // flips with probability ~1e-6; the accepted value d*v itself is exact arithmetic)
#define TQ_SLOG(x) __logf(x)
#define TQ_SRSQRT(x) rsqrtf(x)
#else
#define TQ_SLOG(x) logf(x)
#define TQ_SRSQRT(x) (1.0f / sqrtf(x))
#endif
    // the block of the first trial always comes first in the stream: a caller may have fetched it already (preload), and
    // the draw must not depend on that
    trials.preload(rng);
    double scale = 1.0;
    if (alpha < 1.0f) {
        // u^(1/alpha): the logarithm of a random number needs no more than fp32, the exponential needs double's range
        scale = exp((double)(logf((float)rng.uniform_d()) / alpha));
        alpha += 1.0f;
    }
    const float d = alpha - 1.0f / 3.0f;
    const float c = TQ_SRSQRT(9.0f * d);
    for (int it = 0; it < 64; ++it) {
        float xn, u;
        trials.draw(rng, xn, u);
        const float yv = 1.0f + c * xn;
        if (yv <= 0.0f) continue;       // (the paper redraws the normal alone; discarding the independent uniform with it is the same law)
        const float v = yv * yv * yv;
        const float xx = xn * xn;
        if (u < 1.0f - 0.0331f * xx * xx) return scale * (double)(d * v);
        if (TQ_SLOG(u) < 0.5f * xx + d * (1.0f - v + TQ_SLOG(v))) return scale * (double)(d * v);
    }
    return scale * (double)d;
#undef TQ_SLOG
#undef TQ_SRSQRT
}
TQ_HD double sample_std_gamma_f32(Philox& rng, float alpha) {
    GammaTrials trials;
    return sample_std_gamma_f32(rng, trials, alpha);
}

// Marsaglia & Tsang (2000) standard gamma sampler (doi:10.1145/358407.358414), alpha > 0.
template <typename T> TQ_HD_NOINLINE T sample_std_gamma(Philox& rng, T alpha) {
    using R = Real<T>;
    T scale = T(1);
    if (alpha < T(1)) {
        scale = R::pow(T(rng.uniform_d()), T(1) / alpha);
        alpha += T(1);
    }
    const T d = alpha - T(1) / T(3);
    const T c = T(1) / R::sqrt(T(9) * d);
    for (int it = 0; it < 64; ++it) {
        T xn, yv;
        do {
            xn = T(rng.normal());
            yv = T(1) + c * xn;
        } while (yv <= T(0));
        const T v = yv * yv * yv;
        const T u = T(rng.uniform());
        const T xx = xn * xn;
        if (u < T(1) - T(0.0331) * xx * xx) return scale * d * v;
        if (R::log(u) < T(0.5) * xx + d * (T(1) - v + R::log(v))) return scale * d * v;
    }
    return scale * d;  // unreachable in practice (acceptance > 95 % per trial)
}

}  // namespace tq
