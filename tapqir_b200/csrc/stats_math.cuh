// Inverse CDFs of the Gamma and Beta distributions in double precision (host + device: tests/hostcheck pins them against
// scipy on the CPU).  The textbook scheme: series / continued fraction (modified Lentz) for the regularised incomplete
// function, a closed-form starting point, Halley steps on  F(x) - p  with the density as derivative.
// Reference path replaced: scipy.stats.gamma / beta `.interval(CI)` behind stats.torch_to_scipy_dist
// (tapqir/utils/stats.py:262-293), called by cosmos.compute_params (models/cosmos.py:711-784).
#pragma once
#include <math.h>

#include "tq_math.cuh"

namespace tq {

constexpr double kTiny = 1e-300;
constexpr double kEps = 2.220446049250313e-16;

// regularised lower incomplete gamma P(a, x), a > 0, x >= 0; gln = lgamma(a)
TQ_HD_NOINLINE double gamma_p(double a, double x, double gln) {
    if (x <= 0.0) return 0.0;
    if (x < a + 1.0) {   // series
        double ap = a, del = 1.0 / a, sum = del;
        for (int i = 0; i < 200000; ++i) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * kEps) break;
        }
        return sum * exp(-x + a * log(x) - gln);
    }
    // continued fraction for Q (modified Lentz)
    double b = x + 1.0 - a, c = 1.0 / kTiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 200000; ++i) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < kTiny) d = kTiny;
        c = b + an / c;
        if (fabs(c) < kTiny) c = kTiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= kEps) break;
    }
    return 1.0 - exp(-x + a * log(x) - gln) * h;
}

// x with P(a, x) = p  (0 < p < 1)
TQ_HD_NOINLINE double gamma_p_inv(double p, double a) {
    const double gln = lgamma(a), a1 = a - 1.0;
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return fmax(100.0, a + 100.0 * sqrt(a));
    double x, lna1 = 0.0, afac = 0.0;
    if (a > 1.0) {   // Wilson-Hilferty from a rational normal quantile
        lna1 = log(a1);
        afac = exp(a1 * (lna1 - 1.0) - gln);
        const double pp = p < 0.5 ? p : 1.0 - p;
        const double t = sqrt(-2.0 * log(pp));
        double z = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (p < 0.5) z = -z;
        x = fmax(1e-3, a * pow(1.0 - 1.0 / (9.0 * a) - z / (3.0 * sqrt(a)), 3.0));
    } else {
        const double t = 1.0 - a * (0.253 + a * 0.12);
        x = p < t ? pow(p / t, 1.0 / a) : 1.0 - log(1.0 - (p - t) / (1.0 - t));
    }
    for (int j = 0; j < 40; ++j) {
        if (x <= 0.0) return 0.0;
        const double err = gamma_p(a, x, gln) - p;
        double t = a > 1.0 ? afac * exp(-(x - a1) + a1 * (log(x) - lna1)) : exp(-x + a1 * log(x) - gln);   // density
        if (!(t > 0.0)) break;
        const double u = err / t;
        t = u / (1.0 - 0.5 * fmin(1.0, u * (a1 / x - 1.0)));   // Halley
        x -= t;
        if (x <= 0.0) x = 0.5 * (x + t);
        if (fabs(t) < 1e-13 * x) break;
    }
    return x;
}

// continued fraction of the incomplete beta function (modified Lentz)
TQ_HD_NOINLINE double beta_cf(double a, double b, double x) {
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < kTiny) d = kTiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m < 200000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < kTiny) d = kTiny;
        c = 1.0 + aa / c;
        if (fabs(c) < kTiny) c = kTiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < kTiny) d = kTiny;
        c = 1.0 + aa / c;
        if (fabs(c) < kTiny) c = kTiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= kEps) break;
    }
    return h;
}

// regularised incomplete beta I_x(a, b); lbeta = lgamma(a) + lgamma(b) - lgamma(a + b)
TQ_HD_NOINLINE double beta_i(double a, double b, double x, double lbeta) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double bt = exp(a * log(x) + b * log1p(-x) - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * beta_cf(a, b, x) / a;
    return 1.0 - bt * beta_cf(b, a, 1.0 - x) / b;
}

// x with I_x(a, b) = p
TQ_HD_NOINLINE double beta_i_inv(double p, double a, double b) {
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return 1.0;
    const double a1 = a - 1.0, b1 = b - 1.0;
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    double x;
    if (a >= 1.0 && b >= 1.0) {
        const double pp = p < 0.5 ? p : 1.0 - p;
        const double t = sqrt(-2.0 * log(pp));
        double z = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (p < 0.5) z = -z;
        const double al = (z * z - 3.0) / 6.0;
        const double h = 2.0 / (1.0 / (2.0 * a - 1.0) + 1.0 / (2.0 * b - 1.0));
        const double w = z * sqrt(al + h) / h - (1.0 / (2.0 * b - 1.0) - 1.0 / (2.0 * a - 1.0)) * (al + 5.0 / 6.0 - 2.0 / (3.0 * h));
        x = a / (a + b * exp(2.0 * w));
    } else {
        const double lna = log(a / (a + b)), lnb = log(b / (a + b));
        const double t = exp(a * lna) / a, u = exp(b * lnb) / b;
        const double w = t + u;
        x = p < t / w ? pow(a * w * p, 1.0 / a) : 1.0 - pow(b * w * (1.0 - p), 1.0 / b);
    }
    for (int j = 0; j < 40; ++j) {
        if (x <= 0.0 || x >= 1.0) { x = fmin(fmax(x, 1e-300), 1.0 - 1e-16); }
        const double err = beta_i(a, b, x, lbeta) - p;
        double t = exp(a1 * log(x) + b1 * log1p(-x) - lbeta);   // density
        if (!(t > 0.0)) break;
        const double u = err / t;
        t = u / (1.0 - 0.5 * fmin(1.0, u * (a1 / x - b1 / (1.0 - x))));   // Halley
        x -= t;
        if (x <= 0.0) x = 0.5 * (x + t);
        if (x >= 1.0) x = 0.5 * (x + t + 1.0);
        if (fabs(t) < 1e-13 * x && j > 0) break;
    }
    return x;
}

}  // namespace tq
