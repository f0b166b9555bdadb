// Per-pixel arithmetic of the image likelihood: K rendered spots + background, Gamma noise
// convolved with the empirical offset distribution, for NM spot-presence configurations at once.
//
// Replaces (reference): distributions/util.py:44-64 (render), distributions/ksmogn.py:158-169
// (image, concentration) and ksmogn.py:187-238 (offset-marginalised log-density), plus what
// autograd derives from them.  Maths: SURVEY.md App. C.
//
// Host+device so the same code is exercised on the CPU by tests/hostcheck (no GPU needed) and
// inside the warp-per-patch kernels of ksmogn.cu / cosmos_step.cu.
#pragma once
#include "tq_math.cuh"

namespace tq {

constexpr int kMaxP = 32;  // widest patch the per-warp row/column tables hold

// One patch's spot parameters in absolute pixel coordinates.
template <typename T> struct PatchSpots {
    T h[kK];    // integrated intensity
    T w[kK];    // width
    T cx[kK];   // centre along x (last image axis): x_k + target_x
    T cy[kK];   // centre along y (second-to-last axis): y_k + target_y
    T b;        // background
};

// Reduced (over pixels) outputs of one patch.
template <typename T, int NM> struct PatchOut {
    T logp[NM];
    T g_b, g_h[kK], g_w[kK], g_x[kK], g_y[kK];
    T g_rate;  // d/d(1/gain), explicit + through concentration
    TQ_HD void zero() {
#pragma unroll
        for (int m = 0; m < NM; ++m) logp[m] = T(0);
        g_b = g_rate = T(0);
#pragma unroll
        for (int k = 0; k < kK; ++k) g_h[k] = g_w[k] = g_x[k] = g_y[k] = T(0);
    }
};

// Separable factors of spot k at pixel index i along one axis: exp(-(i-c)^2 / (2 w^2)).
template <typename T> TQ_HD T axis_factor(int i, T c, T w) {
    const T d = T(i) - c;
    return Real<T>::exp(-(d * d) / (T(2) * w * w));
}

// Accumulate one pixel into `out`.
//   D          observed pixel value
//   gxk, gyk   separable spot factors at this pixel's column / row
//   col, row   pixel indices (x, y)
//   mcfg       NM x K spot-presence table (floats: the operator seam accepts fractional m)
//   rate       1/gain, log_rate = log(1/gain)
//   off_s/off_w  O offset samples and log-weights
//   W          upstream weights per configuration (read only when BWD)
template <typename T, int NM, bool BWD>
TQ_HD void pixel_accumulate(T D, const T (&gxk)[kK], const T (&gyk)[kK], int col, int row,
                            const PatchSpots<T>& s, const T (&mcfg)[NM][kK], T rate, T log_rate,
                            int O, const T* __restrict__ off_s, const T* __restrict__ off_w,
                            const T (&W)[NM], PatchOut<T, NM>& out) {
    using R = Real<T>;
    const T two_pi = T(6.283185307179586476925);
    T shape[kK], mu[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        shape[k] = gxk[k] * gyk[k] / (two_pi * s.w[k] * s.w[k]);
        mu[k] = s.h[k] * shape[k];
    }
    T img[NM], a[NM], mx[NM], se[NM], sl[NM], sy[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        T v = s.b;
#pragma unroll
        for (int k = 0; k < kK; ++k) v += mcfg[m][k] * mu[k];
        img[m] = v;
        a[m] = v * rate;
        mx[m] = -R::inf();
        se[m] = sl[m] = sy[m] = T(0);
    }
    // pass 1: running maximum of  w_j + (a-1) log y_j - rate y_j  over the offsets with y_j > 0
    for (int j = 0; j < O; ++j) {
        const T y = D - off_s[j];
        if (y > T(0)) {
            const T l = R::log(y);
            const T base = off_w[j] - rate * y - l;
#pragma unroll
            for (int m = 0; m < NM; ++m) mx[m] = R::max(mx[m], base + a[m] * l);
        }
    }
    // pass 2: sum of exponentials and the softmax moments the gradient needs
    for (int j = 0; j < O; ++j) {
        const T y = D - off_s[j];
        if (y > T(0)) {
            const T l = R::log(y);
            const T base = off_w[j] - rate * y - l;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const T e = R::exp(base + a[m] * l - mx[m]);
                se[m] += e;
                if (BWD) {
                    sl[m] += e * l;
                    sy[m] += e * y;
                }
            }
        }
    }
    T g_img_sum = T(0), S[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) S[k] = T(0);
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        if (se[m] > T(0)) {
            out.logp[m] += a[m] * log_rate - R::lgamma(a[m]) + mx[m] + R::log(se[m]);
            if (BWD) {
                const T inv = T(1) / se[m];
                const T dLda = log_rate - digamma(a[m]) + sl[m] * inv;
                const T gi = W[m] * rate * dLda;
                out.g_rate += W[m] * (img[m] * dLda + img[m] - sy[m] * inv);
                g_img_sum += gi;
#pragma unroll
                for (int k = 0; k < kK; ++k) S[k] += mcfg[m][k] * gi;
            }
        } else {
            out.logp[m] = -R::inf();  // pixel at or below every offset (ksmogn.py:225-236 gives -inf)
        }
    }
    if (BWD) {
        out.g_b += g_img_sum;
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const T iw = T(1) / s.w[k];
            const T dx = T(col) - s.cx[k], dy = T(row) - s.cy[k];
            const T t = S[k] * mu[k];
            out.g_h[k] += S[k] * shape[k];
            out.g_x[k] += t * dx * iw * iw;
            out.g_y[k] += t * dy * iw * iw;
            out.g_w[k] += t * ((dx * dx + dy * dy) * iw * iw * iw - T(2) * iw);
        }
    }
}

}  // namespace tq
