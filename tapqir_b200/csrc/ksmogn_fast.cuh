// fp32 production form of the per-pixel likelihood arithmetic (same quantities as
// ksmogn_core.cuh::pixel_accumulate, which stays the fp64 / exact-parity form).
//
// The kernel is bound by instruction issue and by the MUFU pipe (16 lanes/SM/clk), so this form
//   * works in base 2: lg2.approx / ex2.approx are single MUFU ops, the ln2 factors are folded into
//     per-patch constants;
//   * caches the per-offset terms of a pixel in registers (OC <= 4 offset bins, one pass) instead
//     of two passes over the offsets;
//   * evaluates lgamma(a) and digamma(a) from ONE lg2(a) and ONE reciprocal through their Stirling
//     series (a = image/gain >= 4 directly, recurrence shift below that), SURVEY.md App. C.3;
//   * gets 1/a and 1/sum_exp from one rcp of their product;
//   * specialises the spot-presence table to the enumerated {0,1}^K one (no multiplies by m).
// Accuracy is pinned by tests/test_ksmogn_gpu.py and tests/test_step_gpu.py at the north-star
// tolerance (1e-5 of the fp64 oracle).
#pragma once
#include "ksmogn_core.cuh"

namespace tq {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float f_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
inline float f_lg2(float x) { return log2f(x); }
inline float f_ex2(float x) { return exp2f(x); }
inline float f_rcp(float x) { return 1.0f / x; }
#endif

// reciprocal without the denormal/overflow slow path of 1.0f / x (arguments are widths, heights: normal numbers)
TQ_HD float rcp_newton(float x) {
#ifdef __CUDA_ARCH__
    const float r = f_rcp(x);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}

constexpr float kLn2 = 0.69314718055994530942f;
constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kHalfLn2Pi = 0.91893853320467274178f;
constexpr float kNegInf = -3.0e38f;  // finite stand-in for -inf: keeps (v - max) free of NaNs

// per-patch constants of the fast form
struct FastConst {
    float rate;       // 1 / gain
    float rate2;      // rate * log2(e)
    float log_rate;   // log(rate)
    float gain;       // 1 / rate
};

// Stirling pieces for a >= 4:  lgamma(a) = (a - 1/2) ln a - a + ln(2 pi)/2 + r(a);  psi(a) = ln a - q(a)
TQ_HD void stirling(float ia, float& r, float& q) {
    const float ia2 = ia * ia;
    r = ia * (0.0833333333f - ia2 * (0.00277777778f - ia2 * 0.000793650794f));
    q = ia * (0.5f + ia * (0.0833333333f - ia2 * (0.00833333333f - ia2 * 0.00396825397f)));
}

// OC > 0: exactly OC offset bins cached in registers.  OC == 0: any O, two passes.
// SMALL: some configuration of this patch may have a = image/gain < 4 (needs the recurrence shift);
// a is smallest for the all-absent configuration, so the caller tests background/gain once per patch.
template <int NM, int OC, bool BWD, bool SMALL>
TQ_HD void pixel_accumulate_fast(float D, const float (&gxk)[kK], const float (&gyk)[kK], int col, int row,
                                 const PatchSpots<float>& s, const float (&norm)[kK],
                                 const FastConst& fc, int O, const float* __restrict__ off_s,
                                 const float* __restrict__ off_w2, const float (&W)[NM], const float (&Wr)[NM],
                                 PatchOut<float, NM>& out) {
    static_assert(NM == kM, "fast path is written for the enumerated 2^K table");
    float shape[kK], mu[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        shape[k] = gxk[k] * gyk[k] * norm[k];
        mu[k] = s.h[k] * shape[k];
    }
    float img[NM];
    img[0] = s.b;
    img[1] = s.b + mu[0];
    img[2] = s.b + mu[1];
    img[3] = img[1] + mu[1];

    constexpr int NC = OC > 0 ? OC : 1;
    float y[NC], l2[NC], b2[NC];
    bool any_ok = false;
    if (OC > 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float yy = D - off_s[j];
            const bool ok = yy > 0.0f;
            any_ok |= ok;
            y[j] = ok ? yy : 1.0f;
            l2[j] = f_lg2(y[j]);
            b2[j] = ok ? fmaf(-fc.rate2, yy, off_w2[j]) - l2[j] : kNegInf;
        }
    } else {
        for (int j = 0; j < O; ++j) any_ok |= (D - off_s[j]) > 0.0f;
    }
    if (!any_ok) {  // pixel at or below every offset: log-probability -inf (ksmogn.py:225-236)
#pragma unroll
        for (int m = 0; m < NM; ++m) out.logp[m] = -INFINITY;
        return;
    }
    float gi[NM];
    float g_img_sum = 0.0f;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const float a = img[m] * fc.rate;
        float mx = kNegInf, se = 0.0f, sl = 0.0f, sy = 0.0f;
        if (OC > 0) {
            float v[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                v[j] = fmaf(a, l2[j], b2[j]);
                mx = fmaxf(mx, v[j]);
            }
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float e = f_ex2(v[j] - mx);
                se += e;
                if (BWD) {
                    sl = fmaf(e, l2[j], sl);
                    sy = fmaf(e, y[j], sy);
                }
            }
        } else {
            for (int j = 0; j < O; ++j) {
                const float yy = D - off_s[j];
                if (yy > 0.0f) {
                    const float l = f_lg2(yy);
                    mx = fmaxf(mx, fmaf(a, l, fmaf(-fc.rate2, yy, off_w2[j]) - l));
                }
            }
            for (int j = 0; j < O; ++j) {
                const float yy = D - off_s[j];
                if (yy > 0.0f) {
                    const float l = f_lg2(yy);
                    const float e = f_ex2(fmaf(a, l, fmaf(-fc.rate2, yy, off_w2[j]) - l) - mx);
                    se += e;
                    if (BWD) {
                        sl = fmaf(e, l, sl);
                        sy = fmaf(e, yy, sy);
                    }
                }
            }
        }
        // lgamma / digamma of a through Stirling; shift by 4 when a is small
        float as = a, shift_log = 0.0f, shift_psi = 0.0f;
        if (SMALL) {
            if (a < 4.0f) {
                const float p01 = a * (a + 1.0f), p23 = (a + 2.0f) * (a + 3.0f);
                shift_log = kLn2 * f_lg2(p01 * p23);
                // 1/a + 1/(a+1) + 1/(a+2) + 1/(a+3)
                shift_psi = (2.0f * a + 1.0f) * f_rcp(p01) + (2.0f * a + 5.0f) * f_rcp(p23);
                as = a + 4.0f;
            }
        }
        const float inv = f_rcp(as * se);   // one reciprocal for 1/a and 1/se
        const float ia = inv * se, ise = inv * as;
        const float la = kLn2 * f_lg2(as);
        float r, q;
        stirling(ia, r, q);
        // a log(rate) - lgamma(a) = a log(rate) - [(as - 1/2) la - as + ln(2pi)/2 + r] + shift_log
        const float neg_lgamma = fmaf(0.5f - as, la, as) - kHalfLn2Pi - r + shift_log;
        const float lse = kLn2 * (mx + f_lg2(se));
        out.logp[m] += fmaf(a, fc.log_rate, neg_lgamma) + lse;
        if (BWD) {
            const float psi = la - q - shift_psi;
            const float dLda = fc.log_rate - psi + kLn2 * sl * ise;
            gi[m] = Wr[m] * dLda;   // Wr = W * rate
            out.g_rate = fmaf(W[m], fmaf(img[m], dLda + 1.0f, -sy * ise), out.g_rate);
            g_img_sum += gi[m];
        }
    }
    if (BWD) {
        out.g_b += g_img_sum;
        // spot gradients as moments of t = S_k mu_k about the spot centre; the 1/w^2, 1/w^3, 1/h
        // factors are applied once per patch (finish_spot_moments).  (Moments about the patch centre
        // would save three more ops but cancel for off-centre spots: 1e-5 errors measured.)
        const float S[kK] = {gi[1] + gi[3], gi[2] + gi[3]};
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const float dx = float(col) - s.cx[k], dy = float(row) - s.cy[k];
            const float t = S[k] * mu[k];
            out.g_h[k] += t;
            out.g_x[k] = fmaf(t, dx, out.g_x[k]);
            out.g_y[k] = fmaf(t, dy, out.g_y[k]);
            out.g_w[k] = fmaf(t, fmaf(dx, dx, dy * dy), out.g_w[k]);
        }
    }
}

// ---- packed pairs -------------------------------------------------------------------------------------------
// sm_100a has two-wide FP32 instructions (fma / add / sub / mul .f32x2 -> FFMA2 / FADD2 / FMUL2): one issue slot
// for two operations on a 64-bit register pair, scalar operands broadcast for free.  The likelihood kernel
// is bound by instruction issue, so its common case processes TWO pixels per lane in this form; MUFU ops
// and the max stay scalar.  Host build (tests/hostcheck): plain pairs of floats.
struct F2 { float x, y; };
TQ_HD F2 f2(float a) { return F2{a, a}; }
#ifdef __CUDA_ARCH__
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 d;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
#define TQ_F2_BINARY(name, op)                                                                             \
    __device__ __forceinline__ F2 name(F2 a, F2 b) {                                                       \
        F2 d;                                                                                              \
        asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n" op                    \
            ".rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"                                                \
            : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                              \
        return d;                                                                                          \
    }
TQ_F2_BINARY(add2, "add")
TQ_F2_BINARY(sub2, "sub")
TQ_F2_BINARY(mul2, "mul")
#undef TQ_F2_BINARY
#else
inline F2 fma2(F2 a, F2 b, F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline F2 add2(F2 a, F2 b) { return F2{a.x + b.x, a.y + b.y}; }
inline F2 sub2(F2 a, F2 b) { return F2{a.x - b.x, a.y - b.y}; }
inline F2 mul2(F2 a, F2 b) { return F2{a.x * b.x, a.y * b.y}; }
#endif
TQ_HD F2 lg2_2(F2 a) { return F2{f_lg2(a.x), f_lg2(a.y)}; }
TQ_HD F2 ex2_2(F2 a) { return F2{f_ex2(a.x), f_ex2(a.y)}; }
TQ_HD F2 rcp_2(F2 a) { return F2{f_rcp(a.x), f_rcp(a.y)}; }

// per-patch sums of the pair form; gradients w.r.t. the image are accumulated WITHOUT the factor 1/gain
// (applied once per patch by finish_pair)
struct PairOut {
    F2 logp[kM], g_b, g_rate, g_h[kK], g_w[kK], g_x[kK], g_y[kK];
    F2 common;   // part of the log-probability shared by all configurations (single-offset form)
    TQ_HD void zero() {
#pragma unroll
        for (int m = 0; m < kM; ++m) logp[m] = f2(0.0f);
        g_b = g_rate = common = f2(0.0f);
#pragma unroll
        for (int k = 0; k < kK; ++k) g_h[k] = g_w[k] = g_x[k] = g_y[k] = f2(0.0f);
    }
};

// Two pixels of one COLUMN (same x-factor gxn = gx * norm, same dx) at once.  Common case only: every
// pixel above every offset (no -inf handling) and a = image/gain >= 4 (no recurrence shift); the caller
// checks both per patch.  Same quantities as pixel_accumulate_fast<kM, OC, true, false>.
// gxh[k] = (column factor) x norm x height, dx / dx2 = column distance to the spot centre and its square: per-COLUMN
// values the kernel tabulates once per patch (they used to be re-derived for every pixel pair: 14 scalar FP32 operations
// per pair on the pipe that binds the kernel); gyk / dy: the row factors and distances of the two rows of the pair.
template <int OC>
TQ_HD void pixel_pair_accumulate_fast(F2 D, const float (&gxh)[kK], const F2 (&gyk)[kK], const float (&dx)[kK],
                                      const float (&dx2)[kK], const F2 (&dy)[kK], const PatchSpots<float>& s,
                                      const FastConst& fc, const float (&off_s)[OC], const float (&off_w2)[OC],
                                      const float (&W)[kM], PairOut& out) {
    F2 mu[kK], img[kM];
#pragma unroll
    for (int k = 0; k < kK; ++k) mu[k] = mul2(gyk[k], f2(gxh[k]));
    img[0] = f2(s.b);
    img[1] = add2(mu[0], f2(s.b));
    img[2] = add2(mu[1], f2(s.b));
    img[3] = add2(img[1], mu[1]);
    F2 y[OC], l2[OC], b2[OC];
#pragma unroll
    for (int j = 0; j < OC; ++j) {
        y[j] = sub2(D, f2(off_s[j]));
        l2[j] = lg2_2(y[j]);
        b2[j] = sub2(fma2(f2(-fc.rate2), y[j], f2(off_w2[j])), l2[j]);
    }
    // one offset bin (after merging identical support points: the simulated data, simulate.py:92,103): the
    // log-sum-exp is its single term, softmax weight 1 -- no exponentials, no log / reciprocal of the sum
    const F2 lny = mul2(l2[0], f2(kLn2));                       // ln y_0
    const F2 c1 = add2(lny, f2(fc.log_rate));                   // d/da of [a log(rate) + lse] when OC == 1
    if (OC == 1) out.common = fma2(b2[0], f2(kLn2), out.common);
    F2 gsum = f2(0.0f), S[kK];
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        const F2 a = mul2(img[m], f2(fc.rate));
        F2 ia, dl, ym;   // 1/a,  d lse / d a,  softmax mean of y
        if (OC == 1) {
            ia = rcp_2(a);
            ym = y[0];
        } else {
            F2 v[OC], mx;
#pragma unroll
            for (int j = 0; j < OC; ++j) v[j] = fma2(a, l2[j], b2[j]);
            mx = v[0];
#pragma unroll
            for (int j = 1; j < OC; ++j) { mx.x = fmaxf(mx.x, v[j].x); mx.y = fmaxf(mx.y, v[j].y); }
            F2 se, sl, sy;
#pragma unroll
            for (int j = 0; j < OC; ++j) {
                const F2 e = ex2_2(sub2(v[j], mx));
                if (j == 0) { se = e; sl = mul2(e, l2[0]); sy = mul2(e, y[0]); }
                else { se = add2(se, e); sl = fma2(e, l2[j], sl); sy = fma2(e, y[j], sy); }
            }
            const F2 inv = rcp_2(mul2(a, se));   // one reciprocal for 1/a and 1/se
            ia = mul2(inv, se);
            const F2 ise = mul2(inv, a);
            dl = mul2(mul2(sl, ise), f2(kLn2));
            ym = mul2(sy, ise);
            out.logp[m] = fma2(add2(mx, lg2_2(se)), f2(kLn2), out.logp[m]);
        }
        const F2 la = mul2(lg2_2(a), f2(kLn2));
        const F2 ia2 = mul2(ia, ia);
        // Stirling: r = ia (1/12 + ia2 (-1/360 + ia2/1260)),  q = ia (1/2 + ia (1/12 + ia2 (-1/120 + ia2/252)))
        const F2 r = mul2(ia, fma2(ia2, fma2(ia2, f2(0.000793650794f), f2(-0.00277777778f)), f2(0.0833333333f)));
        const F2 q = mul2(ia, fma2(ia, fma2(ia2, fma2(ia2, f2(0.00396825397f), f2(-0.00833333333f)), f2(0.0833333333f)), f2(0.5f)));
        // -lgamma(a) = (1/2 - a) ln a + a - ln(2 pi)/2 - r
        const F2 nl = sub2(add2(fma2(sub2(f2(0.5f), a), la, a), f2(-kHalfLn2Pi)), r);
        F2 dLda;   // d/da [a log(rate) - lgamma(a) + lse]
        if (OC == 1) {
            out.logp[m] = add2(out.logp[m], fma2(a, c1, nl));
            dLda = sub2(c1, sub2(la, q));
        } else {
            out.logp[m] = add2(out.logp[m], fma2(a, f2(fc.log_rate), nl));
            dLda = add2(dl, sub2(f2(fc.log_rate), sub2(la, q)));
        }
        const F2 gi = mul2(dLda, f2(W[m]));   // times rate: finish_pair
        out.g_rate = fma2(sub2(fma2(img[m], dLda, img[m]), ym), f2(W[m]), out.g_rate);
        gsum = add2(gsum, gi);
        if (m == 1) S[0] = gi;
        if (m == 2) S[1] = gi;
        if (m == 3) { S[0] = add2(S[0], gi); S[1] = add2(S[1], gi); }
    }
    out.g_b = add2(out.g_b, gsum);
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F2 t = mul2(S[k], mu[k]);
        out.g_h[k] = add2(out.g_h[k], t);
        out.g_x[k] = fma2(t, f2(dx[k]), out.g_x[k]);
        out.g_y[k] = fma2(t, dy[k], out.g_y[k]);
        out.g_w[k] = fma2(t, fma2(dy[k], dy[k], f2(dx2[k])), out.g_w[k]);
    }
}

// ---- many offset bins (histograms of real movies: tens to hundreds of distinct bins, glimpse_reader.py:403-424) -------------
// One pass over the bins per pixel pair, everything relative to a REFERENCE bin (the smallest offset delta_ref: its
// y_ref = D - delta_ref is the largest, so it is valid for a pixel whenever any bin is):
//     v_j(m) - v_ref(m) = (a_m - 1) (lg2 y_j - lg2 y_ref) + c_j,    c_j = (w2_j - w2_ref) + rate2 (delta_j - delta_ref)
// -- c_j and y_j - y_ref = delta_ref - delta_j do not depend on the pixel (per-bin constants in shared memory), lg2 y_j is
// shared by the four configurations, and the exponent needs no running maximum: bins spread over a few dozen counts
// keep |v_j - v_ref| within a few tens of bits.  Per pixel and bin: 1 lg2 + 4 ex2, 9 packed FP32 operations per pair
// (the two-pass form this replaces: 2 x 4 lg2 + 4 ex2 per pixel and bin, 6.75 ms per 100 000 patches at O = 64).
// A bin at or below the pixel (y_j <= 0: excluded, ksmogn.py:225-236) has its y_j clamped to 2^-40: with a - 1 >= 3 its
// term underflows to an exact zero.
struct alignas(16) BinConst { float delta, c, dyc, pad; };
struct ManyBinSums {
    F2 se[kM], sl[kM], sy[kM];   // sum_j e_j, sum_j e_j (lg2 y_j - lg2 y_ref), sum_j e_j (y_j - y_ref),  e_j = 2^(v_j - v_ref)
    TQ_HD void zero() {
#pragma unroll
        for (int m = 0; m < kM; ++m) se[m] = sl[m] = sy[m] = f2(0.0f);
    }
};
constexpr float kTinyY = 9.094947017729282e-13f;   // 2^-40

TQ_HD void many_bins_accumulate(F2 D, F2 l_ref, const F2 (&am1)[kM], const BinConst& bc, ManyBinSums& acc) {
    F2 y = sub2(D, f2(bc.delta));
    y.x = fmaxf(y.x, kTinyY);
    y.y = fmaxf(y.y, kTinyY);
    const F2 dl = sub2(lg2_2(y), l_ref);
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        const F2 e = ex2_2(fma2(am1[m], dl, f2(bc.c)));
        acc.se[m] = add2(acc.se[m], e);
        acc.sl[m] = fma2(e, dl, acc.sl[m]);
        acc.sy[m] = fma2(e, f2(bc.dyc), acc.sy[m]);
    }
}

// The rest of a pixel pair once its bins are summed: same quantities as pixel_pair_accumulate_fast.
//   y_ref, l_ref   D - delta_ref and its lg2;  w2_ref  log2-weight of the reference bin
TQ_HD void pixel_pair_finish_many(const F2 (&img)[kM], const F2 (&mu)[kK], F2 y_ref, F2 l_ref, float w2_ref,
                                  const ManyBinSums& sums, const float (&dx)[kK], const float (&dx2)[kK], const F2 (&dy)[kK],
                                  const FastConst& fc, const float (&W)[kM], PairOut& out) {
    F2 gsum = f2(0.0f), S[kK];
    const F2 vb = fma2(f2(-fc.rate2), y_ref, f2(w2_ref));            // v_ref(m) = (a_m - 1) l_ref + vb
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        const F2 a = mul2(img[m], f2(fc.rate));
        const F2 inv = rcp_2(mul2(a, sums.se[m]));                   // one reciprocal for 1/a and 1/se
        const F2 ia = mul2(inv, sums.se[m]), ise = mul2(inv, a);
        const F2 ml2 = fma2(sums.sl[m], ise, l_ref);                 // softmax mean of lg2 y
        const F2 ym = fma2(sums.sy[m], ise, y_ref);                  // softmax mean of y
        const F2 lse2 = add2(fma2(sub2(a, f2(1.0f)), l_ref, vb), lg2_2(sums.se[m]));
        const F2 la = mul2(lg2_2(a), f2(kLn2));
        const F2 ia2 = mul2(ia, ia);
        const F2 r = mul2(ia, fma2(ia2, fma2(ia2, f2(0.000793650794f), f2(-0.00277777778f)), f2(0.0833333333f)));
        const F2 q = mul2(ia, fma2(ia, fma2(ia2, fma2(ia2, f2(0.00396825397f), f2(-0.00833333333f)), f2(0.0833333333f)), f2(0.5f)));
        const F2 nl = sub2(add2(fma2(sub2(f2(0.5f), a), la, a), f2(-kHalfLn2Pi)), r);   // -lgamma(a)
        out.logp[m] = add2(out.logp[m], fma2(lse2, f2(kLn2), fma2(a, f2(fc.log_rate), nl)));
        const F2 dLda = add2(mul2(ml2, f2(kLn2)), sub2(f2(fc.log_rate), sub2(la, q)));
        const F2 gi = mul2(dLda, f2(W[m]));   // times rate: finish_pair
        out.g_rate = fma2(sub2(fma2(img[m], dLda, img[m]), ym), f2(W[m]), out.g_rate);
        gsum = add2(gsum, gi);
        if (m == 1) S[0] = gi;
        if (m == 2) S[1] = gi;
        if (m == 3) { S[0] = add2(S[0], gi); S[1] = add2(S[1], gi); }
    }
    out.g_b = add2(out.g_b, gsum);
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F2 t = mul2(S[k], mu[k]);
        out.g_h[k] = add2(out.g_h[k], t);
        out.g_x[k] = fma2(t, f2(dx[k]), out.g_x[k]);
        out.g_y[k] = fma2(t, dy[k], out.g_y[k]);
        out.g_w[k] = fma2(t, fma2(dy[k], dy[k], f2(dx2[k])), out.g_w[k]);
    }
}

// One pixel pair against all bins (the kernel's loop; also what tests/hostcheck runs): bins[0..O) hold the per-bin
// constants, ref = index of the reference bin.
TQ_HD void pixel_pair_many_bins(F2 D, const float (&gxh)[kK], const F2 (&gyk)[kK], const float (&dx)[kK],
                                const float (&dx2)[kK], const F2 (&dy)[kK], const PatchSpots<float>& s,
                                const FastConst& fc, int O, const BinConst* __restrict__ bins, float delta_ref,
                                float w2_ref, const float (&W)[kM], PairOut& out) {
    F2 mu[kK], img[kM], am1[kM];
#pragma unroll
    for (int k = 0; k < kK; ++k) mu[k] = mul2(gyk[k], f2(gxh[k]));
    img[0] = f2(s.b);
    img[1] = add2(mu[0], f2(s.b));
    img[2] = add2(mu[1], f2(s.b));
    img[3] = add2(img[1], mu[1]);
#pragma unroll
    for (int m = 0; m < kM; ++m) am1[m] = fma2(img[m], f2(fc.rate), f2(-1.0f));
    const F2 y_ref = sub2(D, f2(delta_ref)), l_ref = lg2_2(y_ref);
    ManyBinSums sums;
    sums.zero();
#pragma unroll 2
    for (int j = 0; j < O; ++j) many_bins_accumulate(D, l_ref, am1, bins[j], sums);
    pixel_pair_finish_many(img, mu, y_ref, l_ref, w2_ref, sums, dx, dx2, dy, fc, W, out);
}

// per-bin constants for the form above; returns the reference bin (smallest offset)
TQ_HD int many_bins_reference(int O, const float* off_s) {
    int ref = 0;
    for (int j = 1; j < O; ++j) if (off_s[j] < off_s[ref]) ref = j;
    return ref;
}
TQ_HD BinConst many_bins_const(int j, int ref, const float* off_s, const float* off_w2, float rate2) {
    BinConst b;
    b.delta = off_s[j];
    b.c = (off_w2[j] - off_w2[ref]) + rate2 * (off_s[j] - off_s[ref]);
    b.dyc = off_s[ref] - off_s[j];
    b.pad = 0.0f;
    return b;
}

// ---- single offset bin (the simulator's data after merging its identical bins) ---------------------------------
// With one bin the log-sum-exp is its only term; per pixel-configuration what is left is lgamma / digamma of
// a = image/gain and a handful of FMAs.  The spot-free configuration has the SAME a = b/gain at every pixel, so
// its sums over pixels are closed forms of sum(y) and sum(ln(y/b)) (finish_single_bin); terms common to all
// configurations and the constants of Stirling's formula are added once per patch as well.
struct SingleBinConst {
    float a0;      // b / gain
    float la0p1;   // ln(a0) + 1
    float q0p1;    // ln(a0) - psi(a0) + 1
    float q0;
    float K0;      // ln(a0)/2 + a0 - ln(2 pi)/2 - r(a0)  (= a0 ln a0 - lgamma(a0))
    float neg_lnb; // -ln(b)
};
TQ_HD SingleBinConst single_bin_const(float b, const FastConst& fc) {
    SingleBinConst c;
    c.a0 = b * fc.rate;
    // MUFU logarithms (absolute error 2e-7): the same approximation every pixel's ln(y/b) goes through, and the
    // library logf is ~20 instructions of a per-patch prologue that is a quarter of the kernel
    const float ia = rcp_newton(c.a0), la0 = kLn2 * f_lg2(c.a0);
    float r, q;
    stirling(ia, r, q);
    c.la0p1 = la0 + 1.0f;
    c.q0 = q;
    c.q0p1 = q + 1.0f;
    c.K0 = 0.5f * la0 + c.a0 - kHalfLn2Pi - r;
    c.neg_lnb = -kLn2 * f_lg2(b);
    return c;
}

struct PairOut1 {
    F2 logp[kM], g_b, g_rate, g_h[kK], g_w[kK], g_x[kK], g_y[kK];   // logp[0] unused
    F2 sum_dc, sum_yb;
    TQ_HD void zero() {
#pragma unroll
        for (int m = 0; m < kM; ++m) logp[m] = f2(0.0f);
        g_b = g_rate = sum_dc = sum_yb = f2(0.0f);
#pragma unroll
        for (int k = 0; k < kK; ++k) g_h[k] = g_w[k] = g_x[k] = g_y[k] = f2(0.0f);
    }
};

// Same contract as pixel_pair_accumulate_fast<1>: two pixels of one column, every pixel above the offset, a >= 4.
// ROWFIXED: the caller sweeps ONE row pair (dy fixed) over the columns and applies single_bin_row_fixup at the end of the
// row: the y- and the dy^2-part of the width moment are then products of dy with the height moment and are not accumulated
// per pair (4 packed operations fewer per pair).
template <bool ROWFIXED = false>
TQ_HD void pixel_pair_single_bin(F2 D, const float (&gxh)[kK], const F2 (&gyk)[kK], const float (&dx)[kK],
                                 const float (&dx2)[kK], const F2 (&dy)[kK], const PatchSpots<float>& s,
                                 const FastConst& fc, const SingleBinConst& sc, float off, const float (&W)[kM],
                                 PairOut1& out) {
    F2 mu[kK], img[kM];
#pragma unroll
    for (int k = 0; k < kK; ++k) mu[k] = mul2(gyk[k], f2(gxh[k]));
    img[1] = add2(mu[0], f2(s.b));
    img[2] = add2(mu[1], f2(s.b));
    img[3] = add2(img[1], mu[1]);
    const F2 y = sub2(D, f2(off)), ny = sub2(f2(off), D);
    const F2 dc = fma2(lg2_2(y), f2(kLn2), f2(sc.neg_lnb));   // ln(y / b)
    out.sum_dc = add2(out.sum_dc, dc);
    out.sum_yb = add2(out.sum_yb, sub2(y, f2(s.b)));          // deviations from b: small numbers, accurate sums
    const F2 c1p1 = add2(dc, f2(sc.la0p1));                    // d/da [a log(rate) + lse] + 1
    const F2 nry = mul2(ny, f2(fc.rate));                      // -y / gain
    // spot-free configuration: only its (nonlinear in y) contribution to d/d(1/gain) is per pixel
    out.g_rate = fma2(fma2(f2(s.b), add2(dc, f2(sc.q0p1)), ny), f2(W[0]), out.g_rate);
    F2 gsum = f2(0.0f), S[kK];
#pragma unroll
    for (int m = 1; m < kM; ++m) {
        const F2 a = mul2(img[m], f2(fc.rate));
        const F2 ia = rcp_2(a);
        const F2 la = mul2(lg2_2(a), f2(kLn2));
        const F2 ia2 = mul2(ia, ia);
        // Stirling remainders for a >= 4 (the caller's contract), one term shorter than stirling(): the dropped terms are
        // a^-5 / 1260 <= 7.7e-7 in ln Gamma and a^-6 / 252 <= 9.7e-7 in psi at a = 4, 2e-10 / 4e-11 at a typical a = 21
        const F2 r = mul2(ia, fma2(ia2, f2(-0.00277777778f), f2(0.0833333333f)));
        const F2 q = mul2(ia, fma2(ia, fma2(ia2, f2(-0.00833333333f), f2(0.0833333333f)), f2(0.5f)));
        const F2 d = sub2(c1p1, la);
        // a (c1 + 1 - ln a) - y/gain + ln(a)/2 - r  =  a log(rate) - lgamma(a) + a ln y - y/gain + ln(2 pi)/2:
        // the first two terms nearly cancel (a ~ y/gain), so they are combined BEFORE entering the running sum
        out.logp[m] = add2(out.logp[m], sub2(fma2(la, f2(0.5f), fma2(a, d, nry)), r));
        const F2 e = add2(d, q);                               // dL/da + 1
        out.g_rate = fma2(fma2(img[m], e, ny), f2(W[m]), out.g_rate);
        const F2 gi = fma2(e, f2(W[m]), f2(-W[m]));            // W dL/da (times rate: finish)
        gsum = add2(gsum, gi);
        if (m == 1) S[0] = gi;
        if (m == 2) S[1] = gi;
        if (m == 3) { S[0] = add2(S[0], gi); S[1] = add2(S[1], gi); }
    }
    out.g_b = add2(out.g_b, gsum);
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F2 t = mul2(S[k], mu[k]);
        out.g_h[k] = add2(out.g_h[k], t);
        out.g_x[k] = fma2(t, f2(dx[k]), out.g_x[k]);
        if (ROWFIXED) {
            out.g_w[k] = fma2(t, f2(dx2[k]), out.g_w[k]);
        } else {
            out.g_y[k] = fma2(t, dy[k], out.g_y[k]);
            out.g_w[k] = fma2(t, fma2(dy[k], dy[k], f2(dx2[k])), out.g_w[k]);
        }
    }
}

// end of a row pair swept with ROWFIXED (accumulators zero at its start): sum t dy = dy sum t, sum t dy^2 = dy^2 sum t
TQ_HD void single_bin_row_fixup(PairOut1& p, const F2 (&dy)[kK]) {
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        p.g_y[k] = mul2(p.g_h[k], dy[k]);
        p.g_w[k] = fma2(mul2(dy[k], dy[k]), p.g_h[k], p.g_w[k]);
    }
}

// lane-level fold for the single-bin form; npix = pixels this lane swept; log_w = log weight of the bin
TQ_HD void finish_single_bin(const PairOut1& p, const SingleBinConst& sc, const FastConst& fc, float b, float log_w,
                             float W0, int npix, PatchOut<float, kM>& out) {
    const float n = float(npix);
    const float sdc = p.sum_dc.x + p.sum_dc.y, syb = p.sum_yb.x + p.sum_yb.y;
    // terms shared by all configurations, -y/gain excepted: sum over pixels of  log w - ln y
    const float common = n * (log_w + sc.neg_lnb) - sdc;
    // spot-free configuration: a0 ln(y/b) + [a0 ln a0 - lgamma(a0)] - y/gain, with y/gain = a0 + (y - b)/gain
    out.logp[0] = sc.a0 * sdc + n * (sc.K0 - sc.a0) - fc.rate * syb + common;
#pragma unroll
    for (int m = 1; m < kM; ++m) out.logp[m] = (p.logp[m].x + p.logp[m].y) - n * kHalfLn2Pi + common;
    out.g_b = ((p.g_b.x + p.g_b.y) + W0 * (sdc + n * sc.q0)) * fc.rate;
    out.g_rate = p.g_rate.x + p.g_rate.y;
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        out.g_h[k] = (p.g_h[k].x + p.g_h[k].y) * fc.rate;
        out.g_w[k] = (p.g_w[k].x + p.g_w[k].y) * fc.rate;
        out.g_x[k] = (p.g_x[k].x + p.g_x[k].y) * fc.rate;
        out.g_y[k] = (p.g_y[k].x + p.g_y[k].y) * fc.rate;
    }
}

// ---- single offset bin, ROW form (sweep_patch_rows_single_bin): one row pair per lane, everything in units of a ----------
// The lane's row factors arrive pre-multiplied by 1/gain (gyr = gy * rate), so mu' = gyr * gxh IS the spots' share of
// a = image/gain and no pixel-configuration needs its own multiply; the spot moments are accumulated in the same units
// (t' = S mu' = rate t: exactly the factor finish_single_bin applied afterwards).  d/d(1/gain) needs NO per-pixel work:
// with W_m e_m = gi_m + W_m (gi = W (dL/da), the quantity the spot moments are made of)
//   rate * sum_m W_m (a_m e_m - y/gain)  =  sumW dev + W_0 a_0 (ln(y/b) + q_0) + a_0 gsum + t'_0 + t'_1
//                                           + (W_1 + W_3) mu'_0 + (W_2 + W_3) mu'_1,        dev = -(y - b)/gain
// and every term's sum over the pixels is either accumulated anyway (sum dev, sum ln(y/b), sum gsum = g_b, sum t'_k =
// the height moments) or separable (sum mu'_k = row factor x sum over columns of gxh_k).  60 packed operations per pixel
// pair against 75 (B200, C3: kernel 3.57 -> see DESIGN 4.1).
struct RowOut {
    F2 logp[kM];                       // [0] unused
    F2 g_b, g_h[kK], g_x[kK], g_w[kK];
    F2 sum_dc, sum_dev;
    float colsum[kK];                  // sum over the swept columns of gxh_k
    TQ_HD void zero() {
#pragma unroll
        for (int m = 0; m < kM; ++m) logp[m] = f2(0.0f);
        g_b = sum_dc = sum_dev = f2(0.0f);
#pragma unroll
        for (int k = 0; k < kK; ++k) { g_h[k] = g_x[k] = g_w[k] = f2(0.0f); colsum[k] = 0.0f; }
    }
};
struct RowConst {
    float off, off_rate, offb_rate, neg_rate;   // offset, offset / gain, (offset + b) / gain, -1 / gain
};
TQ_HD RowConst row_const(float off, float b, const FastConst& fc) {
    return RowConst{off, off * fc.rate, (off + b) * fc.rate, -fc.rate};
}

// two pixels (rows r, r + 7) of one column; contract of pixel_pair_single_bin (every pixel above the offset, a >= 4)
TQ_HD void row_pair_single_bin(F2 D, const float (&gxh)[kK], const F2 (&gyr)[kK], const float (&dx)[kK], const float (&dx2)[kK],
                               const SingleBinConst& sc, const RowConst& rc, const float (&W)[kM], RowOut& out) {
    F2 mu[kK], a[kM];
#pragma unroll
    for (int k = 0; k < kK; ++k) mu[k] = mul2(gyr[k], f2(gxh[k]));
    a[1] = add2(mu[0], f2(sc.a0));
    a[2] = add2(mu[1], f2(sc.a0));
    a[3] = add2(a[1], mu[1]);
    const F2 y = sub2(D, f2(rc.off));
    const F2 nry = fma2(D, f2(rc.neg_rate), f2(rc.off_rate));          // -y / gain
    const F2 dev = fma2(D, f2(rc.neg_rate), f2(rc.offb_rate));         // -(y - b) / gain: small numbers, accurate sums
    const F2 dc = fma2(lg2_2(y), f2(kLn2), f2(sc.neg_lnb));            // ln(y / b)
    out.sum_dev = add2(out.sum_dev, dev);
    out.sum_dc = add2(out.sum_dc, dc);
    const F2 c1p1 = add2(dc, f2(sc.la0p1));                            // d/da [a log(rate) + lse] + 1
    F2 gi[kM];
#pragma unroll
    for (int m = 1; m < kM; ++m) {
        const F2 ia = rcp_2(a[m]), l2 = lg2_2(a[m]);
        const F2 ia2 = mul2(ia, ia);
        const F2 d = fma2(l2, f2(-kLn2), c1p1);                        // c1 + 1 - ln a
        // a d - y/gain nearly cancels (a ~ y/gain): formed BEFORE it enters the running sum; ln(a)/2 and the Stirling
        // remainder -r = ia (ia^2 / 360 - 1/12) go in directly (short tails for a >= 4: see pixel_pair_single_bin)
        const F2 t1 = fma2(a[m], d, nry);
        F2 lp = fma2(l2, f2(0.5f * kLn2), out.logp[m]);
        lp = fma2(ia, fma2(ia2, f2(0.00277777778f), f2(-0.0833333333f)), lp);
        out.logp[m] = add2(lp, t1);
        // e = d + q,  q = ia (1/2 + ia (1/12 - ia^2 / 120)):  dL/da + 1
        const F2 e = fma2(ia, fma2(ia, fma2(ia2, f2(-0.00833333333f), f2(0.0833333333f)), f2(0.5f)), d);
        gi[m] = fma2(e, f2(W[m]), f2(-W[m]));                          // W dL/da
    }
    const F2 S[kK] = {add2(gi[1], gi[3]), add2(gi[2], gi[3])};
    out.g_b = add2(out.g_b, add2(S[0], gi[2]));
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F2 t = mul2(S[k], mu[k]);
        out.g_h[k] = add2(out.g_h[k], t);
        out.g_x[k] = fma2(t, f2(dx[k]), out.g_x[k]);
        out.g_w[k] = fma2(t, f2(dx2[k]), out.g_w[k]);
        out.colsum[k] += gxh[k];
    }
}

// end of the lane's row pair: y- and dy^2-moments from the height moment (dy is fixed along a row), then the lane-level fold
// into the scalar record (before the cross-lane reduction).  npix = pixels this lane swept (0 for the idle eighth lane)
TQ_HD void finish_row_single_bin(const RowOut& p, const F2 (&gyr)[kK], const F2 (&dy)[kK], const SingleBinConst& sc,
                                 const FastConst& fc, float log_w, const float (&W)[kM], int npix, PatchOut<float, kM>& out) {
    const float n = float(npix);
    const float sdc = p.sum_dc.x + p.sum_dc.y, sdev = p.sum_dev.x + p.sum_dev.y;
    // terms shared by all configurations, -y/gain excepted: sum over pixels of  log w - ln y
    const float common = n * (log_w + sc.neg_lnb) - sdc;
    // spot-free configuration: a0 ln(y/b) + [a0 ln a0 - lgamma(a0)] - y/gain, with -y/gain = dev - a0
    out.logp[0] = sc.a0 * sdc + n * (sc.K0 - sc.a0) + sdev + common;
#pragma unroll
    for (int m = 1; m < kM; ++m) out.logp[m] = (p.logp[m].x + p.logp[m].y) - n * kHalfLn2Pi + common;
    const float gsum = p.g_b.x + p.g_b.y, w0 = W[0] * (sdc + n * sc.q0);
    out.g_b = (gsum + w0) * fc.rate;
    float grate = (W[0] + W[1] + W[2] + W[3]) * sdev + sc.a0 * (w0 + gsum);
    const float Wk[kK] = {W[1] + W[3], W[2] + W[3]};
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F2 gy = fma2(mul2(dy[k], dy[k]), p.g_h[k], p.g_w[k]);    // sum t' (dx^2 + dy^2)
        const F2 gyy = mul2(p.g_h[k], dy[k]);                          // sum t' dy
        const float h = p.g_h[k].x + p.g_h[k].y;
        out.g_h[k] = h;
        out.g_x[k] = p.g_x[k].x + p.g_x[k].y;
        out.g_y[k] = gyy.x + gyy.y;
        out.g_w[k] = gy.x + gy.y;
        grate += h + Wk[k] * ((gyr[k].x + gyr[k].y) * p.colsum[k]);
    }
    out.g_rate = grate * fc.gain;
}

// lane-level fold of the pair sums into the scalar record (before the cross-lane reduction)
TQ_HD void finish_pair(const PairOut& p, float rate, PatchOut<float, kM>& out) {
#pragma unroll
    for (int m = 0; m < kM; ++m) out.logp[m] = (p.logp[m].x + p.logp[m].y) + (p.common.x + p.common.y);
    out.g_b = (p.g_b.x + p.g_b.y) * rate;
    out.g_rate = p.g_rate.x + p.g_rate.y;
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        out.g_h[k] = (p.g_h[k].x + p.g_h[k].y) * rate;
        out.g_w[k] = (p.g_w[k].x + p.g_w[k].y) * rate;
        out.g_x[k] = (p.g_x[k].x + p.g_x[k].y) * rate;
        out.g_y[k] = (p.g_y[k].x + p.g_y[k].y) * rate;
    }
}

// Per-patch conversion of the accumulated moments (after the cross-lane reduction):
//   A0 = sum t, A1 = sum t dx, A2 = sum t dy, A3 = sum t (dx^2 + dy^2)
//   d/dh = A0 / h,  d/dx = A1 / w^2,  d/dy = A2 / w^2,  d/dw = A3 / w^3 - 2 A0 / w
TQ_HD void finish_spot_moments(const PatchSpots<float>& s, PatchOut<float, kM>& out) {
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const float A0 = out.g_h[k], A3 = out.g_w[k];
        const float iw = rcp_newton(s.w[k]), iw2 = iw * iw;
        out.g_h[k] = A0 * rcp_newton(s.h[k]);
        out.g_x[k] *= iw2;
        out.g_y[k] *= iw2;
        out.g_w[k] = iw * fmaf(iw2, A3, -2.0f * A0);   // (explicit: two kernels evaluate this and must round alike)
    }
}

}  // namespace tq
