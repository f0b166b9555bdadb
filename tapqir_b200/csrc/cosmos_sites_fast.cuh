// Production (fp32) form of site_eval (cosmos_local.cuh): the same quantities -- guide sample, log q,
// d log q / d sample and the linear maps A_p, B_p of the reparameterised gradient -- evaluated in
// single precision without the cancellations that make the textbook formulas need double.
//
// The reference reaches these through torch (Gamma/Beta rsample -> ATen _standard_gamma_grad /
// _dirichlet_grad, log_prob -> lgamma/digamma; models/cosmos.py:408-462, affine_beta.py:33-49).  All
// of their ill-conditioned pieces are functions of ONE small quantity per site:
//
//   Gamma(conc, rate) draw x (standard variate):      u  = x / conc - 1
//   Beta(c1, c0) draw x, mean m1 = c1 / (c1 + c0):     ua = (x - m1) / m1,   ub = -(x - m1) / m0
//
// u / ua are formed in DOUBLE (one exp and one fused multiply-add from the fp32 parameters); everything
// after that is fp32 on (l, f) = (log1p(u), u - log1p(u)) and on the "small parts"
// A = l/u - 1, F = 2 f/u^2 - 1, which an atanh series gives to full relative accuracy:
//
//   lgamma(a+b) - lgamma(a) - lgamma(b) + (a-1) log x + (b-1) log(1-x)
//        = -(c1 fa + c0 fb) - log x - log(1-x) + (log c1 + log c0 - log tot)/2 - ln(2 pi)/2 + r(tot) - r(c1) - r(c0)
//   psi(tot) - psi(c1) + log x  = la + q(c1) - q(tot)                  (r, q: Stirling remainders)
//   Rice expansion of d x / d alpha (ATen, alpha > 8):   stirling (1 + u) [ l/u - H(u)/alpha ]
//   Rice expansion of the Beta gradient (ATen, both > 6): stirling (x / c1) [ la/ua - (m1 m0 / tot) B / d^2 ]
//
// (H, B collect the terms that cancel to second order in ATen's form; tests/test_hostcheck_sites.py
// pins every output of this file against site_eval in double over all regimes.)  Outside the regimes
// handled here (tiny concentrations, samples on the clamps, ...) site_eval_fast returns false and the
// caller runs the double-precision site_eval.
#pragma once
#include "cosmos_local.cuh"

namespace tq {

constexpr float kSiteHalfLn2Pi = 0.91893853320467274178f;

// reciprocal: one MUFU op on the device (1 ulp), exact division on the host
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float site_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
inline float site_rcp(float x) { return 1.0f / x; }
#endif

// (1 + g)^(-3/2) - 1 with full relative accuracy for small g
TQ_HD float pow_m32_m1(float g) {
    if (fabsf(g) <= 0.1f)
        return g * (-1.5f + g * (1.875f + g * (-2.1875f + g * (2.4609375f + g * (-2.70703125f + g * (2.9326171875f
                 + g * (-3.14208984375f + g * 3.338470458984375f)))))));
    // |g| > 0.1: nothing cancels in (1 + g)^(-3/2) - 1 (its magnitude is >= 0.13); one rsqrt instead of log1p + expm1
#ifdef __CUDA_ARCH__
    const float r = rsqrtf(1.0f + g);
    return fmaf(r * r, r, -1.0f);
#else
    return expm1f(-1.5f * log1pf(g));
#endif
}

// l = log1p(u), f = u - l, A = l/u - 1, F = 2 f / u^2 - 1, for u > -1; r = 1 + u rounded from its own
// (double) evaluation, so that l keeps its accuracy when u is close to -1
struct Lp1 { float l, f, A, F; };

TQ_HD Lp1 lp1_parts(float u, float r) {
    Lp1 o;
    if (fabsf(u) <= 0.4f) {
        // log1p(u) = 2 atanh(s), s = u / (2 + u);  u - 2 s = s u
        const float inv = site_rcp(2.0f + u);
        const float s = u * inv, z = s * s;
        const float P = 0.333333333f + z * (0.2f + z * (0.142857143f + z * (0.111111111f + z * (0.0909090909f + z * 0.0769230769f))));
        const float R = z * P;
        o.l = fmaf(2.0f * s, R, 2.0f * s);
        o.f = s * (u - 2.0f * R);
        o.A = (2.0f * R - u) * inv;
        o.F = -(4.0f * (u * inv * inv * P) + u) * inv;
    } else {
        o.l = logf(r);
        o.f = u - o.l;
        const float iu = site_rcp(u);
        o.A = o.l * iu - 1.0f;
        o.F = 2.0f * o.f * (iu * iu) - 1.0f;
    }
    return o;
}

// Stirling remainders for z >= 6 from iz = 1/z (next terms: 5e-12 / 1e-11 at z = 6):
//   lgamma(z) = (z - 1/2) ln z - z + ln(2 pi)/2 + r(z),   psi(z) = ln z - q(z)
TQ_HD float stirling_r(float iz) {
    const float z2 = iz * iz;
    return iz * (0.0833333333f - z2 * (0.00277777778f - z2 * (0.000793650794f - z2 * (0.000595238095f - z2 * 0.000841750842f))));
}
TQ_HD float stirling_q(float iz) {
    const float z2 = iz * iz;
    return iz * (0.5f + iz * (0.0833333333f - z2 * (0.00833333333f - z2 * (0.00396825397f - z2 * (0.00416666667f - z2 * 0.00757575758f)))));
}

// lgamma(x) and digamma(x) in fp32 for x > 0: Stirling series for x >= 8, below that the recurrence shifted
// by 8 with the eight reciprocals folded into two divisions and the eight factors into one logarithm
TQ_HD void lgamma_digamma_f32(float x, float& lg, float& psi) {
    float xs = x, lprod = 0.0f, rsum = 0.0f;
    if (x < 8.0f) {
        const float p01 = x * (x + 1.0f), p23 = (x + 2.0f) * (x + 3.0f), p45 = (x + 4.0f) * (x + 5.0f), p67 = (x + 6.0f) * (x + 7.0f);
        const float tx = 2.0f * x;
        const float pa = p01 * p23, pb = p45 * p67;
        rsum = ((tx + 1.0f) * p23 + (tx + 5.0f) * p01) * site_rcp(pa) + ((tx + 9.0f) * p67 + (tx + 13.0f) * p45) * site_rcp(pb);
        lprod = logf(pa * pb);
        xs = x + 8.0f;
    }
    const float ix = site_rcp(xs), lxs = logf(xs);
    lg = (xs - 0.5f) * lxs - xs + kSiteHalfLn2Pi + stirling_r(ix) - lprod;
    psi = lxs - stirling_q(ix) - rsum;
}

// d x / d alpha of a standard Gamma(alpha) draw outside the Rice regime (ATen _standard_gamma_grad):
// Taylor series of the incomplete gamma function for x < 0.8 (the x^alpha factors of cdf and pdf divided
// out), else the bivariate rational fit in (log(x/alpha), log alpha).  lx = log x, l = log(x / alpha).
TQ_HD float gamma_grad_small(float alpha, float x, float lx, float l, float lalpha, float psi) {
    if (x < 0.8f) {
        float numer = 1.0f, denom = alpha;
        float r = site_rcp(denom);
        float s1 = r, s2 = r * r;
#pragma unroll
        for (int i = 1; i <= 5; ++i) {
            numer *= -x * (1.0f / float(i));
            denom += 1.0f;
            r = site_rcp(denom);
            s1 = fmaf(numer, r, s1);
            s2 = fmaf(numer, r * r, s2);
        }
        const float res = -x * expf(x) * ((lx - psi) * s1 - s2);
        return (res != res) ? 0.0f : res;
    }
    const float c[3][8] = {
        {0.16009398f, -0.094634809f, 0.025146376f, -0.0030648343f, 1.0f, 0.32668115f, 0.10406089f, 0.0014179084f},
        {0.53487893f, 0.1298071f, 0.065735949f, -0.0015649758f, 0.16639465f, 0.020070113f, -0.0035938915f, -0.00058392623f},
        {0.040121004f, -0.0065914022f, -0.0026286047f, -0.0013441777f, 0.017050642f, -0.0021309326f, 0.00085092367f, -1.5247877e-07f},
    };
    float cv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cv[i] = c[0][i] + l * (c[1][i] + l * c[2][i]);
    const float pnum = cv[0] + lalpha * (cv[1] + lalpha * (cv[2] + lalpha * cv[3]));
    const float qden = cv[4] + lalpha * (cv[5] + lalpha * (cv[6] + lalpha * cv[7]));
    return expf(pnum * site_rcp(qden));
}

// d x / d alpha of a standard Gamma(alpha) draw x = alpha (1 + u), alpha > 8, x >= 0.8: ATen's Rice
// expansion and its Taylor patch, returned as  sgg  and  sgg - (1 + u)  (the latter without cancellation)
TQ_HD void gamma_grad_rice(float alpha, float ia, float u, const Lp1& p, float& sgg, float& sgg_m1u) {
    const float st = ia * (0.0833333333f + ia * 0.00347222222f);   // stirling - 1
    if (fabsf(u) <= 0.1f) {
        // n1 n2 / den = (1 + st) (1 + P),  P = u/2 - u^2/6 + (360 + 188 u - 65 u^2) / (4320 alpha)
        const float c = (360.0f + u * (188.0f - 65.0f * u)) * (ia * (1.0f / 4320.0f));
        const float Pm = -u * (0.5f + u * 0.166666667f) + c;   // P - u
        const float P = Pm + u;
        sgg_m1u = st + Pm + st * P;
        sgg = 1.0f + u + sgg_m1u;
        return;
    }
    // H(u) = [ (l/u) g^(-3/2) - (1 + u/2) ] / u^2,  g = 2 f / u^2
    const float E = pow_m32_m1(p.F);                              // g^(-3/2) - 1
    const float B = p.A + E + p.A * E - 0.5f * u;
    const float H = B * site_rcp(u * u);
    const float inner = p.A - H * ia;                             // l/u - H/alpha - 1
    sgg_m1u = (1.0f + u) * (st + inner + st * inner);             // (1+u)(1+st)(1+inner) - (1+u)
    sgg = (1.0f + u) + sgg_m1u;
}

// reciprocal of a positive double from the fp32 MUFU seed and one Newton step (1e-14): the series below need their
// divisions in double, and a DP division costs several times this
TQ_HD double rcp_d(double d) {
#ifdef __CUDA_ARCH__
    const double r = (double)site_rcp((float)d);
    return fma(r, fma(-d, r, 1.0), r);
#else
    return 1.0 / d;
#endif
}

// ATen's Beta reparameterisation gradient d x / d alpha (tq_math.cuh::beta_grad: same regimes, series and coefficient
// tables) OUTSIDE the regime site_eval_fast reformulates (both concentrations > 6 and the draw in the bulk): small
// concentrations -- absent spots, whose guides relax towards the flat prior -- and draws in the tails.  The caller
// passes what it already has:
//   xd, yd       the draw and 1 - draw (double: draws pile up against 0 and 1 at small concentrations)
//   lx, ly       ln x, ln y (log1p of the small one where the other is close to 1)
//   dpsi         psi(total) - psi(alpha);   dc = ln x + dpsi  (= d log q / d alpha, already needed for the density)
//   la, lb       ln(alpha) - ln x  and  ln(x total / alpha)  when have_logs (formed without cancellation at large total)
//   rice         the reformulated Rice value when have_rice (alpha, beta > 6: the draw is then in a tail, where that
//                expansion has nothing left to cancel)
// The two power series alternate with terms up to e^(beta x) times their sum: their ten / eight terms are accumulated in
// double (arguments and result stay fp32; FP64 FMAs run at half the FP32 rate on sm_100, it is double TRANSCENDENTALS
// that are slow).  Returns false where no fp32 form applies (caller: double form).
TQ_HD_NOINLINE bool beta_grad_tierb(double xd, double yd, float lx, float ly, float alpha, float beta, float total, float dpsi,
                                    float dc, bool have_logs, float la, float lb, bool have_rice, float rice, float& g) {
    const float x = (float)xd, y = (float)yd;
    const float boundary = total * x * y;
    if (x <= 0.5f && boundary < 2.5f) {
        // power series in x around 0
        const double f = -(double)dc, a = alpha, b = beta;
        double numer = 1.0, id = rcp_d(a);
        double series = id * (f + id);
        const double inv_i[11] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10};
#pragma unroll
        for (int i = 1; i <= 10; ++i) {
            numer *= ((double)i - b) * xd * inv_i[i];
            id = rcp_d(a + (double)i);
            series = fma(numer * id, f + id, series);
        }
        const float res = x * expf(-beta * ly) * (float)series;       // x (1 - x)^(-beta) series
        g = (res != res) ? 0.0f : res;
        return true;
    }
    if (x >= 0.5f && boundary < 0.75f) {
        // the mirrored series in y = 1 - x
        const double f = dpsi, a = alpha, b = beta;
        double numer = 1.0, betas = 1.0, dbetas = 0.0, series = f * rcp_d(b);
        const double inv_j[9] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8};
#pragma unroll
        for (int i = 1; i <= 8; ++i) {
            numer *= -yd * inv_j[i];
            dbetas = dbetas * (a - (double)i) + betas;
            betas = betas * (a - (double)i);
            series = fma(numer * rcp_d(b + (double)i), dbetas + f * betas, series);
        }
        const float res = expf((1.0f - alpha) * lx) * (float)series;   // -(-(1 - y)^(1 - alpha) series)
        g = (res != res) ? 0.0f : res;
        return true;
    }
    if (alpha > 6.0f && beta > 6.0f) {
        g = rice;
        return have_rice;
    }
    // rational correction to an analytic approximation (coefficients: torch Distributions.h, as in tq_math.cuh)
    // (polynomials accumulated in double: d sample / d size is a difference of two of these gradients that cancels ten- to
    // thirty-fold for lopsided guides, and 72 FP64 FMAs are cheap on sm_100)
    const double c[2][3][3][4] = {
        {{{1.003668233, -0.01061107488, -0.0657888334, 0.01201642863},
          {0.6336835991, -0.3557432599, 0.05486251648, -0.001465281033},
          {-0.03276231906, 0.004474107445, 0.002429354597, -0.0001557569013}},
         {{0.221950385, -0.3187676331, 0.01799915743, 0.01074823814},
          {-0.2951249643, 0.06219954479, 0.01535556598, 0.001550077057},
          {0.02155310298, 0.004170831599, 0.001292462449, 6.976601077e-05}},
         {{-0.05980841433, 0.008441916499, 0.01085618172, 0.002319392565},
          {0.02911413504, 0.01400243777, -0.002721828457, 0.000751041181},
          {0.005900514878, -0.001936558688, -9.495446725e-06, 5.385558597e-05}}},
        {{{1.0, -0.02924021934, -0.04438342661, 0.007285809825},
          {0.6357567472, -0.3473456711, 0.05454656494, -0.002407477521},
          {-0.03301322327, 0.004845219414, 0.00231480583, -0.0002307248149}},
         {{0.5925320577, -0.1757678135, 0.01505928619, 0.000564515273},
          {0.1014815858, -0.06589186703, 0.01272886114, -0.0007316646956},
          {-0.007258481865, 0.001096195486, 0.0003934994223, -4.12701925e-05}},
         {{0.06469649321, -0.0236701437, 0.002902096474, -5.896963079e-05},
          {0.001925008108, -0.002869809258, 0.0008000589141, -6.063713228e-05},
          {-0.0003477407336, 6.959756487e-05, 1.097287507e-05, -1.650964693e-06}}},
    };
    const float uaf = have_logs ? la : logf(alpha) - lx;
    const double ua = uaf, b = have_logs ? lb : logf(total) - uaf, u = lx;
    const double pu[3] = {1.0, u, u * u};
    const double pa[3] = {1.0, ua, ua * ua};
    double p = 0.0, q = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double w = pu[i] * pa[j];
            p = fma(w, c[0][i][j][0] + b * (c[0][i][j][1] + b * (c[0][i][j][2] + b * c[0][i][j][3])), p);
            q = fma(w, c[1][i][j][0] + b * (c[1][i][j][1] + b * (c[1][i][j][2] + b * c[1][i][j][3])), q);
        }
    g = (float)(p * rcp_d(q) * (xd * (double)dpsi * rcp_d((double)beta)));
    return true;
}

// Density pieces of Beta(cs, cb) with total S > 16 and ONE small concentration cs <= 6 (a guide pushed against an edge of
// its interval): Stirling for lgamma(S) - lgamma(cb), the small side directly.
//   xs, lxs, us   the draw on the small side, its log, and (xs - ms) / ms
//   pb            log1p parts of the big side's (xb - mb) / mb;  lmb = ln mb
//   lqs           log q + ln(scale);  d_cs, d_cb = d log q / d c;  dpsi_* = psi(S) - psi(c_*);  b1 = ms d_cs + mb d_cb
struct BetaLopsided { float lqs, d_cs, d_cb, dpsi_s, dpsi_b, b1; };
TQ_HD BetaLopsided beta_density_lopsided(float cs, float cb, float S, float xs, float lxs, float us, const Lp1& pb, float ms,
                                         float mb, float lmb) {
    BetaLopsided o;
    float lgs, psis;
    lgamma_digamma_f32(cs, lgs, psis);
    const float it = site_rcp(S), ib = site_rcp(cb);
    const float qt = stirling_q(it), qb = stirling_q(ib);
    const float lSx = logf(S * xs);
    o.lqs = (cb - 1.0f) * pb.l - 0.5f * lmb + cs * lSx - lxs - cs + stirling_r(it) - stirling_r(ib) - lgs;
    o.d_cs = lSx - qt - psis;
    o.d_cb = pb.l + qb - qt;
    o.dpsi_s = logf(S) - qt - psis;
    o.dpsi_b = qb - qt - lmb;
    o.b1 = ms * (lSx - psis - us) - mb * pb.f + mb * qb - qt;       // ms us + mb ub = 0 used
    return o;
}

// Numerator polynomial of the Taylor patch of ATen's Beta gradient around x = mean
// (beta_grad_alpha_mid, tq_math.cuh):  grad = pn / (1 - x) * poly / pd.  Evaluated in double because the
// size gradient needs poly(x, a, b) - poly(1 - x, b, a), which cancels to first order.
TQ_HD double beta_patch_poly(double x, double alpha, double beta) {
    const double b2 = beta * beta;
    return 47.0 * x * b2 * b2 + alpha * (
               (43.0 + 20.0 * (16.0 + 27.0 * beta) * x) * b2 * beta + alpha * (
               3.0 * (59.0 + 180.0 * beta - 90.0 * x) * b2 + alpha * (
               (453.0 + 1620.0 * beta * (1.0 - x) - 455.0 * x) * beta + alpha * (
               8.0 * (1.0 - x) * (135.0 * beta - 11.0)))));
}

// log-density of v ~ Gamma(conc, rate) and its partials from x = rate v = conc (1 + u):
//   lp, d_v, and the two combinations the site maps need:  conc * d_conc  and  conc * d_conc + rate * d_rate
TQ_HD void gamma_density_fast(float conc, float lconc, float lrate, float rate, float x, float u, const Lp1& p,
                              float& lp, float& d_v, float& cdc, float& cdc_rdr, float& psi, bool known_big = false) {
    const float lx = lconc + p.l;
    d_v = rate * (-fmaf(conc, u, 1.0f)) * site_rcp(x);            // (conc - 1)/v - rate
    if (known_big || conc > 10.0f) {
        const float ic = site_rcp(conc);
        const float q = stirling_q(ic);
        lp = lrate - lx - conc * p.f + 0.5f * lconc - kSiteHalfLn2Pi - stirling_r(ic);
        cdc = conc * (p.l + q);                                     // conc (lx - psi(conc))
        cdc_rdr = conc * (q - p.f);                                 // ... + (conc - x)
        psi = lconc - q;
    } else {
        float lg;
        lgamma_digamma_f32(conc, lg, psi);
        lp = lrate - lx + conc * lx - x - lg;
        cdc = conc * (lx - psi);
        cdc_rdr = cdc - conc * u;
    }
}

// Same arguments and outputs as site_eval, with float records.  Returns SITE_DONE, or -- when the site is
// outside the regimes handled in fp32 -- SITE_FALLBACK (variate holds the base draw: call site_eval in
// replay mode) or SITE_FALLBACK_DRAW (nothing drawn yet: call site_eval as is).
//
// MODE splits the work for site_fast_kernel, whose warps would otherwise drag every lane through every regime any lane is in
// (a trained model: 13.6 of 32 lanes active per instruction): 1 = draw, classify, and evaluate only the BULK regime
// (Gamma: concentration > 10 and draw >= 0.8; Beta: both concentrations > 6 and the draw away from the tails), returning
// SITE_DEFER with `variate` and a regime class in `cls` otherwise; 2 = replay a deferred site (everything but the bulk
// forms and the sampler compiled out); 0 = everything in one call.
enum { SITE_DONE = 0, SITE_FALLBACK = 1, SITE_FALLBACK_DRAW = 2, SITE_DEFER = 3 };
// Regime classes of deferred sites.  A class fixes every data-dependent branch of the replay (MODE 2), so that sites sorted
// by class run side by side in one warp through the same code:
//   Gamma (8 classes):  bit 0: x < 0.8,  bit 1: concentration > 8,  bit 2: concentration > 10
//   Beta (64 classes):  density regime (S <= 16 / c1 <= 6 / c0 <= 6 / both > 6 with a tail draw) x the branch each of the
//                       two beta_grad_tierb calls takes (series in x, series in 1 - x, Rice, rational correction)
constexpr int kSiteClasses = 64;
TQ_HD int beta_tierb_branch(float x, float boundary, float alpha, float beta) {
    if (x <= 0.5f && boundary < 2.5f) return 0;
    if (x >= 0.5f && boundary < 0.75f) return 1;
    return (alpha > 6.0f && beta > 6.0f) ? 2 : 3;
}
// `trials`: the site's Marsaglia-Tsang trial pairs (RNG mode); the caller may have preloaded its first block
template <int MODE>
TQ_HD int site_eval_fast_t(int s, float u0, float u1, float ubm, float ubs, const ModelConst& mc, bool use_rng,
                           Philox* rng, GammaTrials& trials, double& variate, float& sample, float* rec, float* extra, int& cls) {
    if (!(mc.eps < 1e-12)) return SITE_FALLBACK_DRAW;   // fp32 reference conventions move the clamps into range
    if (site_is_gamma(s)) {
        // Gamma(loc * beta, beta): background cosmos.py:408-415, height cosmos.py:428-435
        const float lconc = u0 + u1;
        if (!(lconc > -4.0f && lconc < 13.0f) || fabsf(u1) > 40.0f) return SITE_FALLBACK_DRAW;
        const double ed = exp(-((double)u0 + (double)u1));         // 1 / conc
        const float ic = (float)ed, conc = site_rcp(ic);
        if (use_rng) variate = fmax((double)sample_std_gamma_f32(*rng, trials, conc), mc.tiny);
        const double xd = variate;
        if (!(xd > 1e-18) || !(xd < 1e18)) return SITE_FALLBACK;
        const double rd = xd * ed;                                 // x / conc
        const float u = (float)(rd - 1.0);
        const float x = (float)xd;
        if (x < 0.8f && conc > 30.0f) return SITE_FALLBACK;        // Taylor regime far in the tail: powers underflow in fp32
        const bool bulk = conc > 10.0f && x >= 0.8f;
        if (MODE == 1 && !bulk) {
            cls = (x < 0.8f ? 1 : 0) | (conc > 8.0f ? 2 : 0) | (conc > 10.0f ? 4 : 0);
            return SITE_DEFER;
        }
        const float ibeta = expf(-u1), beta = site_rcp(ibeta), loc = conc * ibeta;
        const float v = x * ibeta;
        const Lp1 p = lp1_parts(u, (float)rd);
        float lp, d_v, cdc, cdc_rdr, psi;
        gamma_density_fast(conc, lconc, u1, beta, x, u, p, lp, d_v, cdc, cdc_rdr, psi, MODE == 1);
        float sgg, sgg_m1u;
        if (MODE == 1 || (x >= 0.8f && conc > 8.0f)) {
            gamma_grad_rice(conc, ic, u, p, sgg, sgg_m1u);
        } else {
            sgg = gamma_grad_small(conc, x, lconc + p.l, p.l, lconc, psi);
            sgg_m1u = sgg - (1.0f + u);
        }
        if (s == S_B) {
            // model prior Gamma(pc, pr), pc = (bm/bs)^2, pr = bm/bs^2 at the same v: pr v / pc = v / bm   cosmos.py:233-239
            const float lpc = 2.0f * (ubm - ubs), lpr = ubm - 2.0f * ubs;
            if (!(lpc > -4.0f && lpc < 13.0f) || fabsf(lpr) > 40.0f) return SITE_FALLBACK;
            const double rpd = xd * exp(-((double)u1 + (double)ubm));
            const float up = (float)(rpd - 1.0);
            if (!(up > -1.0f)) return SITE_FALLBACK;
            const float pc = expf(lpc), pr = expf(lpr);
            const Lp1 pp = lp1_parts(up, (float)rpd);
            float plp, pd_v, pcdc, pcdc_rdr, ppsi;
            gamma_density_fast(pc, lpc, lpr, pr, pr * v, up, pp, plp, pd_v, pcdc, pcdc_rdr, ppsi);
            extra[EX_LP] = plp;
            extra[EX_DP] = pd_v;
            // d pc / d u_bm = 2 pc, d pr / d u_bm = pr;  d pc / d u_bs = -2 pc, d pr / d u_bs = -2 pr
            extra[EX_GBM] = pcdc + pcdc_rdr;
            extra[EX_GBS] = -2.0f * pcdc_rdr;
        }
        sample = v;
        rec[SO_LQ] = lp;
        rec[SO_DQ] = d_v;
        rec[SO_A0] = sgg * loc;
        rec[SO_A1] = sgg_m1u * loc;                                 // sgg loc - v,  v = loc (1 + u)
        rec[SO_B0] = cdc;
        rec[SO_B1] = cdc_rdr;
        return SITE_DONE;
    }
    // AffineBeta(mean, size, lo, hi): width cosmos.py:436-444, x :445-453, y :454-462
    double lod, hid;
    if (s < S_X) { lod = mc.width_min; hid = mc.width_max; } else { hid = 0.5 * (double)(mc.P + 1); lod = -hid; }
    const float scale = (float)(hid - lod);
    if (fabsf(u0) > 15.0f || !(u1 < 11.5f)) return SITE_FALLBACK_DRAW;
    const double ed = exp(-(double)u0);                           // m1 = 1 / (1 + e), m0 = e m1
    const float e = (float)ed;
    const float m1 = site_rcp(1.0f + e), m0 = e * m1;
    const float sz = expf(u1), S = 2.0f + sz;                     // size - 2 = d size / d u, size
    const float c1 = S * m1, c0 = S * m0;
    if (use_rng) {
        const double g1 = sample_std_gamma_f32(*rng, trials, c1), g2 = sample_std_gamma_f32(*rng, trials, c0);
        variate = beta01_from_gammas(g1, g2, mc);
    }
    const double x01d = variate;   // (v - low) / scale of the reference, to rounding
    const float ua = (float)fma(x01d, ed, x01d - 1.0);            // (x - m1) / m1
    const float x = (float)x01d, y = (float)(1.0 - x01d);
    const float ie = site_rcp(e), ub = -ua * ie;                  // -(x - m1) / m0
    const bool bulk = c1 > 6.0f && c0 > 6.0f && S * x * y >= 2.5f && ua > -1.0f && ub > -1.0f;
    if (MODE == 1 && !bulk) {
        const float boundary = S * x * y;
        cls = (S <= 16.0f ? 0 : (c1 <= 6.0f ? 1 : (c0 <= 6.0f ? 2 : 3))) * 16 + beta_tierb_branch(x, boundary, c1, c0) * 4
              + beta_tierb_branch(y, boundary, c0, c1);
        return SITE_DEFER;
    }
    if (MODE != 1 && (MODE == 2 || !bulk)) {
        // ---- small concentrations (absent spots: the guide relaxes towards the flat prior), a guide pushed against an
        // edge of its interval, or a draw in a tail: per-gradient regimes of beta_grad_tierb, and for the density
        //   S <= 16                      the textbook formulas (nothing large enough to cancel)
        //   S > 16, both c > 6           the Stirling / log1p forms of the bulk regime below
        //   S > 16, one c <= 6           beta_density_lopsided (the other is then >= 10)
        if (!(x > 1e-30f) || !(y > 1e-30f) || !(ua > -1.0f) || !(ub > -1.0f)) return SITE_FALLBACK;
        const bool big = c1 > 6.0f && c0 > 6.0f;
        const float lx = x > 0.5f ? log1pf(-y) : logf(x), ly = y > 0.5f ? log1pf(-x) : logf(y);
        const float d = ua * m1, mm = m1 * m0;
        float lqs, d_c1, d_c0, dpsi1, dpsi0, b1, la1 = 0.0f, lb1 = 0.0f, la0 = 0.0f, lb0 = 0.0f, rice1 = 0.0f, rice0 = 0.0f;
        bool have_logs = false;
        Lp1 pa, pb;
        if (big || S > 16.0f) {
            pa = lp1_parts(ua, x * (1.0f + e));
            pb = lp1_parts(ub, y * (1.0f + ie));
        }
        if (big) {
            // Rice expansion of the bulk regime (below): here the draw is in a tail, d is not small
            const float it = site_rcp(S), i1 = it * (1.0f + e), i0 = it * (1.0f + ie);
            const float Gs = m0 * pa.F + m1 * pb.F;
            const float E = pow_m32_m1(Gs);
            const float h = 0.5f * d * site_rcp(mm);
            const float Ba = pa.A + E + pa.A * E - h * (m0 - 2.0f * m1);
            const float Bb = pb.A + E + pb.A * E + h * (m1 - 2.0f * m0);
            const float w = mm * site_rcp(S * d * d);
            const float stir = (1.0f + i1 * (0.0833333333f + i1 * 0.00347222222f)) * (1.0f + i0 * (0.0833333333f + i0 * 0.00347222222f))
                             * site_rcp(1.0f + it * (0.0833333333f + it * 0.00347222222f));
            rice1 = stir * x * site_rcp(c1) * (1.0f + (pa.A - w * Ba));
            rice0 = stir * y * site_rcp(c0) * (1.0f + (pb.A - w * Bb));
        }
        if (S <= 16.0f) {
            float lgt, pt, lg1, p1, lg0, p0;
            lgamma_digamma_f32(S, lgt, pt);
            lgamma_digamma_f32(c1, lg1, p1);
            lgamma_digamma_f32(c0, lg0, p0);
            dpsi1 = pt - p1;
            dpsi0 = pt - p0;
            d_c1 = lx + dpsi1;
            d_c0 = ly + dpsi0;
            lqs = (c1 - 1.0f) * lx + (c0 - 1.0f) * ly + lgt - lg1 - lg0;
            b1 = m1 * d_c1 + m0 * d_c0;
        } else if (big) {
            const float it = site_rcp(S), i1 = it * (1.0f + e), i0 = it * (1.0f + ie);
            const float q1 = stirling_q(i1), q0 = stirling_q(i0), qt = stirling_q(it);
            const float lm1 = -log1pf(e), lm0 = -log1pf(ie), lt = logf(S);
            const float kl = m1 * pa.f + m0 * pb.f;
            lqs = -S * kl - (lm1 + pa.l) - (lm0 + pb.l) + 0.5f * (lt + lm1 + lm0) - kSiteHalfLn2Pi
                  + stirling_r(it) - stirling_r(i1) - stirling_r(i0);
            d_c1 = pa.l + q1 - qt;
            d_c0 = pb.l + q0 - qt;
            dpsi1 = q1 - qt - lm1;
            dpsi0 = q0 - qt - lm0;
            b1 = m1 * q1 + m0 * q0 - qt - kl;
            have_logs = true;
            lb1 = pa.l; la1 = lt - lb1; lb0 = pb.l; la0 = lt - lb0;
        } else if (c1 <= 6.0f) {
            const float lm0 = -log1pf(ie), lt = logf(S);
            const BetaLopsided o = beta_density_lopsided(c1, c0, S, x, lx, ua, pb, m1, m0, lm0);
            lqs = o.lqs; d_c1 = o.d_cs; d_c0 = o.d_cb; dpsi1 = o.dpsi_s; dpsi0 = o.dpsi_b; b1 = o.b1;
            have_logs = true;
            lb1 = pa.l; la1 = lt - lb1; lb0 = pb.l; la0 = lt - lb0;
        } else {
            const float lm1 = -log1pf(e), lt = logf(S);
            const BetaLopsided o = beta_density_lopsided(c0, c1, S, y, ly, ub, pa, m0, m1, lm1);
            lqs = o.lqs; d_c0 = o.d_cs; d_c1 = o.d_cb; dpsi0 = o.dpsi_s; dpsi1 = o.dpsi_b; b1 = o.b1;
            have_logs = true;
            lb1 = pa.l; la1 = lt - lb1; lb0 = pb.l; la0 = lt - lb0;
        }
        const double y01d = 1.0 - x01d;
        float bg1, bg0;
        if (!beta_grad_tierb(x01d, y01d, lx, ly, c1, c0, S, dpsi1, d_c1, have_logs, la1, lb1, big, rice1, bg1) ||
            !beta_grad_tierb(y01d, x01d, ly, lx, c0, c1, S, dpsi0, d_c0, have_logs, la0, lb0, big, rice0, bg0)) return SITE_FALLBACK;
        const float dv_dc1 = scale * y * bg1, dv_dc0 = -scale * x * bg0;
        sample = (float)(lod + (hid - lod) * variate);
        rec[SO_LQ] = lqs - logf(scale);
        rec[SO_DQ] = (-S * d + (x - y)) * site_rcp(x * y * scale);
        rec[SO_A0] = (dv_dc1 - dv_dc0) * (S * mm);
        rec[SO_B0] = (d_c1 - d_c0) * (S * mm);
        rec[SO_A1] = sz * (dv_dc1 * m1 + dv_dc0 * m0);
        rec[SO_B1] = sz * b1;
        return SITE_DONE;
    }
    const float d = ua * m1;                                      // x - m1
    const Lp1 pa = lp1_parts(ua, x * (1.0f + e)), pb = lp1_parts(ub, y * (1.0f + ie));
    const float it = site_rcp(S), i1 = it * (1.0f + e), i0 = it * (1.0f + ie);
    const float q1 = stirling_q(i1), q0 = stirling_q(i0), qt = stirling_q(it);
    const float lm1 = -log1pf(e), lm0 = lm1 - u0;                 // log m1, log m0
    const float lt = logf(S);
    const float kl = m1 * pa.f + m0 * pb.f;                       // S kl = c1 log(m1/x) + c0 log(m0/y) >= 0
    const float lp = -S * kl - (lm1 + pa.l) - (lm0 + pb.l) + 0.5f * (lt + lm1 + lm0) - kSiteHalfLn2Pi
                     + stirling_r(it) - stirling_r(i1) - stirling_r(i0) - logf(scale);
    const float d_v = (-S * d + (x - y)) * site_rcp(x * y * scale);
    const float d_c1 = pa.l + q1 - qt, d_c0 = pb.l + q0 - qt;
    const float mm = m1 * m0;
    float A0, A1;
    if (d * d * (S + 1.0f) <= 0.01f * mm) {
        // y bg1 = K P1 / c1,  x bg0 = K P0 / c0,  K = pn / (12960 c1^2 c0^2 (1 + 12 S))
        const double P1 = beta_patch_poly(x01d, (double)c1, (double)c0), P0 = beta_patch_poly(1.0 - x01d, (double)c0, (double)c1);
        const float pn = (1.0f + 12.0f * c1) * (1.0f + 12.0f * c0) * (it * it);
        const float K = pn * site_rcp(12960.0f * (c1 * c1) * (c0 * c0) * (1.0f + 12.0f * S));
        A0 = scale * K * ((float)P1 * m0 + (float)P0 * m1);
        A1 = scale * sz * it * K * (float)(P1 - P0);
    } else {
        const float Gs = m0 * pa.F + m1 * pb.F;                   // 2 m1 m0 kl / d^2 - 1
        const float E = pow_m32_m1(Gs);
        const float h = 0.5f * d * site_rcp(mm);
        const float Ba = pa.A + E + pa.A * E - h * (m0 - 2.0f * m1);
        const float Bb = pb.A + E + pb.A * E + h * (m1 - 2.0f * m0);
        const float w = mm * site_rcp(S * d * d);
        const float ta = pa.A - w * Ba, tb = pb.A - w * Bb;      // bg1 = stir (x/c1) (1 + ta), bg0 = stir (y/c0) (1 + tb)
        const float stir = (1.0f + i1 * (0.0833333333f + i1 * 0.00347222222f)) * (1.0f + i0 * (0.0833333333f + i0 * 0.00347222222f))
                         * site_rcp(1.0f + it * (0.0833333333f + it * 0.00347222222f));
        const float sxy = scale * stir * x * y;
        A0 = sxy * (m0 * (1.0f + ta) + m1 * (1.0f + tb));
        A1 = sxy * sz * it * (ta - tb);
    }
    sample = (float)(lod + (hid - lod) * variate);
    rec[SO_LQ] = lp;
    rec[SO_DQ] = d_v;
    rec[SO_A0] = A0;
    rec[SO_A1] = A1;
    rec[SO_B0] = S * mm * (d_c1 - d_c0);
    rec[SO_B1] = sz * (m1 * q1 + m0 * q0 - qt - kl);
    return SITE_DONE;
}

template <int MODE>
TQ_HD int site_eval_fast_t(int s, float u0, float u1, float ubm, float ubs, const ModelConst& mc, bool use_rng,
                           Philox* rng, double& variate, float& sample, float* rec, float* extra, int& cls) {
    GammaTrials trials;
    return site_eval_fast_t<MODE>(s, u0, u1, ubm, ubs, mc, use_rng, rng, trials, variate, sample, rec, extra, cls);
}
TQ_HD int site_eval_fast(int s, float u0, float u1, float ubm, float ubs, const ModelConst& mc, bool use_rng,
                         Philox* rng, double& variate, float& sample, float* rec, float* extra) {
    int cls = 0;
    return site_eval_fast_t<0>(s, u0, u1, ubm, ubs, mc, use_rng, rng, variate, sample, rec, extra, cls);
}

}  // namespace tq
