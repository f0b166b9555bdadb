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
    return expm1f(-1.5f * log1pf(g));
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

// ATen's Beta reparameterisation gradient (tq_math.cuh::beta_grad: same regimes, series and coefficient tables) in
// fp32 for SMALL concentrations (total <= 64: nothing in the plain formulas is large enough to cancel), with what the
// caller already has passed in instead of recomputed: y = 1 - x (formed in double: draws pile up against 0 and 1 at
// these concentrations), lx = ln x, ly = ln y, and the digammas psi(alpha), psi(total).  Returns false where fp32 is
// not enough (caller: double form).
TQ_HD bool beta_grad_f32(float x, float y, float lx, float ly, float alpha, float total, float psi_a, float psi_t, float& g) {
    const float beta = total - alpha;
    const float boundary = total * x * y;
    if (x <= 0.5f && boundary < 2.5f) {
        // power series in x around 0; terms alternate with size ~ (beta x)^i / i!: fine in fp32 while beta x < 2
        // (measured: 4e-6 at 2, 1e-5 at 2.5, 3e-5 at 3.5)
        if (!(beta * x < 2.0f)) return false;
        const float factor = psi_a - psi_t - lx;
        float numer = 1.0f, id = site_rcp(alpha);
        float series = id * (factor + id);
#pragma unroll
        for (int i = 1; i <= 10; ++i) {
            numer *= (float(i) - beta) * x * (1.0f / float(i));
            id = site_rcp(alpha + float(i));
            series = fmaf(numer * id, factor + id, series);
        }
        const float res = x * expf(-beta * ly) * series;           // x (1 - x)^(-beta) series
        g = (res != res) ? 0.0f : res;
        return true;
    }
    if (x >= 0.5f && boundary < 0.75f) {
        // the mirrored series in y = 1 - x
        if (!(alpha * y < 2.0f)) return false;
        const float factor = psi_t - psi_a;
        float numer = 1.0f, betas = 1.0f, dbetas = 0.0f, series = factor * site_rcp(beta);
#pragma unroll
        for (int i = 1; i <= 8; ++i) {
            numer *= -y * (1.0f / float(i));
            dbetas = dbetas * (alpha - float(i)) + betas;
            betas = betas * (alpha - float(i));
            series = fmaf(numer * site_rcp(beta + float(i)), dbetas + factor * betas, series);
        }
        const float res = expf((1.0f - alpha) * lx) * series;      // -(-(1 - y)^(1 - alpha) series)
        g = (res != res) ? 0.0f : res;
        return true;
    }
    if (alpha > 6.0f && beta > 6.0f) return false;   // Rice expansion outside the reformulated regime: double
    // rational correction to an analytic approximation (coefficients: torch Distributions.h, as in tq_math.cuh)
    const float c[2][3][3][4] = {
        {{{1.003668233f, -0.01061107488f, -0.0657888334f, 0.01201642863f},
          {0.6336835991f, -0.3557432599f, 0.05486251648f, -0.001465281033f},
          {-0.03276231906f, 0.004474107445f, 0.002429354597f, -0.0001557569013f}},
         {{0.221950385f, -0.3187676331f, 0.01799915743f, 0.01074823814f},
          {-0.2951249643f, 0.06219954479f, 0.01535556598f, 0.001550077057f},
          {0.02155310298f, 0.004170831599f, 0.001292462449f, 6.976601077e-05f}},
         {{-0.05980841433f, 0.008441916499f, 0.01085618172f, 0.002319392565f},
          {0.02911413504f, 0.01400243777f, -0.002721828457f, 0.000751041181f},
          {0.005900514878f, -0.001936558688f, -9.495446725e-06f, 5.385558597e-05f}}},
        {{{1.0f, -0.02924021934f, -0.04438342661f, 0.007285809825f},
          {0.6357567472f, -0.3473456711f, 0.05454656494f, -0.002407477521f},
          {-0.03301322327f, 0.004845219414f, 0.00231480583f, -0.0002307248149f}},
         {{0.5925320577f, -0.1757678135f, 0.01505928619f, 0.000564515273f},
          {0.1014815858f, -0.06589186703f, 0.01272886114f, -0.0007316646956f},
          {-0.007258481865f, 0.001096195486f, 0.0003934994223f, -4.12701925e-05f}},
         {{0.06469649321f, -0.0236701437f, 0.002902096474f, -5.896963079e-05f},
          {0.001925008108f, -0.002869809258f, 0.0008000589141f, -6.063713228e-05f},
          {-0.0003477407336f, 6.959756487e-05f, 1.097287507e-05f, -1.650964693e-06f}}},
    };
    const float ua = logf(alpha) - lx;
    const float b = logf(total) - ua;
    const float pu[3] = {1.0f, lx, lx * lx};
    const float pa[3] = {1.0f, ua, ua * ua};
    float p = 0.0f, q = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float w = pu[i] * pa[j];
            p = fmaf(w, c[0][i][j][0] + b * (c[0][i][j][1] + b * (c[0][i][j][2] + b * c[0][i][j][3])), p);
            q = fmaf(w, c[1][i][j][0] + b * (c[1][i][j][1] + b * (c[1][i][j][2] + b * c[1][i][j][3])), q);
        }
    g = p * site_rcp(q) * (x * (psi_t - psi_a) * site_rcp(beta));
    return true;
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
                              float& lp, float& d_v, float& cdc, float& cdc_rdr, float& psi) {
    const float lx = lconc + p.l;
    d_v = rate * (-fmaf(conc, u, 1.0f)) * site_rcp(x);            // (conc - 1)/v - rate
    if (conc > 10.0f) {
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
enum { SITE_DONE = 0, SITE_FALLBACK = 1, SITE_FALLBACK_DRAW = 2 };
TQ_HD int site_eval_fast(int s, float u0, float u1, float ubm, float ubs, const ModelConst& mc, bool use_rng,
                         Philox* rng, double& variate, float& sample, float* rec, float* extra) {
    if (!(mc.eps < 1e-12)) return SITE_FALLBACK_DRAW;   // fp32 reference conventions move the clamps into range
    if (site_is_gamma(s)) {
        // Gamma(loc * beta, beta): background cosmos.py:408-415, height cosmos.py:428-435
        const float lconc = u0 + u1;
        if (!(lconc > -4.0f && lconc < 13.0f) || fabsf(u1) > 40.0f) return SITE_FALLBACK_DRAW;
        const double ed = exp(-((double)u0 + (double)u1));         // 1 / conc
        const float ic = (float)ed, conc = site_rcp(ic);
        if (use_rng) variate = fmax((double)sample_std_gamma_f32(*rng, conc), mc.tiny);
        const double xd = variate;
        if (!(xd > 1e-18) || !(xd < 1e18)) return SITE_FALLBACK;
        const double rd = xd * ed;                                 // x / conc
        const float u = (float)(rd - 1.0);
        const float x = (float)xd;
        if (x < 0.8f && conc > 30.0f) return SITE_FALLBACK;        // Taylor regime far in the tail: powers underflow in fp32
        const float ibeta = expf(-u1), beta = site_rcp(ibeta), loc = conc * ibeta;
        const float v = x * ibeta;
        const Lp1 p = lp1_parts(u, (float)rd);
        float lp, d_v, cdc, cdc_rdr, psi;
        gamma_density_fast(conc, lconc, u1, beta, x, u, p, lp, d_v, cdc, cdc_rdr, psi);
        float sgg, sgg_m1u;
        if (x >= 0.8f && conc > 8.0f) {
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
        const double g1 = sample_std_gamma_f32(*rng, c1), g2 = sample_std_gamma_f32(*rng, c0);
        variate = beta01_from_gammas(g1, g2, mc);
    }
    const double x01d = variate;   // (v - low) / scale of the reference, to rounding
    const float ua = (float)fma(x01d, ed, x01d - 1.0);            // (x - m1) / m1
    const float x = (float)x01d, y = (float)(1.0 - x01d);
    const float ie = site_rcp(e), ub = -ua * ie;                  // -(x - m1) / m0
    if (!(c1 > 6.0f && c0 > 6.0f && S * x * y >= 2.5f && ua > -1.0f && ub > -1.0f)) {
        // ---- small concentrations (absent spots: the guide relaxes towards the flat prior), or a draw in the far tail:
        // the textbook formulas in fp32 -- nothing large enough to cancel while total <= 64
        if (!(S <= 64.0f) || !(x > 1e-30f) || !(y > 1e-30f)) return SITE_FALLBACK;
        float lgt, pt, lg1, p1, lg0, p0;
        lgamma_digamma_f32(S, lgt, pt);
        lgamma_digamma_f32(c1, lg1, p1);
        lgamma_digamma_f32(c0, lg0, p0);
        const float lx = logf(x), ly = logf(y);
        const float d_c1 = lx + pt - p1, d_c0 = ly + pt - p0;
        float bg1, bg0;
        if (!beta_grad_f32(x, y, lx, ly, c1, S, p1, pt, bg1) || !beta_grad_f32(y, x, ly, lx, c0, S, p0, pt, bg0)) return SITE_FALLBACK;
        const float dv_dc1 = scale * y * bg1, dv_dc0 = -scale * x * bg0;
        const float km = S * m1 * m0, k1 = m1 * sz, k0 = m0 * sz;
        sample = (float)(lod + (hid - lod) * variate);
        rec[SO_LQ] = (c1 - 1.0f) * lx + (c0 - 1.0f) * ly + lgt - lg1 - lg0 - logf(scale);
        rec[SO_DQ] = ((c1 - 1.0f) * site_rcp(x) - (c0 - 1.0f) * site_rcp(y)) * site_rcp(scale);
        rec[SO_A0] = (dv_dc1 - dv_dc0) * km;
        rec[SO_B0] = (d_c1 - d_c0) * km;
        rec[SO_A1] = dv_dc1 * k1 + dv_dc0 * k0;
        rec[SO_B1] = d_c1 * k1 + d_c0 * k0;
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

}  // namespace tq
