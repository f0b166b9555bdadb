// Per-unit ("local") part of the cosmos ELBO: everything models/cosmos.py:216-327 (model) and
// :393-462 (guide) do for one (AOI, frame, channel) patch except the pixel likelihood itself,
// written out with its analytic reverse mode.  SURVEY.md App. A.3 gives the ELBO; this file is its
// per-unit summand and the chain rule down to the unconstrained variational parameters.
//
// Two stages around the likelihood kernel:
//
//   site_eval  one call per (guide site, unit); 9 sites per unit at K = 2: background, and height,
//              width, x, y of each spot.  unconstrained params -> constrained -> guide sample (replayed
//              variate or in-kernel Philox) -> log q, d log q / d sample, and the LINEAR map that turns
//              the total derivative G = dELBO/dsample into the gradients of the site's two
//              variational parameters:   grad_p = G * A_p - w * B_p   (w = q(m_k = 1), 1 for background).
//              Always evaluated in double: ATen's implicit reparameterisation gradients and the
//              Beta/Gamma log-normalisers cancel catastrophically in fp32 ("precision").
//   unit_post  one call per unit, cheap (F = float in production): priors, the (z, theta)
//              log-sum-exp T(m), q(m)-weighted ELBO summand, d/d m_probs, G for every sample, parameter
//              gradients through the site maps, and the per-channel accumulators for the globals.
//
// Host+device: tests/hostcheck runs the same code on the CPU against the oracle.
#pragma once
#include "tq_math.cuh"

namespace tq {

constexpr int kMaxC = 4;  // channels the global tables are sized for

// Prior hyper-parameters (models/cosmos.py:55-64) + numeric conventions of the reference dtype.
struct ModelConst {
    double bg_mean_std, bg_std_std, lamda_rate, height_std, width_min, width_max, proximity_rate, gain_std;
    double eps;    // torch.finfo(reference dtype).eps: clamps of Categorical/Bernoulli/sigmoid/AffineBeta
    double tiny;   // torch.finfo(reference dtype).tiny
    double logit_lim;  // log((1 - eps) / eps): logit of the Bernoulli probs clamp
    int P;
};

// Per-channel tables derived from the sampled globals (built by globals_pre, cosmos_globals.cuh).
template <typename A> struct ChannelTables {
    A logpz[2][kZ];             // [is_ontarget][z]           cosmos.py:242-246, util.py:133-151
    A logptheta[2][kTheta];     // [min(z,1)][theta]          cosmos.py:247-255, util.py:154-173
    A logpm[kTheta][kK][2];     // [theta][k][m_k]            cosmos.py:262-267, util.py:94-130
    A logptrans[2][kZ][kZ];     // hmm only: [is_ontarget][z'][z]  hmm.py:165-169 (logpz is then the initial distribution)
};
template <typename A> struct GlobalTables {
    A gain, rate, log_rate;
    A size1;   // AffineBeta sample size of the target-specific spot: ((P+1)/(2 proximity))^2 - 1
    // log AffineBeta(x; 0, size1) + log AffineBeta(y; 0, size1)
    //    = (size1/2 - 1) [log1p(-tx^2) + log1p(-ty^2)] + cxy1,   tx = 2 x / (P+1)
    A cxy1;    // 2 [lgamma(size1) - 2 lgamma(size1/2)] - 4 (size1/2 - 1) ln 2 - 2 ln(P+1)
    A dcxy1;   // d cxy1-part / d size1 at fixed x, y: 2 [psi(size1) - psi(size1/2)] - 2 ln 2
    A lxy0;    // uniform (size 2) prior on x and y: -2 ln(P+1)
    ChannelTables<A> ch[kMaxC];
    template <typename B> TQ_HD void convert_from(const GlobalTables<B>& o) {
        gain = (A)o.gain; rate = (A)o.rate; log_rate = (A)o.log_rate; size1 = (A)o.size1;
        cxy1 = (A)o.cxy1; dcxy1 = (A)o.dcxy1; lxy0 = (A)o.lxy0;
        for (int q = 0; q < kMaxC; ++q) {
            for (int a = 0; a < 2; ++a) for (int z = 0; z < kZ; ++z) ch[q].logpz[a][z] = (A)o.ch[q].logpz[a][z];
            for (int a = 0; a < 2; ++a) for (int t = 0; t < kTheta; ++t) ch[q].logptheta[a][t] = (A)o.ch[q].logptheta[a][t];
            for (int t = 0; t < kTheta; ++t) for (int k = 0; k < kK; ++k) for (int m = 0; m < 2; ++m)
                ch[q].logpm[t][k][m] = (A)o.ch[q].logpm[t][k][m];
            for (int a = 0; a < 2; ++a) for (int z = 0; z < kZ; ++z) for (int y = 0; y < kZ; ++y)
                ch[q].logptrans[a][z][y] = (A)o.ch[q].logptrans[a][z][y];
        }
    }
};

// indices into the per-channel accumulator vector reduced over units
enum {
    ACC_ELBO_FRAME = 0,   // sum mu_n * (frame-level ELBO summand), unscaled
    ACC_ELBO_AOI = 1,     // sum mu_n * (AOI-level prior terms), once per (AOI, channel)
    ACC_LOGPZ = 2,        // [kZ]            d/d logpz[ontarget=1][z]
    ACC_LOGPM = 4,        // [kTheta][kK][2] d/d logpm
    ACC_SIZE1 = 16,       // d/d size1
    ACC_RATE = 17,        // d/d (1/gain) from the likelihood
    NACC = 18
};

template <typename A> struct Transformed { A v, d; };  // constrained value and d v / d unconstrained

template <typename A> TQ_HD A sigmoid_clamped(A u, const ModelConst& mc, bool& active) {
    // SigmoidTransform._call: clamp(sigmoid(u), finfo.tiny, 1 - finfo.eps)
    A s = A(1) / (A(1) + Real<A>::exp(-u));
    active = true;
    if (s < A(mc.tiny)) { s = A(mc.tiny); active = false; }
    if (s > A(1) - A(mc.eps)) { s = A(1) - A(mc.eps); active = false; }
    return s;
}
template <typename A> TQ_HD Transformed<A> t_positive(A u) { const A e = Real<A>::exp(u); return {e, e}; }
template <typename A> TQ_HD Transformed<A> t_greater_than(A u, A lb) { const A e = Real<A>::exp(u); return {lb + e, e}; }
template <typename A> TQ_HD Transformed<A> t_interval(A u, A lo, A hi, const ModelConst& mc) {
    bool act;
    const A s = sigmoid_clamped(u, mc, act);
    return {lo + (hi - lo) * s, act ? (hi - lo) * s * (A(1) - s) : A(0)};
}

// The AOI-local variational parameters of one unit.  Order = order of the flat gradient record and
// of the tensors in the flat parameter buffer (tapqir_b200/models/layout.py).
enum {
    LP_BM = 0, LP_BS, LP_B_LOC, LP_B_BETA,
    LP_M_PROBS,                 // + k
    LP_H_LOC = LP_M_PROBS + kK, // + k
    LP_H_BETA = LP_H_LOC + kK,
    LP_W_MEAN = LP_H_BETA + kK,
    LP_W_SIZE = LP_W_MEAN + kK,
    LP_X_MEAN = LP_W_SIZE + kK,
    LP_Y_MEAN = LP_X_MEAN + kK,
    LP_SIZE = LP_Y_MEAN + kK,
    NLOCAL = LP_SIZE + kK       // 20 at K = 2
};

// Guide sites of one unit = samples of one unit: background, then height, width, x, y per spot.
enum { S_B = 0, S_H = 1, S_W = S_H + kK, S_X = S_W + kK, S_Y = S_X + kK, NSAMP = S_Y + kK };  // 9 at K=2

// Per-site record written by site_eval (SoA rows of the (NREC, U) buffer)
enum { SO_LQ = 0, SO_DQ, SO_A0, SO_B0, SO_A1, SO_B1, NSO };
// extra rows of the background site: its model prior Gamma((bm/bs)^2, bm/bs^2)
enum { EX_LP = 0, EX_DP, EX_GBM, EX_GBS, NEX };
constexpr int NREC = NSAMP * NSO + NEX;  // 58 at K = 2

// which two local parameters parameterise site s, and which family it is
TQ_HD int site_param0(int s) {
    if (s == S_B) return LP_B_LOC;
    if (s < S_W) return LP_H_LOC + (s - S_H);
    if (s < S_X) return LP_W_MEAN + (s - S_W);
    if (s < S_Y) return LP_X_MEAN + (s - S_X);
    return LP_Y_MEAN + (s - S_Y);
}
TQ_HD int site_param1(int s) {
    if (s == S_B) return LP_B_BETA;
    if (s < S_W) return LP_H_BETA + (s - S_H);
    if (s < S_X) return LP_W_SIZE + (s - S_W);
    if (s < S_Y) return LP_SIZE + (s - S_X);
    return LP_SIZE + (s - S_Y);
}
TQ_HD bool site_is_gamma(int s) { return s < S_W; }

// Beta on [low, low+scale] in mean / sample-size form (affine_beta.py:33-49)
template <typename A> struct AffBeta {
    A low, scale, c1, c0;
    TQ_HD AffBeta(A mean, A size, A lo, A hi) : low(lo), scale(hi - lo), c1(size * (mean - lo) / (hi - lo)), c0(size * (hi - mean) / (hi - lo)) {}
    TQ_HD A clamp(A v, const ModelConst& mc) const {
        const A e = A(mc.eps) * scale;  // pyro AffineBeta.rsample clamp [third party]
        return Real<A>::min(Real<A>::max(v, low + e), low + scale - e);
    }
};

template <typename A> TQ_HD A beta01_from_gammas(A g1, A g2, const ModelConst& mc) {
    // _sample_dirichlet: normalise, clamp to [tiny, 1 - eps]
    A v = g1 / (g1 + g2);
    v = Real<A>::max(v, A(mc.tiny));
    return Real<A>::min(v, A(1) - A(mc.eps));
}

// ---- densities and their partials ----------------------------------------------------------------
TQ_HD void special(double x, double& lg, double& psi) { lgamma_digamma(x, lg, psi); }
TQ_HD void special(float x, float& lg, float& psi) { lg = lgammaf(x); psi = digamma<float>(x); }
template <typename A> struct GammaSite {
    // value v ~ Gamma(conc, rate): log-density and partials
    A lp, d_v, d_conc, d_rate;
    TQ_HD GammaSite(A v, A conc, A rate) {
        using R = Real<A>;
        const A lv = R::log(v), lr = R::log(rate);
        A lg, psi;
        special(conc, lg, psi);
        lp = conc * lr + (conc - A(1)) * lv - rate * v - lg;
        d_v = (conc - A(1)) / v - rate;
        d_conc = lr + lv - psi;
        d_rate = conc / rate - v;
    }
};
template <typename A> struct BetaSite {
    // value v = low + scale * x01, x01 ~ Beta(c1, c0): log-density of v and partials
    A lp, d_v, d_c1, d_c0, x01;
    TQ_HD BetaSite(A v, const AffBeta<A>& d) {
        using R = Real<A>;
        x01 = (v - d.low) / d.scale;
        const A l1 = R::log(x01), l0 = R::log(A(1) - x01);
        const A tot = d.c1 + d.c0;
        A lgt, pt, lg1, p1, lg0, p0;
        special(tot, lgt, pt);
        special(d.c1, lg1, p1);
        special(d.c0, lg0, p0);
        lp = (d.c1 - A(1)) * l1 + (d.c0 - A(1)) * l0 + lgt - lg1 - lg0 - R::log(d.scale);
        d_v = ((d.c1 - A(1)) / x01 - (d.c0 - A(1)) / (A(1) - x01)) / d.scale;
        d_c1 = l1 + pt - p1;
        d_c0 = l0 + pt - p0;
    }
};

// ---- site_eval -----------------------------------------------------------------------------------------
//   s         site index (S_B, S_H + k, S_W + k, S_X + k, S_Y + k)
//   u0, u1    unconstrained values of the site's two parameters (site_param0/1)
//   ubm, ubs  unconstrained background_mean_loc / background_std_loc (read only for s == S_B)
//   variate   base draw: standard gamma (Gamma sites) or (0,1) Beta draw; filled here when use_rng
//   rec[NSO]  site record;  extra[NEX] background-prior record (s == S_B only)
// Returns the guide sample.
//
// Both families have a "usual regime" path that reuses logarithms and reciprocals (the exp
// transforms make log(conc), log(rate) free; lgamma and digamma share one log and one reciprocal;
// the two Beta reparameterisation gradients share theirs) and fall back to the general functions
// outside it.  Same formulas either way; tests/test_hostcheck_step.py pins both against the oracle.

// log-density of v ~ Gamma(conc, rate) and partials, logs supplied: lv = log v, lr = log rate,
// lconc = log conc
TQ_HD void gamma_density(double v, double conc, double rate, double lv, double lr, double lconc,
                         double& lp, double& d_v, double& d_conc, double& d_rate) {
    double lg, psi;
    if (conc > 10.0) lgamma_digamma_known(conc, lconc, 1.0 / conc, lg, psi);
    else lgamma_digamma(conc, lg, psi);
    lp = conc * lr + (conc - 1.0) * lv - rate * v - lg;
    d_v = (conc - 1.0) / v - rate;
    d_conc = lr + lv - psi;
    d_rate = conc / rate - v;
}

TQ_HD double site_eval(int s, double u0, double u1, double ubm, double ubs, const ModelConst& mc, bool use_rng,
                       Philox* rng, double& variate, double* rec, double* extra) {
    using A = double;
    const A tiny = mc.tiny;
    if (site_is_gamma(s)) {
        // Gamma(loc * beta, beta): background cosmos.py:408-415, height cosmos.py:428-435
        const A loc = exp(u0), beta = exp(u1);
        const A conc = loc * beta;
        if (use_rng) variate = fmax((A)sample_std_gamma_f32(*rng, (float)conc), tiny);
        const A v = fmax(variate / beta, tiny);
        // exp transforms: log(rate) = u1 and log(conc) = u0 + u1 exactly
        const A lvar = log(v * beta);
        const A lv = lvar - u1, lconc = u0 + u1;
        A lp, d_v, d_conc, d_rate;
        gamma_density(v, conc, beta, lv, u1, lconc, lp, d_v, d_conc, d_rate);
        const A x = v * beta;
        const A sgg = (conc > 8.0 && x >= 0.8) ? std_gamma_grad_large(conc, x, lvar - lconc, 1.0 / conc)
                                               : std_gamma_grad<A>(conc, x);
        rec[SO_LQ] = lp;
        rec[SO_DQ] = d_v;
        // v = variate / beta:  dv/du_loc = sgg * loc,  dv/du_beta = sgg * loc - v
        rec[SO_A0] = sgg * loc;
        rec[SO_A1] = sgg * loc - v;
        rec[SO_B0] = d_conc * conc;
        rec[SO_B1] = d_conc * conc + d_rate * beta;
        if (s == S_B) {
            // model prior Gamma((bm/bs)^2, bm/bs^2)                                       cosmos.py:233-239
            const A bm = exp(ubm), bs = exp(ubs);
            const A ibs = 1.0 / bs;
            const A pc = (bm * ibs) * (bm * ibs), pr = bm * ibs * ibs;
            A plp, pd_v, pd_conc, pd_rate;
            gamma_density(v, pc, pr, lv, ubm - 2.0 * ubs, 2.0 * (ubm - ubs), plp, pd_v, pd_conc, pd_rate);
            extra[EX_LP] = plp;
            extra[EX_DP] = pd_v;
            // d/d u_bm, d/d u_bs (exp transforms: times bm, bs)
            extra[EX_GBM] = (pd_conc * (A(2) * bm * ibs * ibs) + pd_rate * ibs * ibs) * bm;
            extra[EX_GBS] = (pd_conc * (-A(2) * bm * bm * ibs * ibs * ibs) + pd_rate * (-A(2) * bm * ibs * ibs * ibs)) * bs;
        }
        return v;
    }
    // AffineBeta(mean, size, lo, hi): width cosmos.py:436-444, x :445-453, y :454-462
    const A half = A(mc.P + 1) / A(2), eps = mc.eps;
    A lo, hi;
    if (s < S_X) { lo = mc.width_min; hi = mc.width_max; } else { lo = -half; hi = half; }
    const Transformed<A> mean = t_interval<A>(u0, lo + eps, hi - eps, mc);
    const Transformed<A> size = t_greater_than<A>(u1, A(2));
    const AffBeta<A> d(mean.v, size.v, lo, hi);
    if (use_rng) {
        // the base draws are made in fp32 (their value is random); everything done with them is double
        GammaTrials trials;
        const A g1 = (A)sample_std_gamma_f32(*rng, trials, (float)d.c1), g2 = (A)sample_std_gamma_f32(*rng, trials, (float)d.c0);
        variate = beta01_from_gammas(g1, g2, mc);
    }
    const A v = d.clamp(d.low + d.scale * variate, mc);
    const A iscale = 1.0 / d.scale;
    const A x01 = (v - d.low) * iscale, y01 = 1.0 - x01;
    const A tot = d.c1 + d.c0;
    const A boundary = tot * x01 * y01;
    A lp, d_v, d_c1, d_c0, bg1, bg0;
    if (d.c1 > 10.0 && d.c0 > 10.0 && !(boundary < 2.5)) {
        // ---- usual regime: five logarithms, five reciprocals, shared by the density, its partials
        // and the pair of Rice expansions (beta_grad_alpha_mid for (x, c1, c0) and (1-x, c0, c1))
        const A lx = log(x01), ly = log(y01), lt = log(tot), l1 = log(d.c1), l0 = log(d.c0);
        const A it = 1.0 / tot, i1 = 1.0 / d.c1, i0 = 1.0 / d.c0, ix = 1.0 / x01, iy = 1.0 / y01;
        A lgt, pt, lg1, p1, lg0, p0;
        lgamma_digamma_known(tot, lt, it, lgt, pt);
        lgamma_digamma_known(d.c1, l1, i1, lg1, p1);
        lgamma_digamma_known(d.c0, l0, i0, lg0, p0);
        lp = (d.c1 - 1.0) * lx + (d.c0 - 1.0) * ly + lgt - lg1 - lg0 - log(d.scale);
        d_v = ((d.c1 - 1.0) * ix - (d.c0 - 1.0) * iy) * iscale;
        d_c1 = lx + pt - p1;
        d_c0 = ly + pt - p0;
        const A m1 = d.c1 * it;                                   // mean of x
        const A sd = sqrt(d.c1 * d.c0 / (tot + 1.0)) * it;
        if (m1 - 0.1 * sd <= x01 && x01 <= m1 + 0.1 * sd) {
            // Taylor patches around x = mean (both calls hit theirs together: |y - mean_y| = |x - mean_x|)
            bg1 = beta_grad_alpha_mid<A>(x01, d.c1, d.c0);
            bg0 = beta_grad_alpha_mid<A>(y01, d.c0, d.c1);
        } else {
            const A r2 = 1.4142135623730950488;
            const A sab = sqrt(d.c1 * d.c0 * it), isab = 1.0 / sab;
            const A st1 = 1.0 + i1 * (1.0 / 12.0) + i1 * i1 * (1.0 / 288.0);
            const A st0 = 1.0 + i0 * (1.0 / 12.0) + i0 * i0 * (1.0 / 288.0);
            const A stt = 1.0 + it * (1.0 / 12.0) + it * it * (1.0 / 288.0);
            const A stirling = st1 * st0 / stt;
            const A axbx = d.c0 * x01 - d.c1 * y01;
            const A iax = 1.0 / axbx;
            const A rab = sqrt(d.c1 * i0), irab = sqrt(d.c0 * i1);  // sqrt(alpha/beta) and its inverse
            const A i15 = iax * iax * it * sqrt(it) * (1.0 / r2);     // 1 / (sqrt2 total^1.5 axbx^2)
            const A L1 = l1 - lt - lx, L0 = l0 - lt - ly;
            const A t4b = d.c0 * L0 + d.c1 * L1;
            const A t4 = 1.0 / (t4b * sqrt(t4b));
            const A s8 = 2.8284271247461900976 * sab;
            {
                const A t1 = (-2.0 * d.c1 * d.c1 * y01 - d.c1 * d.c0 * y01 - x01 * d.c0 * d.c0) * (irab * i15);
                const A t3 = s8 * iax;
                bg1 = stirling * (-x01 * isab * (1.0 / r2)) * (t1 + 0.5 * L1 * (t3 + (x01 < m1 ? t4 : -t4)));
            }
            {
                const A t1 = (-2.0 * d.c0 * d.c0 * x01 - d.c1 * d.c0 * x01 - y01 * d.c1 * d.c1) * (rab * i15);
                const A t3 = -s8 * iax;
                bg0 = stirling * (-y01 * isab * (1.0 / r2)) * (t1 + 0.5 * L0 * (t3 + (y01 < d.c0 * it ? t4 : -t4)));
            }
        }
    } else {
        const BetaSite<A> q(v, d);
        lp = q.lp; d_v = q.d_v; d_c1 = q.d_c1; d_c0 = q.d_c0;
        beta_grad_pair<A>(q.x01, d.c1, d.c0, bg1, bg0);
    }
    const A dv_dc1 = d.scale * y01 * bg1;
    const A dv_dc0 = -d.scale * x01 * bg0;
    const A km = size.v * iscale * mean.d;                   // d c1 / d u_mean = - d c0 / d u_mean
    const A k1 = (mean.v - d.low) * iscale * size.d;         // d c1 / d u_size
    const A k0 = (d.low + d.scale - mean.v) * iscale * size.d;
    rec[SO_LQ] = lp;
    rec[SO_DQ] = d_v;
    rec[SO_A0] = (dv_dc1 - dv_dc0) * km;
    rec[SO_B0] = (d_c1 - d_c0) * km;
    rec[SO_A1] = dv_dc1 * k1 + dv_dc0 * k0;
    rec[SO_B1] = d_c1 * k1 + d_c0 * k0;
    return v;
}

// q(m_k = 1), q(m_k = 0), their logs and d q1 / d u for Bernoulli(sigmoid(u)) with the reference's
// clamps (unit_interval transform clamp + Bernoulli probs clamp to [eps, 1 - eps]), evaluated on the
// logit so that fp32 never produces an exact 0 or 1.                                  cosmos.py:419-425, 481-485
template <typename F> struct SpotPresence {
    F q1, q0, lq1, lq0, dq1;
    TQ_HD SpotPresence(F u, const ModelConst& mc) {
        using R = Real<F>;
        const F lim = F(mc.logit_lim);
        const bool inside = (u >= -lim) && (u <= lim);
        const F uc = R::min(R::max(u, -lim), lim);
        const F e = R::exp(-R::abs(uc));
        const F sp = R::log1p(e);                 // softplus(-|uc|)
        const F big = F(1) / (F(1) + e), small = e / (F(1) + e);
        if (uc >= F(0)) { q1 = big; q0 = small; lq1 = -sp; lq0 = -uc - sp; }
        else            { q1 = small; q0 = big; lq1 = uc - sp; lq0 = -sp; }
        dq1 = inside ? q1 * q0 : F(0);
    }
};

template <typename F> TQ_HD void presence_weights(const F (&q1)[kK], const F (&q0)[kK], F (&qm)[kM]) {
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        F q = F(1);
#pragma unroll
        for (int k = 0; k < kK; ++k) q *= ((m >> k) & 1) ? q1[k] : q0[k];
        qm[m] = q;
    }
}

// Outputs of unit_post for one unit.
template <typename F> struct UnitGrads {
    F g[NLOCAL];   // d ELBO_unit / d unconstrained (unit-level, unscaled, before the mask)
    F acc[NACC];   // contributions to the per-channel accumulators (unscaled, before the mask)
};

// ---- unit_post -------------------------------------------------------------------------------------------
//   rec        (NREC) site records of this unit (rows s * NSO + SO_*, then the NEX extra rows)
//   sample     guide samples (the values the likelihood kernel used)
//   L          log-likelihood of the 4 configurations from K1
//   gs         d (sum_m q(m) L(m)) / d sample from K1 (weights W(m) = q(m))
//   g_rate     d (sum_m q(m) L(m)) / d (1/gain) from K1
//   u_mp       unconstrained m_probs of the K spots; u_bm/u_bs unconstrained AOI-level parameters
//   first_frame  this unit carries the AOI-level prior terms of its (AOI, channel)
// (rec / sample / gs: anything indexable -- plain arrays, or strided views of shared memory in the fused kernel, which
// lets the compiler load each entry where it is used instead of holding all 76 in registers)
template <typename F, typename RecT, typename SampT, typename GsT>
TQ_HD void unit_post(const RecT& rec, const SampT& sample, const F (&L)[kM], const GsT& gs,
                     F g_rate, const F (&u_mp)[kK], F u_bm, F u_bs, const ModelConst& mc,
                     const GlobalTables<F>& gt, int c, bool ontarget, bool first_frame, UnitGrads<F>& out) {
    using R = Real<F>;
    const ChannelTables<F>& ct = gt.ch[c];
    const F half = F(mc.P + 1) / F(2);
#pragma unroll
    for (int i = 0; i < NLOCAL; ++i) out.g[i] = F(0);
#pragma unroll
    for (int i = 0; i < NACC; ++i) out.acc[i] = F(0);
    auto R_ = [&](int s, int j) -> F { return rec[s * NSO + j]; };

    F q1[kK], q0[kK], lq1[kK], lq0[kK], dq1[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const SpotPresence<F> sp(u_mp[k], mc);
        q1[k] = sp.q1; q0[k] = sp.q0; lq1[k] = sp.lq1; lq0[k] = sp.lq0; dq1[k] = sp.dq1;
    }

    // ---- background ------------------------------------------------------------------------------------------
    auto EX_ = [&](int j) -> F { return rec[NSAMP * NSO + j]; };
    F elbo = EX_(EX_LP) - R_(S_B, SO_LQ);
    {
        const F G = EX_(EX_DP) - R_(S_B, SO_DQ) + gs[S_B];
        out.g[LP_B_LOC] = G * R_(S_B, SO_A0) - R_(S_B, SO_B0);
        out.g[LP_B_BETA] = G * R_(S_B, SO_A1) - R_(S_B, SO_B1);
        out.g[LP_BM] = EX_(EX_GBM);
        out.g[LP_BS] = EX_(EX_GBS);
    }

    // ---- per-spot terms that do not depend on (z, theta) ------------------------------------------------------
    F spot_term[kK], lxy1[kK], dlxy1_dx[kK], dlxy1_dy[kK], dlxy1_dsize[kK];
    const F cs1 = gt.size1 * F(0.5) - F(1);
    const F hs = F(mc.height_std);
    const F lp_w = -R::log(F(mc.width_max) - F(mc.width_min));   // AffineBeta(1.5, 2, ..) = uniform  :274-282
    const F c_hn = -R::log(hs) + F(0.5) * R::log(F(2) / F(3.14159265358979323846));
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F h = sample[S_H + k], x = sample[S_X + k], y = sample[S_Y + k];
        const F lp_h = -h * h / (F(2) * hs * hs) + c_hn;          // HalfNormal(height_std)           :270-273
        spot_term[k] = lp_h + lp_w - R_(S_H + k, SO_LQ) - R_(S_W + k, SO_LQ) - R_(S_X + k, SO_LQ) - R_(S_Y + k, SO_LQ);
        // target-specific prior AffineBeta(0, size1, -half, half) on x and y                 :283-300
        const F tx = x / half, ty = y / half;
        const F ox = F(1) - tx * tx, oy = F(1) - ty * ty;
        const F lsum = R::log1p(-tx * tx) + R::log1p(-ty * ty);
        lxy1[k] = cs1 * lsum + gt.cxy1;
        dlxy1_dx[k] = cs1 * (-F(2) * tx / ox) / half;
        dlxy1_dy[k] = cs1 * (-F(2) * ty / oy) / half;
        dlxy1_dsize[k] = F(0.5) * lsum + gt.dcxy1;
    }

    // ---- enumerated part: T(m) = logsumexp over (z, theta), weighted by q(m)   ---------------
    F sum_qC = F(0);
    F gmp[kK], wk_x[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) gmp[k] = wk_x[k] = F(0);
    const int ot = ontarget ? 1 : 0;
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        F lj[kZ][kTheta];
        F mx = -R::inf();
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                F v = ct.logpz[ot][z] + ct.logptheta[z][th];
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const int mk = (m >> k) & 1;
                    v += ct.logpm[th][k][mk];
                    if (mk) v += (th == k + 1) ? lxy1[k] : gt.lxy0;
                }
                lj[z][th] = v;
                mx = R::max(mx, v);
            }
        F se = F(0);
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                lj[z][th] = R::exp_fast(lj[z][th] - mx);
                se += lj[z][th];
            }
        const F T = mx + R::log_fast(se);
        F q = F(1), lq = F(0), Cm = T + L[m];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int mk = (m >> k) & 1;
            q *= mk ? q1[k] : q0[k];
            lq += mk ? lq1[k] : lq0[k];
            if (mk) Cm += spot_term[k];
        }
        Cm -= lq;
        sum_qC += q * Cm;
        // d / d u_mprobs_k: q(m) (m_k - q1_k) (C(m) - 1)   [= q(m) dlog q(m_k)/dp * dp/du * (C - 1)]
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int mk = (m >> k) & 1;
            gmp[k] += q * (mk ? q0[k] : -q1[k]) * (Cm - F(1));
        }
        // posterior responsibilities r(z, theta | m) drive the table gradients
        const F qi = q / se;
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                const F r = qi * lj[z][th];
                if (ontarget) out.acc[ACC_LOGPZ + z] += r;
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const int mk = (m >> k) & 1;
                    out.acc[ACC_LOGPM + (th * kK + k) * 2 + mk] += r;
                    if (mk && th == k + 1) wk_x[k] += r;
                }
            }
    }
    elbo += sum_qC;
    out.acc[ACC_ELBO_FRAME] = elbo;
    out.acc[ACC_RATE] = g_rate;
#pragma unroll
    for (int k = 0; k < kK; ++k) out.acc[ACC_SIZE1] += wk_x[k] * dlxy1_dsize[k];

    // ---- parameter gradients through the site maps ------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        out.g[LP_M_PROBS + k] = dq1[k] > F(0) ? gmp[k] : F(0);
        const F qk = q1[k];  // sum_m q(m) m_k
        const F Gh = qk * (-sample[S_H + k] / (hs * hs) - R_(S_H + k, SO_DQ)) + gs[S_H + k];
        const F Gw = -qk * R_(S_W + k, SO_DQ) + gs[S_W + k];
        const F Gx = -qk * R_(S_X + k, SO_DQ) + gs[S_X + k] + wk_x[k] * dlxy1_dx[k];
        const F Gy = -qk * R_(S_Y + k, SO_DQ) + gs[S_Y + k] + wk_x[k] * dlxy1_dy[k];
        out.g[LP_H_LOC + k] = Gh * R_(S_H + k, SO_A0) - qk * R_(S_H + k, SO_B0);
        out.g[LP_H_BETA + k] = Gh * R_(S_H + k, SO_A1) - qk * R_(S_H + k, SO_B1);
        out.g[LP_W_MEAN + k] = Gw * R_(S_W + k, SO_A0) - qk * R_(S_W + k, SO_B0);
        out.g[LP_W_SIZE + k] = Gw * R_(S_W + k, SO_A1) - qk * R_(S_W + k, SO_B1);
        out.g[LP_X_MEAN + k] = Gx * R_(S_X + k, SO_A0) - qk * R_(S_X + k, SO_B0);
        out.g[LP_Y_MEAN + k] = Gy * R_(S_Y + k, SO_A0) - qk * R_(S_Y + k, SO_B0);
        out.g[LP_SIZE + k] = (Gx * R_(S_X + k, SO_A1) - qk * R_(S_X + k, SO_B1))
                           + (Gy * R_(S_Y + k, SO_A1) - qk * R_(S_Y + k, SO_B1));
    }

    // ---- AOI-level prior terms, carried by the first minibatch frame of each (AOI, channel) ---------------------
    // HalfNormal(bm; bg_mean_std) + HalfNormal(bs; bg_std_std), Delta guide                    :221-227, 397-404
    if (first_frame) {
        const F bm = R::exp(u_bm), bs = R::exp(u_bs);
        const F s1 = F(mc.bg_mean_std), s2 = F(mc.bg_std_std);
        const F c0 = F(0.5) * R::log(F(2) / F(3.14159265358979323846));
        out.acc[ACC_ELBO_AOI] = (-bm * bm / (F(2) * s1 * s1) - R::log(s1) + c0) + (-bs * bs / (F(2) * s2 * s2) - R::log(s2) + c0);
    }
}

// ---- posterior of (z, theta) for one guide draw (models/cosmos.py:609-672, "next" row N1) -------------------
// p(z, theta | m, x, y, globals) from the model's m/x/y/z/theta log-probs (data hidden), averaged
// over m with the guide weights q(m):  pz[z] = sum_theta,  pth[k] = sum_z at theta = k + 1.
template <typename F>
TQ_HD void unit_ztheta_posterior(F x_[kK], F y_[kK], const F (&u_mp)[kK], const ModelConst& mc,
                                 const GlobalTables<F>& gt, int c, bool ontarget, F (&pz)[kZ], F (&pth)[kK]) {
    using R = Real<F>;
    const ChannelTables<F>& ct = gt.ch[c];
    const F half = F(mc.P + 1) / F(2);
    const F cs1 = gt.size1 * F(0.5) - F(1);
    F q1[kK], q0[kK], lxy1[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const SpotPresence<F> sp(u_mp[k], mc);
        q1[k] = sp.q1; q0[k] = sp.q0;
        const F tx = x_[k] / half, ty = y_[k] / half;
        lxy1[k] = cs1 * (R::log1p(-tx * tx) + R::log1p(-ty * ty)) + gt.cxy1;
    }
#pragma unroll
    for (int z = 0; z < kZ; ++z) pz[z] = F(0);
#pragma unroll
    for (int k = 0; k < kK; ++k) pth[k] = F(0);
    const int ot = ontarget ? 1 : 0;
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        F lj[kZ][kTheta];
        F mx = -R::inf();
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                F v = ct.logpz[ot][z] + ct.logptheta[z][th];
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const int mk = (m >> k) & 1;
                    v += ct.logpm[th][k][mk];
                    if (mk) v += (th == k + 1) ? lxy1[k] : gt.lxy0;
                }
                lj[z][th] = v;
                mx = R::max(mx, v);
            }
        F se = F(0);
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                lj[z][th] = R::exp_fast(lj[z][th] - mx);
                se += lj[z][th];
            }
        F q = F(1);
#pragma unroll
        for (int k = 0; k < kK; ++k) q *= ((m >> k) & 1) ? q1[k] : q0[k];
        const F qi = q / se;
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                const F r = qi * lj[z][th];
                pz[z] += r;
                if (th > 0) pth[th - 1] += r;
            }
    }
}

// gradient of the AOI-level prior terms w.r.t. the unconstrained (bm, bs); scaled by s_N (not s_N s_F)
TQ_HD void aoi_prior_grad(double u_bm, double u_bs, const ModelConst& mc, double& g_bm, double& g_bs) {
    const double bm = exp(u_bm), bs = exp(u_bs);
    g_bm = -bm / (mc.bg_mean_std * mc.bg_mean_std) * bm;
    g_bs = -bs / (mc.bg_std_std * mc.bg_std_std) * bs;
}

}  // namespace tq
