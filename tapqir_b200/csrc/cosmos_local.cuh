// Per-unit ("local") part of the cosmos ELBO: everything models/cosmos.py:216-327 (model) and
// :393-462 (guide) do for one (AOI, frame, channel) patch except the pixel likelihood itself,
// written out with its analytic reverse mode.  SURVEY.md App. A.3 gives the ELBO, this file is its
// per-unit summand and the chain rule down to the unconstrained variational parameters.
//
//   local_pre   : unconstrained params -> constrained -> guide samples (replayed variates or
//                 in-kernel Philox) -> q(m) weights for the likelihood kernel
//   local_post  : priors, guide log-densities, the (z, theta) log-sum-exp T(m), ELBO summand,
//                 d/d(samples) (+ likelihood part from K1) -> reparameterisation -> d/d(unconstrained)
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
    int P;
};

// Per-channel tables derived from the sampled globals (built by globals_pre, cosmos_globals.cuh).
template <typename A> struct ChannelTables {
    A logpz[2][kZ];             // [is_ontarget][z]           cosmos.py:242-246, util.py:133-151
    A logptheta[2][kTheta];     // [min(z,1)][theta]          cosmos.py:247-255, util.py:154-173
    A logpm[kTheta][kK][2];     // [theta][k][m_k]            cosmos.py:262-267, util.py:94-130
};
template <typename A> struct GlobalTables {
    A gain, rate, log_rate;
    A size1;        // AffineBeta sample size of the target-specific spot: ((P+1)/(2 proximity))^2 - 1
    A lnorm1;       // lgamma(size1) - 2 lgamma(size1/2)      (Beta normaliser, per axis)
    A dlnorm1;      // d lnorm1 / d size1 = psi(size1) - psi(size1/2)
    ChannelTables<A> ch[kMaxC];
};

// indices into the per-channel accumulator vector reduced over units
enum {
    ACC_ELBO_FRAME = 0,   // sum mu_n * (frame-level ELBO summand), unscaled
    ACC_ELBO_AOI = 1,     // sum mu_n * (AOI-level prior terms), once per (AOI, channel)
    ACC_LOGPZ = 2,        // [kZ]           d/d logpz[ontarget=1][z]
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
template <typename A> TQ_HD Transformed<A> t_unit_interval(A u, const ModelConst& mc) {
    bool act;
    const A s = sigmoid_clamped(u, mc, act);
    return {s, act ? s * (A(1) - s) : A(0)};
}

// The 18 AOI-local variational parameters of one unit.  Order = order of the flat gradient record.
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

template <typename A> struct UnitParams { Transformed<A> p[NLOCAL]; };

// unconstrained -> constrained for one unit (constraints: models/cosmos.py:481-485, 530-598)
template <typename A> TQ_HD void transform_unit(const A (&u)[NLOCAL], const ModelConst& mc, UnitParams<A>& out) {
    const A half = A(mc.P + 1) / A(2), eps = A(mc.eps);
    out.p[LP_BM] = t_positive(u[LP_BM]);
    out.p[LP_BS] = t_positive(u[LP_BS]);
    out.p[LP_B_LOC] = t_positive(u[LP_B_LOC]);
    out.p[LP_B_BETA] = t_positive(u[LP_B_BETA]);
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        out.p[LP_M_PROBS + k] = t_unit_interval(u[LP_M_PROBS + k], mc);
        out.p[LP_H_LOC + k] = t_positive(u[LP_H_LOC + k]);
        out.p[LP_H_BETA + k] = t_positive(u[LP_H_BETA + k]);
        out.p[LP_W_MEAN + k] = t_interval(u[LP_W_MEAN + k], A(mc.width_min) + eps, A(mc.width_max) - eps, mc);
        out.p[LP_W_SIZE + k] = t_greater_than(u[LP_W_SIZE + k], A(2));
        out.p[LP_X_MEAN + k] = t_interval(u[LP_X_MEAN + k], -half + eps, half - eps, mc);
        out.p[LP_Y_MEAN + k] = t_interval(u[LP_Y_MEAN + k], -half + eps, half - eps, mc);
        out.p[LP_SIZE + k] = t_greater_than(u[LP_SIZE + k], A(2));
    }
}

// Guide samples of one unit: background, and per spot height, width, x, y.  Order of the record.
enum { S_B = 0, S_H = 1, S_W = S_H + kK, S_X = S_W + kK, S_Y = S_X + kK, NSAMP = S_Y + kK };  // 9 at K=2

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

// ---- local_pre ---------------------------------------------------------------------------------
// variates: standard-gamma draws for S_B, S_H+k; (0,1) Beta draws for S_W.., S_X.., S_Y.. (replay),
// or nullptr-equivalent `use_rng` to draw them here.
template <typename A>
TQ_HD void local_pre(const UnitParams<A>& up, const ModelConst& mc, bool use_rng, Philox* rng,
                     A (&variate)[NSAMP], A (&sample)[NSAMP], A (&qm)[kM]) {
    const A half = A(mc.P + 1) / A(2);
    const A tiny = A(mc.tiny);
    // background ~ Gamma(b_loc * b_beta, b_beta)                                    cosmos.py:408-415
    {
        const A conc = up.p[LP_B_LOC].v * up.p[LP_B_BETA].v;
        if (use_rng) variate[S_B] = Real<A>::max(sample_std_gamma<A>(*rng, conc), tiny);
        sample[S_B] = Real<A>::max(variate[S_B] / up.p[LP_B_BETA].v, tiny);
    }
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        // height ~ Gamma(h_loc * h_beta, h_beta)                                     cosmos.py:428-435
        const A conc = up.p[LP_H_LOC + k].v * up.p[LP_H_BETA + k].v;
        if (use_rng) variate[S_H + k] = Real<A>::max(sample_std_gamma<A>(*rng, conc), tiny);
        sample[S_H + k] = Real<A>::max(variate[S_H + k] / up.p[LP_H_BETA + k].v, tiny);
        // width, x, y ~ AffineBeta                                                   cosmos.py:436-462
        const AffBeta<A> dw(up.p[LP_W_MEAN + k].v, up.p[LP_W_SIZE + k].v, A(mc.width_min), A(mc.width_max));
        const AffBeta<A> dx(up.p[LP_X_MEAN + k].v, up.p[LP_SIZE + k].v, -half, half);
        const AffBeta<A> dy(up.p[LP_Y_MEAN + k].v, up.p[LP_SIZE + k].v, -half, half);
        if (use_rng) {
            A g1 = sample_std_gamma<A>(*rng, dw.c1), g2 = sample_std_gamma<A>(*rng, dw.c0);
            variate[S_W + k] = beta01_from_gammas(g1, g2, mc);
            g1 = sample_std_gamma<A>(*rng, dx.c1); g2 = sample_std_gamma<A>(*rng, dx.c0);
            variate[S_X + k] = beta01_from_gammas(g1, g2, mc);
            g1 = sample_std_gamma<A>(*rng, dy.c1); g2 = sample_std_gamma<A>(*rng, dy.c0);
            variate[S_Y + k] = beta01_from_gammas(g1, g2, mc);
        }
        sample[S_W + k] = dw.clamp(dw.low + dw.scale * variate[S_W + k], mc);
        sample[S_X + k] = dx.clamp(dx.low + dx.scale * variate[S_X + k], mc);
        sample[S_Y + k] = dy.clamp(dy.low + dy.scale * variate[S_Y + k], mc);
    }
    // q(m) = prod_k Bernoulli(m_k; m_probs_k) with torch's eps clamp                   cosmos.py:419-425
    A q1[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const A p = up.p[LP_M_PROBS + k].v;
        q1[k] = Real<A>::min(Real<A>::max(p, A(mc.eps)), A(1) - A(mc.eps));
    }
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        A q = A(1);
#pragma unroll
        for (int k = 0; k < kK; ++k) q *= ((m >> k) & 1) ? q1[k] : (A(1) - q1[k]);
        qm[m] = q;
    }
}

// ---- densities and their partials ----------------------------------------------------------------
template <typename A> struct GammaSite {
    // value v ~ Gamma(conc, rate): log-density and partials
    A lp, d_v, d_conc, d_rate;
    TQ_HD GammaSite(A v, A conc, A rate) {
        using R = Real<A>;
        const A lv = R::log(v), lr = R::log(rate);
        lp = conc * lr + (conc - A(1)) * lv - rate * v - R::lgamma(conc);
        d_v = (conc - A(1)) / v - rate;
        d_conc = lr + lv - digamma(conc);
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
        const A pt = digamma(tot);
        lp = (d.c1 - A(1)) * l1 + (d.c0 - A(1)) * l0 + R::lgamma(tot) - R::lgamma(d.c1) - R::lgamma(d.c0) - R::log(d.scale);
        d_v = ((d.c1 - A(1)) / x01 - (d.c0 - A(1)) / (A(1) - x01)) / d.scale;
        d_c1 = l1 + pt - digamma(d.c1);
        d_c0 = l0 + pt - digamma(d.c0);
    }
};

// Outputs of local_post for one unit.
template <typename A> struct UnitGrads {
    A g[NLOCAL];   // d ELBO_unit / d unconstrained (unit-level, unscaled, before the mask)
    A acc[NACC];   // contributions to the per-channel accumulators (unscaled, before the mask)
};

// ---- local_post --------------------------------------------------------------------------------
//   sample     guide samples (same values the likelihood kernel used)
//   L          log-likelihood of the 4 configurations from K1
//   gs         d (sum_m q(m) L(m)) / d sample from K1 (weights W(m) = q(m))
//   g_rate     d (sum_m q(m) L(m)) / d (1/gain) from K1
//   first_frame  this unit carries the AOI-level prior terms of its (AOI, channel)
template <typename A>
TQ_HD void local_post(const UnitParams<A>& up, const ModelConst& mc, const GlobalTables<A>& gt, int c,
                      bool ontarget, bool first_frame, const A (&sample)[NSAMP], const A (&L)[kM],
                      const A (&gs)[NSAMP], A g_rate, UnitGrads<A>& out) {
    using R = Real<A>;
    const ChannelTables<A>& ct = gt.ch[c];
    const A half = A(mc.P + 1) / A(2);
#pragma unroll
    for (int i = 0; i < NLOCAL; ++i) out.g[i] = A(0);
#pragma unroll
    for (int i = 0; i < NACC; ++i) out.acc[i] = A(0);

    // q(m_k), log q(m_k)
    A q1[kK], lq1[kK], lq0[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        q1[k] = R::min(R::max(up.p[LP_M_PROBS + k].v, A(mc.eps)), A(1) - A(mc.eps));
        lq1[k] = R::log(q1[k]);
        lq0[k] = R::log1p(-q1[k]);
    }

    // ---- background: model Gamma((bm/bs)^2, bm/bs^2), guide Gamma(b_loc b_beta, b_beta)  :233-239, 408-415
    const A bm = up.p[LP_BM].v, bs = up.p[LP_BS].v;
    const A b = sample[S_B];
    const A pc = (bm / bs) * (bm / bs), pr = bm / (bs * bs);
    const GammaSite<A> pb(b, pc, pr);
    const A qc = up.p[LP_B_LOC].v * up.p[LP_B_BETA].v, qr = up.p[LP_B_BETA].v;
    const GammaSite<A> qb(b, qc, qr);
    A elbo = pb.lp - qb.lp;
    // total derivative w.r.t. the sample b, then through the reparameterisation
    const A Gb = pb.d_v - qb.d_v + gs[S_B];
    {
        const A variate = b * qr;
        const A db_dconc = std_gamma_grad<A>(qc, variate) / qr;
        const A db_drate = -b / qr;
        const A g_conc = Gb * db_dconc - qb.d_conc;
        const A g_rate_q = Gb * db_drate - qb.d_rate;
        // conc = loc * beta, rate = beta
        out.g[LP_B_LOC] = g_conc * up.p[LP_B_BETA].v * up.p[LP_B_LOC].d;
        out.g[LP_B_BETA] = (g_conc * up.p[LP_B_LOC].v + g_rate_q) * up.p[LP_B_BETA].d;
        // prior parameters: conc_p = (bm/bs)^2, rate_p = bm/bs^2
        const A g_bm = pb.d_conc * (A(2) * bm / (bs * bs)) + pb.d_rate / (bs * bs);
        const A g_bs = pb.d_conc * (-A(2) * bm * bm / (bs * bs * bs)) + pb.d_rate * (-A(2) * bm / (bs * bs * bs));
        out.g[LP_BM] = g_bm * up.p[LP_BM].d;
        out.g[LP_BS] = g_bs * up.p[LP_BS].d;
    }

    // ---- per-spot continuous sites ------------------------------------------------------------------
    A spot_term[kK];           // log p - log q of (h, w, x, y)_k excluding the theta-dependent x,y prior
    A lxy1[kK], dlxy1_dx[kK], dlxy1_dy[kK], dlxy1_dsize[kK];  // target-specific x,y prior and partials
    const A lxy0 = -A(2) * R::log(A(2) * half);  // size 2 => Beta(1,1): uniform on the patch      :283-300
    A Gh[kK], Gw[kK], Gx[kK], Gy[kK];            // running d/d sample
    // guide sites kept for the chain rule
    A h_dconc[kK], h_drate[kK], w_dc1[kK], w_dc0[kK], x_dc1[kK], x_dc0[kK], y_dc1[kK], y_dc0[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const A h = sample[S_H + k], w = sample[S_W + k], x = sample[S_X + k], y = sample[S_Y + k];
        const A hs = A(mc.height_std);
        // HalfNormal(h; height_std)                                                            :270-273
        const A lp_h = -h * h / (A(2) * hs * hs) - R::log(hs) + A(0.5) * R::log(A(2) / A(3.14159265358979323846));
        // AffineBeta(1.5, 2, wmin, wmax) is Beta(1,1): uniform                                  :274-282
        const A lp_w = -R::log(A(mc.width_max) - A(mc.width_min));
        const GammaSite<A> qh(h, up.p[LP_H_LOC + k].v * up.p[LP_H_BETA + k].v, up.p[LP_H_BETA + k].v);
        const AffBeta<A> dw(up.p[LP_W_MEAN + k].v, up.p[LP_W_SIZE + k].v, A(mc.width_min), A(mc.width_max));
        const AffBeta<A> dx(up.p[LP_X_MEAN + k].v, up.p[LP_SIZE + k].v, -half, half);
        const AffBeta<A> dy(up.p[LP_Y_MEAN + k].v, up.p[LP_SIZE + k].v, -half, half);
        const BetaSite<A> qw(w, dw), qx(x, dx), qy(y, dy);
        spot_term[k] = lp_h + lp_w - qh.lp - qw.lp - qx.lp - qy.lp;
        // d spot_term / d sample (weighted by q(m_k = 1) below)
        Gh[k] = -h / (hs * hs) - qh.d_v;
        Gw[k] = -qw.d_v;
        Gx[k] = -qx.d_v;
        Gy[k] = -qy.d_v;
        h_dconc[k] = qh.d_conc; h_drate[k] = qh.d_rate;
        w_dc1[k] = qw.d_c1; w_dc0[k] = qw.d_c0;
        x_dc1[k] = qx.d_c1; x_dc0[k] = qx.d_c0;
        y_dc1[k] = qy.d_c1; y_dc0[k] = qy.d_c0;
        // target-specific prior AffineBeta(0, size1, -half, half) on x and y                   :283-300
        const A cs = gt.size1 / A(2);
        const A Lx = R::log(qx.x01) + R::log(A(1) - qx.x01);
        const A Ly = R::log(qy.x01) + R::log(A(1) - qy.x01);
        lxy1[k] = (cs - A(1)) * (Lx + Ly) + A(2) * gt.lnorm1 + lxy0;
        dlxy1_dx[k] = (cs - A(1)) * (A(1) / qx.x01 - A(1) / (A(1) - qx.x01)) / (A(2) * half);
        dlxy1_dy[k] = (cs - A(1)) * (A(1) / qy.x01 - A(1) / (A(1) - qy.x01)) / (A(2) * half);
        dlxy1_dsize[k] = A(0.5) * (Lx + Ly) + A(2) * gt.dlnorm1;
    }

    // ---- enumerated part: T(m) = logsumexp over (z, theta), weighted by q(m)   ----
    A sum_qC = A(0);
    A gmp[kK];  // d / d (clamped m_probs_k)
#pragma unroll
    for (int k = 0; k < kK; ++k) gmp[k] = A(0);
    A wk_x[kK], wk_s[kK];  // sum_m q(m) m_k R_k(m): weight of the target-specific prior derivative
#pragma unroll
    for (int k = 0; k < kK; ++k) wk_x[k] = wk_s[k] = A(0);
    const int ot = ontarget ? 1 : 0;
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        A lj[kZ][kTheta];
        A mx = -R::inf();
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                A v = ct.logpz[ot][z] + ct.logptheta[z][th];
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const int mk = (m >> k) & 1;
                    v += ct.logpm[th][k][mk];
                    if (mk) v += (th == k + 1) ? lxy1[k] : lxy0;
                }
                lj[z][th] = v;
                mx = R::max(mx, v);
            }
        A se = A(0);
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                lj[z][th] = R::exp(lj[z][th] - mx);
                se += lj[z][th];
            }
        const A T = mx + R::log(se);
        A q = A(1), lq = A(0), Cm = T + L[m];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int mk = (m >> k) & 1;
            q *= mk ? q1[k] : (A(1) - q1[k]);
            lq += mk ? lq1[k] : lq0[k];
            if (mk) Cm += spot_term[k];
        }
        Cm -= lq;
        sum_qC += q * Cm;
        // d / d m_probs_k: q(m) * dlog q(m_k)/dp * (C(m) - 1)
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int mk = (m >> k) & 1;
            gmp[k] += q * (mk ? A(1) / q1[k] : -A(1) / (A(1) - q1[k])) * (Cm - A(1));
        }
        // posterior responsibilities r(z, theta | m) drive the table gradients
        const A qi = q / se;
#pragma unroll
        for (int z = 0; z < kZ; ++z)
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                const A r = qi * lj[z][th];
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

    // ---- chain rule for the spot sites ---------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        // m_probs: clamp (identity inside) -> sigmoid
        const A p = up.p[LP_M_PROBS + k].v;
        const bool inside = (p >= A(mc.eps)) && (p <= A(1) - A(mc.eps));
        out.g[LP_M_PROBS + k] = inside ? gmp[k] * up.p[LP_M_PROBS + k].d : A(0);

        const A qk = q1[k];  // sum_m q(m) m_k
        // height
        {
            const A h = sample[S_H + k];
            const A conc = up.p[LP_H_LOC + k].v * up.p[LP_H_BETA + k].v, rate = up.p[LP_H_BETA + k].v;
            const A G = qk * Gh[k] + gs[S_H + k];
            const A dv_dconc = std_gamma_grad<A>(conc, h * rate) / rate;
            const A g_conc = G * dv_dconc - qk * h_dconc[k];
            const A g_rt = G * (-h / rate) - qk * h_drate[k];
            out.g[LP_H_LOC + k] = g_conc * rate * up.p[LP_H_LOC + k].d;
            out.g[LP_H_BETA + k] = (g_conc * up.p[LP_H_LOC + k].v + g_rt) * up.p[LP_H_BETA + k].d;
        }
        // width / x / y share the AffineBeta chain
        A g_size = A(0);
        auto beta_chain = [&](A G, A v, const AffBeta<A>& d, A dq_c1, A dq_c0, A mean, A size, A& g_mean, A& g_sz) {
            const A x01 = (v - d.low) / d.scale;
            const A tot = d.c1 + d.c0;
            const A dv_dc1 = d.scale * (A(1) - x01) * beta_grad<A>(x01, d.c1, tot);
            const A dv_dc0 = -d.scale * x01 * beta_grad<A>(A(1) - x01, d.c0, tot);
            const A g_c1 = G * dv_dc1 - qk * dq_c1;
            const A g_c0 = G * dv_dc0 - qk * dq_c0;
            g_mean = (g_c1 - g_c0) * size / d.scale;
            g_sz = g_c1 * (mean - d.low) / d.scale + g_c0 * (d.low + d.scale - mean) / d.scale;
        };
        {
            const AffBeta<A> dw(up.p[LP_W_MEAN + k].v, up.p[LP_W_SIZE + k].v, A(mc.width_min), A(mc.width_max));
            A gm, gsz;
            beta_chain(qk * Gw[k] + gs[S_W + k], sample[S_W + k], dw, w_dc1[k], w_dc0[k], up.p[LP_W_MEAN + k].v,
                       up.p[LP_W_SIZE + k].v, gm, gsz);
            out.g[LP_W_MEAN + k] = gm * up.p[LP_W_MEAN + k].d;
            out.g[LP_W_SIZE + k] = gsz * up.p[LP_W_SIZE + k].d;
        }
        {
            const AffBeta<A> dx(up.p[LP_X_MEAN + k].v, up.p[LP_SIZE + k].v, -half, half);
            A gm, gsz;
            beta_chain(qk * Gx[k] + gs[S_X + k] + wk_x[k] * dlxy1_dx[k], sample[S_X + k], dx, x_dc1[k], x_dc0[k],
                       up.p[LP_X_MEAN + k].v, up.p[LP_SIZE + k].v, gm, gsz);
            out.g[LP_X_MEAN + k] = gm * up.p[LP_X_MEAN + k].d;
            g_size += gsz;
        }
        {
            const AffBeta<A> dy(up.p[LP_Y_MEAN + k].v, up.p[LP_SIZE + k].v, -half, half);
            A gm, gsz;
            beta_chain(qk * Gy[k] + gs[S_Y + k] + wk_x[k] * dlxy1_dy[k], sample[S_Y + k], dy, y_dc1[k], y_dc0[k],
                       up.p[LP_Y_MEAN + k].v, up.p[LP_SIZE + k].v, gm, gsz);
            out.g[LP_Y_MEAN + k] = gm * up.p[LP_Y_MEAN + k].d;
            g_size += gsz;
        }
        out.g[LP_SIZE + k] = g_size * up.p[LP_SIZE + k].d;
    }
    (void)wk_s;

    // ---- AOI-level prior terms, carried by the first minibatch frame of each (AOI, channel) ---------
    // HalfNormal(bm; bg_mean_std) + HalfNormal(bs; bg_std_std), Delta guide                    :221-227, 397-404
    if (first_frame) {
        const A s1 = A(mc.bg_mean_std), s2 = A(mc.bg_std_std);
        const A c0 = A(0.5) * R::log(A(2) / A(3.14159265358979323846));
        out.acc[ACC_ELBO_AOI] = (-bm * bm / (A(2) * s1 * s1) - R::log(s1) + c0) + (-bs * bs / (A(2) * s2 * s2) - R::log(s2) + c0);
    }
}

// gradient of the AOI-level prior terms w.r.t. the unconstrained (bm, bs); scaled by s_N (not s_N s_F)
template <typename A> TQ_HD void aoi_prior_grad(const UnitParams<A>& up, const ModelConst& mc, A& g_bm, A& g_bs) {
    const A s1 = A(mc.bg_mean_std), s2 = A(mc.bg_std_std);
    g_bm = -up.p[LP_BM].v / (s1 * s1) * up.p[LP_BM].d;
    g_bs = -up.p[LP_BS].v / (s2 * s2) * up.p[LP_BS].d;
}

}  // namespace tq
