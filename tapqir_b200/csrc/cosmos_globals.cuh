// Global ("replicated") part of the cosmos step: the four global latent sites gain, pi, lamda,
// proximity (models/cosmos.py:170-191 model, :342-368 guide), the prior tables derived from them
// (distributions/util.py:67-173) and the reverse mode from the per-channel accumulators back to the
// 4 + 5Q unconstrained global parameters.  Always evaluated in double by a single thread: it is a
// few hundred flops per step, identical on every rank.
#pragma once
#include "cosmos_local.cuh"

namespace tq {

// flat layout of the global variational parameters (and of their gradient / Adam moments)
struct GlobalLayout {
    int Q;
    bool hmm = false;   // hmm variant (models/hmm.py): pi_* are init_*, plus trans_mean (Q,2,2) and trans_size (Q,2)
    TQ_HD int gain_loc() const { return 0; }
    TQ_HD int gain_beta() const { return 1; }
    TQ_HD int prox_loc() const { return 2; }
    TQ_HD int prox_size() const { return 3; }
    TQ_HD int pi_mean(int q, int z) const { return 4 + q * kZ + z; }
    TQ_HD int pi_size(int q) const { return 4 + kZ * Q + q; }
    TQ_HD int lamda_loc(int q) const { return 4 + (kZ + 1) * Q + q; }
    TQ_HD int lamda_beta(int q) const { return 4 + (kZ + 2) * Q + q; }
    TQ_HD int trans_mean(int q, int zr, int z) const { return 4 + (kZ + 3) * Q + (q * kZ + zr) * kZ + z; }
    TQ_HD int trans_size(int q, int zr) const { return 4 + (kZ + 3) * Q + kZ * kZ * Q + q * kZ + zr; }
    TQ_HD int count() const { return 4 + (kZ + 3) * Q + (hmm ? (kZ * kZ + kZ) * Q : 0); }
    // flat layout of the global base variates / samples: gain, proximity, pi (Q,2), lamda (Q) [, trans (Q,2,2)]
    TQ_HD int n_gain() const { return 0; }
    TQ_HD int n_prox() const { return 1; }
    TQ_HD int n_pi(int q, int z) const { return 2 + q * kZ + z; }
    TQ_HD int n_lamda(int q) const { return 2 + kZ * Q + q; }
    TQ_HD int n_trans(int q, int zr, int z) const { return 2 + (kZ + 1) * Q + (q * kZ + zr) * kZ + z; }
    TQ_HD int n_count() const { return 2 + (kZ + 1) * Q + (hmm ? kZ * kZ * Q : 0); }
};

constexpr int kMaxGlobals = 4 + (kZ + 3 + kZ * kZ + kZ) * kMaxC;
constexpr int kMaxGlobalNoise = 2 + (kZ + 1 + kZ * kZ) * kMaxC;

TQ_HD double clamp_prob(double p, const ModelConst& mc) {
    return fmin(fmax(p, mc.eps), 1.0 - mc.eps);
}

// p(m_k = 1 | theta = 0) and p(m_k = 1 | theta = other spot) for K = 2 and their d/d lamda
// (truncated Poisson, distributions/util.py:67-130)
TQ_HD void probs_m_k2(double lam, double& p0, double& dp0, double& p1, double& dp1) {
    const double e = exp(-lam);
    p0 = 0.5 * (lam * e + 2.0 * (1.0 - e * (1.0 + lam)));
    dp0 = 0.5 * e * (1.0 + lam);
    p1 = 1.0 - e;
    dp1 = e;
}

// The global sites are independent of each other, so both directions are written per site and the
// kernels give each site its own lane (a serial single-thread version cost 80 us per step at C2
// scale, as much as a third of the likelihood kernel).  Site ids: 0 gain, 1 proximity, 2+q pi_q,
// 2+Q+q lamda_q.
// hmm: + one Dirichlet site per (channel, row of the transition matrix): 2 + 2Q + 2q + z'
TQ_HD int global_site_count(int Q, bool hmm = false) { return 2 + 2 * Q + (hmm ? kZ * Q : 0); }

// ---- forward: variates -> samples -> tables --------------------------------------------------------
// u: unconstrained global params; variate: base draws (replay) or filled here from `rng`;
// sample: gain, proximity, pi, lamda in the GlobalLayout::n_* order.  Writes only the entries of
// `variate`, `sample` and `gt` that belong to `site`.
TQ_HD void globals_pre_site(int site, const double* u, const GlobalLayout& gl, const ModelConst& mc, bool use_rng,
                            Philox* rng, double* variate, double* sample, GlobalTables<double>& gt) {
    static_assert(kK == 2, "probs_m closed form below is written for K = 2");
    const double hi = (mc.P + 1) / sqrt(12.0);
    if (site == 0) {
        // gain ~ Gamma(gain_loc * gain_beta, gain_beta)                                 cosmos.py:342-348
        const double loc = exp(u[gl.gain_loc()]), beta = exp(u[gl.gain_beta()]);
        if (use_rng) variate[gl.n_gain()] = fmax(sample_std_gamma<double>(*rng, loc * beta), mc.tiny);
        const double gain = fmax(variate[gl.n_gain()] / beta, mc.tiny);
        sample[gl.n_gain()] = gain;
        gt.gain = gain;
        gt.rate = 1.0 / gain;
        gt.log_rate = log(gt.rate);
        return;
    }
    if (site == 1) {
        // proximity ~ AffineBeta(loc, size, 0, (P+1)/sqrt(12))                          cosmos.py:360-368
        const Transformed<double> loc = t_interval<double>(u[gl.prox_loc()], 0.0, hi - mc.eps, mc);
        const Transformed<double> size = t_greater_than<double>(u[gl.prox_size()], 2.0);
        const AffBeta<double> d(loc.v, size.v, 0.0, hi);
        if (use_rng) {
            const double g1 = sample_std_gamma<double>(*rng, d.c1), g2 = sample_std_gamma<double>(*rng, d.c0);
            variate[gl.n_prox()] = beta01_from_gammas(g1, g2, mc);
        }
        const double prox = d.clamp(d.low + d.scale * variate[gl.n_prox()], mc);
        sample[gl.n_prox()] = prox;
        const double r = (mc.P + 1) / (2.0 * prox);
        gt.size1 = r * r - 1.0;                                                        // cosmos.py:185-191
        const double ln2 = 0.69314718055994530942, lP1 = log((double)(mc.P + 1));
        double lg1, ps1, lgh, psh;
        lgamma_digamma(gt.size1, lg1, ps1);
        lgamma_digamma(0.5 * gt.size1, lgh, psh);
        gt.cxy1 = 2.0 * (lg1 - 2.0 * lgh) - 4.0 * (0.5 * gt.size1 - 1.0) * ln2 - 2.0 * lP1;
        gt.dcxy1 = 2.0 * (ps1 - psh) - 2.0 * ln2;
        gt.lxy0 = -2.0 * lP1;
        return;
    }
    const double le = log(mc.eps), l1e = log(1.0 - mc.eps);
    if (site < 2 + gl.Q || site >= 2 + 2 * gl.Q) {
        // pi_q ~ Dirichlet(pi_mean * pi_size) (cosmos.py:349-352; hmm: init_q, hmm.py:279-284), or row z' of the hmm
        // transition matrix trans_q ~ Dirichlet(trans_mean * trans_size) (hmm.py:285-290)
        const bool is_trans = site >= 2 + 2 * gl.Q;
        const int idx = is_trans ? site - 2 - 2 * gl.Q : site - 2;
        const int q = is_trans ? idx / kZ : idx, zr = is_trans ? idx % kZ : 0;
        const double u0 = u[is_trans ? gl.trans_mean(q, zr, 0) : gl.pi_mean(q, 0)];
        const double u1 = u[is_trans ? gl.trans_mean(q, zr, 1) : gl.pi_mean(q, 1)];
        const int n0 = is_trans ? gl.n_trans(q, zr, 0) : gl.n_pi(q, 0), n1 = n0 + 1;
        const double mxu = fmax(u0, u1);
        const double e0 = exp(u0 - mxu), e1 = exp(u1 - mxu);
        const double size = exp(u[is_trans ? gl.trans_size(q, zr) : gl.pi_size(q)]);
        const double a0 = e0 / (e0 + e1) * size, a1 = e1 / (e0 + e1) * size;
        if (use_rng) {
            const double g0 = sample_std_gamma<double>(*rng, a0), g1 = sample_std_gamma<double>(*rng, a1);
            variate[n0] = fmin(fmax(g0 / (g0 + g1), mc.tiny), 1.0 - mc.eps);
            variate[n1] = fmin(fmax(g1 / (g0 + g1), mc.tiny), 1.0 - mc.eps);
        }
        const double p0 = variate[n0], p1 = variate[n1];
        sample[n0] = p0;
        sample[n1] = p1;
        ChannelTables<double>& ct = gt.ch[q];
        // Categorical(probs).logits = log(clamp(probs / sum)); off-target rows are [1, 0]          util.py:133-151
        if (is_trans) {
            ct.logptrans[0][zr][0] = l1e;
            ct.logptrans[0][zr][1] = le;
            ct.logptrans[1][zr][0] = log(clamp_prob(p0 / (p0 + p1), mc));
            ct.logptrans[1][zr][1] = log(clamp_prob(p1 / (p0 + p1), mc));
            return;
        }
        ct.logpz[0][0] = l1e;
        ct.logpz[0][1] = le;
        ct.logpz[1][0] = log(clamp_prob(p0 / (p0 + p1), mc));
        ct.logpz[1][1] = log(clamp_prob(p1 / (p0 + p1), mc));
        // p(theta | z): z = 0 -> e_0; z = 1 -> uniform over 1..K                          util.py:154-173
        for (int th = 0; th < kTheta; ++th) {
            ct.logptheta[0][th] = (th == 0) ? l1e : le;
            ct.logptheta[1][th] = (th == 0) ? le : log(clamp_prob(1.0 / kK, mc));
        }
        return;
    }
    {
        const int q = site - 2 - gl.Q;
        // lamda_q ~ Gamma(lamda_loc * lamda_beta, lamda_beta)                              cosmos.py:353-359
        const double lloc = exp(u[gl.lamda_loc(q)]), lbeta = exp(u[gl.lamda_beta(q)]);
        if (use_rng) variate[gl.n_lamda(q)] = fmax(sample_std_gamma<double>(*rng, lloc * lbeta), mc.tiny);
        const double lam = fmax(variate[gl.n_lamda(q)] / lbeta, mc.tiny);
        sample[gl.n_lamda(q)] = lam;
        ChannelTables<double>& ct = gt.ch[q];
        double pm0, d0, pm1, d1;
        probs_m_k2(lam, pm0, d0, pm1, d1);
        for (int th = 0; th < kTheta; ++th)
            for (int k = 0; k < kK; ++k) {
                const double p = clamp_prob(th == 0 ? pm0 : (th == k + 1 ? 1.0 : pm1), mc);
                ct.logpm[th][k][1] = log(p);
                ct.logpm[th][k][0] = log1p(-p);
            }
    }
}

// ---- reverse: accumulators -> loss and d loss / d unconstrained globals ------------------------------
// The data reach the reverse mode of a global site only through one or two LINEAR functionals of the
// accumulators (its "drive"); everything else -- densities, digammas, implicit reparameterisation
// gradients: the expensive part -- depends on the sample alone.  So the gradient is affine in the drive:
//     grad(drive) = grad(0) + drive[0] (grad(e0) - grad(0)) + drive[1] (grad(e1) - grad(0)),
// which lets the kernels evaluate grad(0), grad(e0), grad(e1) right after sampling, off the critical
// path (globals_prepare), and finish with a handful of FMAs once the accumulators exist (globals_finish).
//
// acc: [Q][NACC] sums over all units of all ranks (unscaled, masked); sN = Nt/nb, sF = F/fb.
// hacc (hmm only): [Q][NHACC] sums over (AOI, frame) of the chain's ELBO terms and of the expected initial-state /
// transition counts (cosmos_hmm.cuh), unscaled, masked, on-target AOIs only for the counts.
TQ_HD void globals_drive(int site, const GlobalLayout& gl, const ModelConst& mc, const double* sample, const double* acc,
                         const double* hacc, double sN, double sF, double (&drive)[2], double& elbo_data) {
    const double s = sN * sF;
    drive[0] = drive[1] = 0.0;
    elbo_data = 0.0;
    constexpr int kHaccElbo = 0, kHaccInit = 1, kHaccTrans = 1 + kZ, kNHacc = 1 + kZ + kZ * kZ;   // = cosmos_hmm.cuh HACC_*
    if (site == 0) {
        for (int q = 0; q < gl.Q; ++q) {
            const double* a = acc + q * NACC;
            elbo_data += s * a[ACC_ELBO_FRAME] + sN * a[ACC_ELBO_AOI];
            if (gl.hmm) elbo_data += s * hacc[q * kNHacc + kHaccElbo];
            drive[0] += s * a[ACC_RATE];
        }
    } else if (site == 1) {
        for (int q = 0; q < gl.Q; ++q) drive[0] += s * acc[q * NACC + ACC_SIZE1];
    } else if (site < 2 + gl.Q) {
        const int q = site - 2;
        for (int z = 0; z < kZ; ++z) drive[z] = gl.hmm ? s * hacc[q * kNHacc + kHaccInit + z] : s * acc[q * NACC + ACC_LOGPZ + z];
    } else if (site >= 2 + 2 * gl.Q) {
        const int idx = site - 2 - 2 * gl.Q, q = idx / kZ, zr = idx % kZ;
        for (int z = 0; z < kZ; ++z) drive[z] = s * hacc[q * kNHacc + kHaccTrans + zr * kZ + z];
    } else {
        const int q = site - 2 - gl.Q;
        const double* a = acc + q * NACC;
        double pm0, d0, pm1, d1;
        probs_m_k2(sample[gl.n_lamda(q)], pm0, d0, pm1, d1);
        for (int th = 0; th < kTheta; ++th)
            for (int k = 0; k < kK; ++k) {
                if (th == k + 1) continue;  // certain spot: probability 1, no lamda dependence
                const double p = th == 0 ? pm0 : pm1, dp = th == 0 ? d0 : d1;
                if (p < mc.eps || p > 1.0 - mc.eps) continue;
                const double* t = a + ACC_LOGPM + (th * kK + k) * 2;
                drive[0] += s * (t[1] / p - t[0] / (1.0 - p)) * dp;
            }
    }
}

// Returns the site's own part of the ELBO (log prior - log q of its sample); writes grad[i] = d loss / d u[i]
// for the parameters of `site` only.
TQ_HD double globals_post_site_driven(int site, const double* u, const GlobalLayout& gl, const ModelConst& mc,
                                      const double* sample, const double (&drive)[2], double* grad) {
    const double hi = (mc.P + 1) / sqrt(12.0);
    double elbo = 0.0;
    if (site == 0) {
        const double rate_grad = drive[0];
        const double loc = exp(u[gl.gain_loc()]), beta = exp(u[gl.gain_beta()]);
        const double conc = loc * beta;
        const double g = sample[gl.n_gain()];
        const double sd = mc.gain_std;
        const double lp = -g * g / (2 * sd * sd) - log(sd) + 0.5 * log(2.0 / 3.14159265358979323846);  // HalfNormal :170
        const GammaSite<double> qg(g, conc, beta);
        elbo += lp - qg.lp;
        const double G = rate_grad * (-1.0 / (g * g)) - g / (sd * sd) - qg.d_v;
        const double dv_dconc = std_gamma_grad<double>(conc, g * beta) / beta;
        const double g_conc = G * dv_dconc - qg.d_conc;
        const double g_rate = G * (-g / beta) - qg.d_rate;
        grad[gl.gain_loc()] = -(g_conc * beta * loc);
        grad[gl.gain_beta()] = -((g_conc * loc + g_rate) * beta);
        return elbo;
    }
    if (site == 1) {
        const double size1_grad = drive[0];
        const Transformed<double> loc = t_interval<double>(u[gl.prox_loc()], 0.0, hi - mc.eps, mc);
        const Transformed<double> size = t_greater_than<double>(u[gl.prox_size()], 2.0);
        const AffBeta<double> d(loc.v, size.v, 0.0, hi);
        const double v = sample[gl.n_prox()];
        const BetaSite<double> qs(v, d);
        elbo += (log(mc.proximity_rate) - mc.proximity_rate * v) - qs.lp;  // Exponential prior :182-184
        const double half = 0.5 * (mc.P + 1);
        const double dsize1 = -2.0 * half * half / (v * v * v);
        const double G = size1_grad * dsize1 - mc.proximity_rate - qs.d_v;
        double bg1, bg0;
        beta_grad_pair<double>(qs.x01, d.c1, d.c0, bg1, bg0);
        const double dv_dc1 = d.scale * (1.0 - qs.x01) * bg1;
        const double dv_dc0 = -d.scale * qs.x01 * bg0;
        const double g_c1 = G * dv_dc1 - qs.d_c1, g_c0 = G * dv_dc0 - qs.d_c0;
        grad[gl.prox_loc()] = -((g_c1 - g_c0) * size.v / d.scale * loc.d);
        grad[gl.prox_size()] = -((g_c1 * (loc.v - d.low) / d.scale + g_c0 * (d.low + d.scale - loc.v) / d.scale) * size.d);
        return elbo;
    }
    if (site < 2 + gl.Q || site >= 2 + 2 * gl.Q) {
        // Dirichlet over two states: pi_q / init_q, or row z' of trans_q (same prior Dirichlet(1/2, 1/2), hmm.py:87-98)
        const bool is_trans = site >= 2 + 2 * gl.Q;
        const int idx = is_trans ? site - 2 - 2 * gl.Q : site - 2;
        const int q = is_trans ? idx / kZ : idx, zr = is_trans ? idx % kZ : 0;
        const int i_m0 = is_trans ? gl.trans_mean(q, zr, 0) : gl.pi_mean(q, 0), i_m1 = i_m0 + 1;
        const int i_sz = is_trans ? gl.trans_size(q, zr) : gl.pi_size(q);
        const int n0 = is_trans ? gl.n_trans(q, zr, 0) : gl.n_pi(q, 0);
        const double u0 = u[i_m0], u1 = u[i_m1];
        const double mxu = fmax(u0, u1);
        const double e0 = exp(u0 - mxu), e1 = exp(u1 - mxu);
        const double mean[2] = {e0 / (e0 + e1), e1 / (e0 + e1)};
        const double size = exp(u[i_sz]);
        const double conc[2] = {mean[0] * size, mean[1] * size};
        const double x[2] = {sample[n0], sample[n0 + 1]};
        const double tot = conc[0] + conc[1], sx = x[0] + x[1];
        const double prior_c = 1.0 / kZ;  // Dirichlet(1/(S+1))                               cosmos.py:171-174
        double lp = lgamma_pos(prior_c * kZ), lq = lgamma_pos(tot);
        double gx[2];
        for (int z = 0; z < kZ; ++z) {
            lp += (prior_c - 1.0) * log(x[z]) - lgamma_pos(prior_c);
            lq += (conc[z] - 1.0) * log(x[z]) - lgamma_pos(conc[z]);
            // table path: logpz[1][z] = log(clamp(x_z / sum x))
            const double pz = x[z] / sx;
            const bool inside = pz >= mc.eps && pz <= 1.0 - mc.eps;
            gx[z] = (prior_c - 1.0) / x[z] - (conc[z] - 1.0) / x[z];
            if (inside) gx[z] += drive[z] / x[z];
        }
        for (int z = 0; z < kZ; ++z) {
            const double pz = x[z] / sx;
            if (pz >= mc.eps && pz <= 1.0 - mc.eps) {
                gx[0] -= drive[z] / sx;
                gx[1] -= drive[z] / sx;
            }
        }
        elbo += lp - lq;
        // Dirichlet reparameterisation (torch dirichlet.py _Dirichlet_backward) + direct -dlogq/dconc
        const double dot = x[0] * gx[0] + x[1] * gx[1];
        const double pt = digamma<double>(tot);
        double g_conc[2];
        for (int z = 0; z < kZ; ++z)
            g_conc[z] = beta_grad<double>(x[z], conc[z], tot) * (gx[z] - dot) - (log(x[z]) + pt - digamma<double>(conc[z]));
        // conc = softmax(u_mean) * exp(u_size)
        const double gm[2] = {g_conc[0] * size, g_conc[1] * size};
        const double gmdot = mean[0] * gm[0] + mean[1] * gm[1];
        grad[i_m0] = -(mean[0] * (gm[0] - gmdot));
        grad[i_m1] = -(mean[1] * (gm[1] - gmdot));
        grad[i_sz] = -((g_conc[0] * mean[0] + g_conc[1] * mean[1]) * size);
        return elbo;
    }
    {
        const int q = site - 2 - gl.Q;
        const double loc = exp(u[gl.lamda_loc(q)]), beta = exp(u[gl.lamda_beta(q)]);
        const double conc = loc * beta;
        const double lam = sample[gl.n_lamda(q)];
        const GammaSite<double> ql(lam, conc, beta);
        elbo += (log(mc.lamda_rate) - mc.lamda_rate * lam) - ql.lp;  // Exponential prior :176-181
        const double G = -mc.lamda_rate - ql.d_v + drive[0];
        const double dv_dconc = std_gamma_grad<double>(conc, lam * beta) / beta;
        const double g_conc = G * dv_dconc - ql.d_conc;
        const double g_rate = G * (-lam / beta) - ql.d_rate;
        grad[gl.lamda_loc(q)] = -(g_conc * beta * loc);
        grad[gl.lamda_beta(q)] = -((g_conc * loc + g_rate) * beta);
        return elbo;
    }
}

// one-shot form: this site's part of the ELBO (site 0 also carries the data terms) and its gradients
TQ_HD double globals_post_site(int site, const double* u, const GlobalLayout& gl, const ModelConst& mc,
                               const double* sample, const double* acc, double sN, double sF, double* grad,
                               const double* hacc = nullptr) {
    double drive[2], elbo_data;
    globals_drive(site, gl, mc, sample, acc, hacc, sN, sF, drive, elbo_data);
    return elbo_data + globals_post_site_driven(site, u, gl, mc, sample, drive, grad);
}

// global site owning parameter i of the flat GlobalLayout
TQ_HD int global_param_site(int i, int Q) {
    if (i < 2) return 0;
    if (i < 4) return 1;
    if (i < 4 + kZ * Q) return 2 + (i - 4) / kZ;           // pi_mean (q, z)
    if (i < 4 + (kZ + 1) * Q) return 2 + (i - 4 - kZ * Q);  // pi_size q
    if (i < 4 + (kZ + 2) * Q) return 2 + Q + (i - 4 - (kZ + 1) * Q);
    if (i < 4 + (kZ + 3) * Q) return 2 + Q + (i - 4 - (kZ + 2) * Q);
    const int j = i - 4 - (kZ + 3) * Q;                     // hmm: trans_mean (q, z', z) then trans_size (q, z')
    if (j < kZ * kZ * Q) return 2 + 2 * Q + j / kZ;
    return 2 + 2 * Q + (j - kZ * kZ * Q);
}

// prepared reverse mode: for every site, grad at drive = 0, e0, e1 (full GlobalLayout vectors; only the
// site's own entries are meaningful) and the site's own ELBO part
constexpr int kMaxGlobalSites = 2 + (2 + kZ) * kMaxC;
struct GlobalPrep {
    double grad[kMaxGlobalSites][3][kMaxGlobals];
    double elbo[kMaxGlobalSites];
};

// serial forms (host check, tests)
TQ_HD void globals_pre(const double* u, const GlobalLayout& gl, const ModelConst& mc, bool use_rng, Philox* rng,
                       double* variate, double* sample, GlobalTables<double>& gt) {
    for (int site = 0; site < global_site_count(gl.Q, gl.hmm); ++site)
        globals_pre_site(site, u, gl, mc, use_rng, rng, variate, sample, gt);
}
TQ_HD double globals_post(const double* u, const GlobalLayout& gl, const ModelConst& mc, const double* sample,
                          const double* acc, double sN, double sF, double* grad) {
    double elbo = 0.0;
    for (int site = 0; site < global_site_count(gl.Q, gl.hmm); ++site)
        elbo += globals_post_site(site, u, gl, mc, sample, acc, sN, sF, grad);
    return elbo;
}

}  // namespace tq
