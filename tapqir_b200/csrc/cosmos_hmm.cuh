// Per-unit part of the *hmm* variant of cosmos (reference: tapqir/models/hmm.py; BASELINE config 5).
//
// Differences from cosmos: the guide enumerates a Markov chain z_f with AOI-local transition
// tables z_trans (hmm.py:355-364) and m_k conditional on z_f (m_probs has a leading (1+S) axis, :368-377); the model
// enumerates theta (:178-186); there is no frame subsampling (:127-131).  Pyro's TraceEnum_ELBO then computes, per
// (AOI, channel), with a_f the forward marginals of the guide's chain,
//
//   sum_f sum_{z',z} a_{f-1}(z') q_f(z|z') [log p_f(z|z') - log q_f(z|z')]
//   + sum_f [ log p(b_f) - log q(b_f) + sum_z a_f(z) V_f(z) ],
//   V_f(z) = sum_m q_f(m|z) ( T_z(m) + L(m) + sum_k m_k (masked h,w,x,y terms) - log q_f(m|z) ),
//   T_z(m) = log sum_theta p(theta|z) prod_k p(m_k|theta) (p(x_k|theta) p(y_k|theta))^{m_k}.
//
// The image likelihood L(m) is the SAME 4-configuration kernel as cosmos: its upstream weight
// W_f(m) = sum_z a_f(z) q_f(m|z) depends on guide parameters only, so forward and reverse mode are still one sweep.
// This file: the per-unit emission terms (unit_post_hmm) and the chain's forward / backward recursions.
#pragma once
#include "cosmos_local.cuh"

namespace tq {

// Outputs of unit_post_hmm on top of UnitGrads (whose g[LP_M_PROBS + k] are unused here)
template <typename F> struct HmmUnitOut {
    F gmp[kZ][kK];   // d / d unconstrained m_probs[z][k] (unit-level, unscaled, before the mask)
    F v[kZ];         // V_f(z) - L(m = 0): emission value of state z, centred for fp32 (only V(1) - V(0) propagates)
};

// W(m) = sum_z a(z) prod_k q(m_k | z): the likelihood kernel's weights
template <typename F>
TQ_HD void hmm_presence_weights(const F (&u_mp)[kZ][kK], const F (&a)[kZ], const ModelConst& mc, F (&qm)[kM]) {
#pragma unroll
    for (int m = 0; m < kM; ++m) qm[m] = F(0);
#pragma unroll
    for (int z = 0; z < kZ; ++z) {
        F q1[kK], q0[kK], w[kM];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const SpotPresence<F> sp(u_mp[z][k], mc);
            q1[k] = sp.q1; q0[k] = sp.q0;
        }
        presence_weights<F>(q1, q0, w);
#pragma unroll
        for (int m = 0; m < kM; ++m) qm[m] += a[z] * w[m];
    }
}

// Same arguments as unit_post, with the guide's per-state m_probs and the forward marginals a(z) of this frame.
template <typename F>
TQ_HD void unit_post_hmm(const F (&rec)[NREC], const F (&sample)[NSAMP], const F (&L)[kM], const F (&gs)[NSAMP],
                         F g_rate, const F (&u_mp)[kZ][kK], const F (&a)[kZ], F u_bm, F u_bs, const ModelConst& mc,
                         const GlobalTables<F>& gt, int c, bool first_frame, UnitGrads<F>& out, HmmUnitOut<F>& ho) {
    using R = Real<F>;
    const ChannelTables<F>& ct = gt.ch[c];
    const F half = F(mc.P + 1) / F(2);
#pragma unroll
    for (int i = 0; i < NLOCAL; ++i) out.g[i] = F(0);
#pragma unroll
    for (int i = 0; i < NACC; ++i) out.acc[i] = F(0);
    auto R_ = [&](int s, int j) -> F { return rec[s * NSO + j]; };

    // ---- background (as cosmos) ----------------------------------------------------------------------------------------
    const F* ex = rec + NSAMP * NSO;
    F elbo = ex[EX_LP] - R_(S_B, SO_LQ);
    {
        const F G = ex[EX_DP] - R_(S_B, SO_DQ) + gs[S_B];
        out.g[LP_B_LOC] = G * R_(S_B, SO_A0) - R_(S_B, SO_B0);
        out.g[LP_B_BETA] = G * R_(S_B, SO_A1) - R_(S_B, SO_B1);
        out.g[LP_BM] = ex[EX_GBM];
        out.g[LP_BS] = ex[EX_GBS];
    }

    // ---- per-spot terms that do not depend on (z, theta) (as cosmos) -------------------------------------------------------
    F spot_term[kK], lxy1[kK], dlxy1_dx[kK], dlxy1_dy[kK], dlxy1_dsize[kK];
    const F cs1 = gt.size1 * F(0.5) - F(1);
    const F hs = F(mc.height_std);
    const F lp_w = -R::log(F(mc.width_max) - F(mc.width_min));
    const F c_hn = -R::log(hs) + F(0.5) * R::log(F(2) / F(3.14159265358979323846));
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F h = sample[S_H + k], x = sample[S_X + k], y = sample[S_Y + k];
        const F lp_h = -h * h / (F(2) * hs * hs) + c_hn;
        spot_term[k] = lp_h + lp_w - R_(S_H + k, SO_LQ) - R_(S_W + k, SO_LQ) - R_(S_X + k, SO_LQ) - R_(S_Y + k, SO_LQ);
        const F tx = x / half, ty = y / half;
        const F ox = F(1) - tx * tx, oy = F(1) - ty * ty;
        const F lsum = R::log1p(-tx * tx) + R::log1p(-ty * ty);
        lxy1[k] = cs1 * lsum + gt.cxy1;
        dlxy1_dx[k] = cs1 * (-F(2) * tx / ox) / half;
        dlxy1_dy[k] = cs1 * (-F(2) * ty / oy) / half;
        dlxy1_dsize[k] = F(0.5) * lsum + gt.dcxy1;
    }

    // ---- enumerated part: for each guide state z, theta summed out, weighted by a(z) q(m | z) -----------------------------
    F wk_x[kK], wq[kK];   // posterior weight of "spot k is the target-specific one" / marginal presence of spot k
#pragma unroll
    for (int k = 0; k < kK; ++k) wk_x[k] = wq[k] = F(0);
    F sum_av = F(0);
#pragma unroll
    for (int z = 0; z < kZ; ++z) {
        F q1[kK], q0[kK], lq1[kK], lq0[kK], dq1[kK];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const SpotPresence<F> sp(u_mp[z][k], mc);
            q1[k] = sp.q1; q0[k] = sp.q0; lq1[k] = sp.lq1; lq0[k] = sp.lq0; dq1[k] = sp.dq1;
            wq[k] += a[z] * sp.q1;
            ho.gmp[z][k] = F(0);
        }
        F V = F(0);
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            F lj[kTheta];
            F mx = -R::inf();
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                F v = ct.logptheta[z][th];
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const int mk = (m >> k) & 1;
                    v += ct.logpm[th][k][mk];
                    if (mk) v += (th == k + 1) ? lxy1[k] : gt.lxy0;
                }
                lj[th] = v;
                mx = R::max(mx, v);
            }
            F se = F(0);
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                lj[th] = R::exp_fast(lj[th] - mx);
                se += lj[th];
            }
            const F T = mx + R::log_fast(se);
            F q = F(1), lq = F(0), Cm = T + (L[m] - L[0]);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                const int mk = (m >> k) & 1;
                q *= mk ? q1[k] : q0[k];
                lq += mk ? lq1[k] : lq0[k];
                if (mk) Cm += spot_term[k];
            }
            Cm -= lq;
            V += q * Cm;
            const F w = a[z] * q;
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                const int mk = (m >> k) & 1;
                ho.gmp[z][k] += w * (mk ? q0[k] : -q1[k]) * (Cm - F(1));
            }
            const F wi = w / se;
#pragma unroll
            for (int th = 0; th < kTheta; ++th) {
                const F r = wi * lj[th];
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const int mk = (m >> k) & 1;
                    out.acc[ACC_LOGPM + (th * kK + k) * 2 + mk] += r;
                    if (mk && th == k + 1) wk_x[k] += r;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) if (!(dq1[k] > F(0))) ho.gmp[z][k] = F(0);
        ho.v[z] = V;
        sum_av += a[z] * V;
    }
    elbo += sum_av + L[0];
    out.acc[ACC_ELBO_FRAME] = elbo;
    out.acc[ACC_RATE] = g_rate;
#pragma unroll
    for (int k = 0; k < kK; ++k) out.acc[ACC_SIZE1] += wk_x[k] * dlxy1_dsize[k];

    // ---- parameter gradients through the site maps (as cosmos, with the marginal presence wq) -----------------------------
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const F qk = wq[k];
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
    if (first_frame) {
        const F bm = R::exp(u_bm), bs = R::exp(u_bs);
        const F s1 = F(mc.bg_mean_std), s2 = F(mc.bg_std_std);
        const F c0 = F(0.5) * R::log(F(2) / F(3.14159265358979323846));
        out.acc[ACC_ELBO_AOI] = (-bm * bm / (F(2) * s1 * s1) - R::log(s1) + c0) + (-bs * bs / (F(2) * s2 * s2) - R::log(s2) + c0);
    }
}

// ---- posterior of theta for one guide draw, given the chain state (hmm.py:541-625, "next" row N1 for hmm) --------------
// p(theta | z, m, x, y) from the model's theta / m / x / y log-probs, averaged over m with the guide weights q(m | z):
// pth[k] = that probability at theta = k + 1.  (Terms that do not depend on theta -- the chain's p(z_f | z_{f-1}) -- cancel
// in the normalisation over theta, which is why only z itself is needed.)
template <typename F>
TQ_HD void unit_theta_given_z(const F (&x_)[kK], const F (&y_)[kK], const F (&u_mp)[kK], int z, const ModelConst& mc,
                              const GlobalTables<F>& gt, int c, F (&pth)[kK]) {
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
        pth[k] = F(0);
    }
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        F lj[kTheta];
        F mx = -R::inf();
#pragma unroll
        for (int th = 0; th < kTheta; ++th) {
            F v = ct.logptheta[z][th];
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                const int mk = (m >> k) & 1;
                v += ct.logpm[th][k][mk];
                if (mk) v += (th == k + 1) ? lxy1[k] : gt.lxy0;
            }
            lj[th] = v;
            mx = R::max(mx, v);
        }
        F se = F(0);
#pragma unroll
        for (int th = 0; th < kTheta; ++th) {
            lj[th] = R::exp(lj[th] - mx);
            se += lj[th];
        }
        F q = F(1);
#pragma unroll
        for (int k = 0; k < kK; ++k) q *= ((m >> k) & 1) ? q1[k] : q0[k];
        const F qi = q / se;
#pragma unroll
        for (int k = 0; k < kK; ++k) pth[k] += qi * lj[k + 1];
    }
}

// ---- the guide's chain ---------------------------------------------------------------------------------------------------
// Row z' of one frame's transition table: q(z | z') = clamp(softmax(u[z'][:])) and its log (Categorical(probs).logits,
// hmm.py:355-364 with torch's probs clamp).
struct ChainRow { double q[kZ], lq[kZ]; };
TQ_HD ChainRow chain_row(double u0, double u1, const ModelConst& mc) {
    ChainRow r;
    const double mx = fmax(u0, u1);
    const double e0 = exp(u0 - mx), e1 = exp(u1 - mx), inv = 1.0 / (e0 + e1);
    r.q[0] = fmin(fmax(e0 * inv, mc.eps), 1.0 - mc.eps);
    r.q[1] = fmin(fmax(e1 * inv, mc.eps), 1.0 - mc.eps);
    r.lq[0] = log(r.q[0]);
    r.lq[1] = log(r.q[1]);
    return r;
}

// per-(AOI, channel) sums the chain contributes to the globals' reverse mode and to the ELBO
enum { HACC_ELBO = 0, HACC_INIT = 1, HACC_TRANS = HACC_INIT + kZ, NHACC = HACC_TRANS + kZ * kZ };

}  // namespace tq
