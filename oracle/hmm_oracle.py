"""
CPU fp64 restatement of one SVI step of the reference's *hmm* model (tapqir/models/hmm.py) --
TEST INFRASTRUCTURE ONLY (imported by tests/, never by tapqir_b200/).

Parity status: the continuous sites, the prior tables and the image likelihood are the functions of
oracle/cosmos_oracle.py (pinned against the reference's own distributions code).  The ELBO assembly -- what Pyro's
``TraceEnum_ELBO`` computes for a guide that enumerates the Markov chain ``z_f`` and the spot presences ``m_k``
(hmm.py:355-377) and a model that enumerates ``theta`` (hmm.py:178-186) -- is the closed form of SURVEY.md App. B.2
below.  It is PINNED against the reference's own ``models/hmm.py`` (init_parameters, guide, model) run verbatim in its
sequential form (``vectorized=False``: pyro.markov + TraceEnum_ELBO) by tests/golden/make_golden_step.py ->
tests/golden/ref_step_hmm.pt, where the expectation over the enumerated chain is taken by brute force over every
enumerated dimension: losses 1e-15, gradients 7e-10, Adam trajectory 1e-11 (tests/test_hmm_cpu.py).  Pyro itself
(absent, not installable; funsor for the vectorised form likewise) is restated by tests/golden/minipyro.py -- "parity
unpinned" for Pyro's own machinery, as for cosmos.

    a_f(z)      = sum_z' a_{f-1}(z') q_f(z | z'),  a_{-1} = e_0          (row 0 of z_trans is used at f = 0)
    ELBO_nc     = sum_f sum_{z',z} a_{f-1}(z') q_f(z|z') [log p_f(z|z') - log q_f(z|z')]
                + sum_f [ log p(b_f) - log q(b_f) + sum_z a_f(z) sum_m q_f(m|z) ( T_z(m) + L(m)
                          + sum_k m_k (masked h, w, x, y terms) - log q_f(m|z) ) ]
    T_z(m)      = log sum_theta p(theta | z) prod_k p(m_k | theta) (p(x_k | theta) p(y_k | theta))^{m_k}
"""

import math
from typing import Dict

import torch
import torch.distributions as D
from torch.distributions import constraints, transform_to

from oracle import cosmos_oracle as O

DEFAULT_PRIORS = O.DEFAULT_PRIORS


def param_constraints(P, dtype):
    """hmm.py:433-467 on top of cosmos.py:471-598 (``pi_*`` replaced by ``init_*`` / ``trans_*``)."""
    cons = dict(O.param_constraints(P, dtype))
    del cons["pi_mean"], cons["pi_size"]
    cons.update(init_mean=constraints.simplex, init_size=constraints.positive, trans_mean=constraints.simplex,
                trans_size=constraints.positive, z_trans=constraints.simplex)
    return cons


GLOBAL_PARAMS = ["init_mean", "init_size", "trans_mean", "trans_size", "proximity_loc", "proximity_size", "lamda_loc",
                 "lamda_beta", "gain_loc", "gain_beta"]


def init_constrained(data: O.OracleData, K=2, S=1):
    """hmm.py:433-467 + cosmos.py ``_init_parameters``."""
    dt = data.dtype
    Q, Nt, F, C = data.C, data.Nt, data.F, data.C
    out = O.init_constrained(data, K, S)
    del out["pi_mean"], out["pi_size"]
    out["init_mean"] = torch.ones(Q, S + 1, dtype=dt)
    out["init_size"] = torch.full((Q, 1), 2.0, dtype=dt)
    out["trans_mean"] = torch.ones(Q, S + 1, S + 1, dtype=dt)
    out["trans_size"] = torch.full((Q, S + 1, 1), 2.0, dtype=dt)
    out["z_trans"] = torch.ones(Nt, F, C, 1 + S, 1 + S, dtype=dt)
    out["m_probs"] = torch.full((1 + S, K, Nt, F, C), 0.5, dtype=dt)
    return out


def to_unconstrained(constrained: Dict[str, torch.Tensor], P, dtype):
    cons = param_constraints(P, dtype)
    return {k: transform_to(cons[k]).inv(v).detach().clone() for k, v in constrained.items()}


def to_constrained(unconstrained: Dict[str, torch.Tensor], P, dtype):
    cons = param_constraints(P, dtype)
    return {k: transform_to(cons[k])(v) for k, v in unconstrained.items()}


def _gather_local(p, ndx):
    """Vindex gathers of hmm.py:337-417 over all frames (no frame subsampling, hmm.py:127-131)."""
    out = {}
    for name in ["h_loc", "h_beta", "w_mean", "w_size", "x_mean", "y_mean", "size"]:
        out[name] = p[name][:, ndx]            # (K,nb,F,C)
    out["m_probs"] = p["m_probs"][:, :, ndx]   # (1+S,K,nb,F,C)
    for name in ["b_loc", "b_beta", "background_mean_loc", "background_std_loc", "z_trans"]:
        out[name] = p[name][ndx]
    return out


def guide_dists(p, loc, P, priors):
    half = (P + 1) / 2
    return {
        "gain": (p["gain_loc"] * p["gain_beta"], p["gain_beta"]),                       # hmm.py:272-278
        "init": p["init_mean"] * p["init_size"],                                        # :279-284
        "trans": p["trans_mean"] * p["trans_size"],                                     # :285-290
        "lamda": (p["lamda_loc"] * p["lamda_beta"], p["lamda_beta"]),                   # :291-297
        "proximity": O.AffineBeta(p["proximity_loc"], p["proximity_size"], 0, (P + 1) / math.sqrt(12)),  # :298-306
        "background": (loc["b_loc"] * loc["b_beta"], loc["b_beta"]),                    # :345-352
        "height": (loc["h_loc"] * loc["h_beta"], loc["h_beta"]),                        # :380-387
        "width": O.AffineBeta(loc["w_mean"], loc["w_size"], priors["width_min"], priors["width_max"]),  # :388-396
        "x": O.AffineBeta(loc["x_mean"], loc["size"], -half, half),                     # :397-405
        "y": O.AffineBeta(loc["y_mean"], loc["size"], -half, half),                     # :406-414
    }


@torch.no_grad()
def draw_noise(unconstrained, data: O.OracleData, ndx, generator=None, priors=DEFAULT_PRIORS):
    """Base variates of one guide execution (same ATen samplers as ``rsample``)."""
    p = to_constrained(unconstrained, data.P, data.dtype)
    loc = _gather_local(p, ndx)
    g = guide_dists(p, loc, data.P, priors)
    sg = lambda c: torch._standard_gamma(c.contiguous(), generator=generator)

    def beta01(ab):
        conc = torch.stack(torch.broadcast_tensors(ab.concentration1, ab.concentration0), -1).contiguous()
        return torch._sample_dirichlet(conc, generator=generator)[..., 0]

    noise = {"gain": sg(g["gain"][0]),
             "init": torch._sample_dirichlet(g["init"].contiguous(), generator=generator),
             "trans": torch._sample_dirichlet(g["trans"].contiguous(), generator=generator),
             "lamda": sg(g["lamda"][0]), "proximity": beta01(g["proximity"]), "background": sg(g["background"][0]),
             "height": sg(g["height"][0]), "width": beta01(g["width"]), "x": beta01(g["x"]), "y": beta01(g["y"])}
    return noise


def elbo(unconstrained, data: O.OracleData, ndx, noise, priors=DEFAULT_PRIORS, K=2, S=1, return_parts=False, plate_n=None):
    """ELBO of one guide + model execution over AOIs ``ndx`` and ALL frames, differentiable w.r.t. ``unconstrained``.
    ``plate_n``: size of the AOI plate when ``data`` holds only the gathered AOIs of a larger dataset."""
    assert S == 1
    dt, P = data.dtype, data.P
    Q = C = data.C
    nb, F = len(ndx), data.F
    sN = (plate_n if plate_n is not None else data.Nt) / nb
    half = (P + 1) / 2
    p = to_constrained(unconstrained, P, dt)
    loc = _gather_local(p, ndx)
    g = guide_dists(p, loc, P, priors)
    t = lambda v: torch.as_tensor(v, dtype=dt)
    fdx = torch.arange(F)

    # ---- global sites (hmm.py:86-118 model, :272-306 guide) -----------------------------------------------------
    gain = O.rsample_gamma(*g["gain"], noise["gain"])
    init = O._ReplayDirichlet.apply(g["init"].contiguous(), noise["init"])
    trans = O._ReplayDirichlet.apply(g["trans"].contiguous(), noise["trans"])
    lamda = O.rsample_gamma(*g["lamda"], noise["lamda"])
    proximity = O.rsample_affine_beta(g["proximity"], noise["proximity"])
    e_global = (
        D.HalfNormal(t(priors["gain_std"])).log_prob(gain) - D.Gamma(*g["gain"]).log_prob(gain)
        + (D.Dirichlet(torch.ones(Q, S + 1, dtype=dt) / (S + 1)).log_prob(init) - D.Dirichlet(g["init"]).log_prob(init)).sum()
        + (D.Dirichlet(torch.ones(Q, S + 1, S + 1, dtype=dt) / (S + 1)).log_prob(trans)
           - D.Dirichlet(g["trans"]).log_prob(trans)).sum()
        + (D.Exponential(torch.full((Q,), priors["lamda_rate"], dtype=dt)).log_prob(lamda)
           - D.Gamma(*g["lamda"]).log_prob(lamda)).sum()
        + D.Exponential(t(priors["proximity_rate"])).log_prob(proximity) - g["proximity"].log_prob(proximity)
    )
    size = torch.stack([torch.full_like(proximity, 2.0), ((P + 1) / (2 * proximity)) ** 2 - 1], -1)

    # ---- AOI-level sites --------------------------------------------------------------------------------------------
    mask = data.mask[ndx].to(dt)[:, None, None]
    bm, bs = loc["background_mean_loc"], loc["background_std_loc"]
    e_aoi = (mask * (D.HalfNormal(t(priors["background_mean_std"])).log_prob(bm)
                     + D.HalfNormal(t(priors["background_std_std"])).log_prob(bs))).sum()

    # ---- frame-level continuous sites (same families as cosmos) --------------------------------------------------------
    background = O.rsample_gamma(*g["background"], noise["background"])
    e_b = D.Gamma((bm / bs) ** 2, bm / bs**2).log_prob(background) - D.Gamma(*g["background"]).log_prob(background)
    height = O.rsample_gamma(*g["height"], noise["height"])
    width = O.rsample_affine_beta(g["width"], noise["width"])
    x = O.rsample_affine_beta(g["x"], noise["x"])
    y = O.rsample_affine_beta(g["y"], noise["y"])
    spot_terms = (
        D.HalfNormal(t(priors["height_std"])).log_prob(height)
        + O.AffineBeta(t(1.5), t(2.0), priors["width_min"], priors["width_max"]).log_prob(width)
        - D.Gamma(*g["height"]).log_prob(height) - g["width"].log_prob(width) - g["x"].log_prob(x) - g["y"].log_prob(y)
    )  # (K,nb,F,C)

    # ---- guide: Markov chain over z, m_k | z (hmm.py:355-377) -----------------------------------------------------
    mcfg = O.m_configs(K, dt)
    M = mcfg.shape[0]
    logq_z = D.Categorical(probs=loc["z_trans"], validate_args=False).logits        # (nb,F,C,z',z)
    q_z = logq_z.exp()
    mp = loc["m_probs"]                                                                # (1+S,K,nb,F,C)
    logq_mk = D.Categorical(probs=torch.stack([1 - mp, mp], -1), validate_args=False).logits   # (1+S,K,nb,F,C,2)
    logq_m = sum(logq_mk[:, k][..., mcfg[:, k].long()] for k in range(K))            # (1+S,nb,F,C,M)
    q_m = logq_m.exp()

    # ---- model tables (hmm.py:164-197) ----------------------------------------------------------------------------------
    ont = data.is_ontarget[ndx].long()
    logp_init = D.Categorical(probs=O.expand_offtarget(init)[:, :, ont].permute(2, 0, 1), validate_args=False).logits   # (nb,C,z)
    logp_trans = D.Categorical(probs=O.expand_offtarget(trans)[:, :, :, ont].permute(3, 0, 1, 2),
                               validate_args=False).logits                                                              # (nb,C,z',z)
    logp_theta = D.Categorical(probs=O.probs_theta(K, dt), validate_args=False).logits
    pm = O.probs_m(lamda, K)                                                           # (Q,1+K,K)
    logp_mk = D.Categorical(probs=torch.stack([1 - pm, pm], -1), validate_args=False).logits   # (Q,1+K,K,2)
    xy_prior = [O.AffineBeta(t(0.0), size[s], -half, half) for s in range(2)]
    logp_xy = torch.stack([d.log_prob(x) + d.log_prob(y) for d in xy_prior])         # (2,K,nb,F,C)

    # T_z(m): theta summed out for a fixed z
    Tz = []
    for z in range(S + 1):
        terms = []
        for th in range(K + 1):
            term = logp_theta[min(z, 1), th] + torch.zeros(nb, F, C, M, dtype=dt)
            for k in range(K):
                mk = mcfg[:, k]
                spec = int(th == k + 1)
                term = term + logp_mk[:, th, k][:, mk.long()][None, None] + mk * logp_xy[spec, k][..., None]
            terms.append(term)
        Tz.append(torch.logsumexp(torch.stack(terms), 0))
    Tz = torch.stack(Tz)                                                                # (1+S,nb,F,C,M)

    stk = lambda v: v.permute(1, 2, 3, 0)
    target = data.xy[ndx[:, None], fdx[None, :]]
    obs = data.images[ndx[:, None], fdx[None, :]]
    L = O.ksmogn_log_prob(stk(height), stk(width), stk(x), stk(y), target, background, gain, data.offset_samples,
                          data.offset_logits, P, obs, m=mcfg[:, None, None, None, :])   # (M,nb,F,C)
    L = L.permute(1, 2, 3, 0)                                                          # (nb,F,C,M)
    masked = sum(mcfg[:, k] * spot_terms[k][..., None] for k in range(K))              # (nb,F,C,M)
    emission = (q_m * (Tz + L + masked - logq_m)).sum(-1)                              # (1+S,nb,F,C): E_z per unit

    # ---- forward marginals and the chain's ELBO ---------------------------------------------------------------------------
    a_prev = torch.zeros(nb, C, S + 1, dtype=dt)
    a_prev[..., 0] = 1.0
    e_chain = torch.zeros(nb, C, dtype=dt)
    e_emit = torch.zeros(nb, C, dtype=dt)
    for f in range(F):
        joint = a_prev[..., :, None] * q_z[:, f]                                        # (nb,C,z',z)
        logp = logp_init[:, :, None, :].expand(nb, C, S + 1, S + 1) if f == 0 else logp_trans
        e_chain = e_chain + (joint * (logp - logq_z[:, f])).sum((-1, -2))
        a_prev = joint.sum(-2)
        e_emit = e_emit + (a_prev.permute(2, 0, 1) * emission[:, :, f]).sum(0)
    m2 = mask[:, 0, :]                                                                  # (nb,1) -> broadcast over C
    e_frame = (mask * e_b).sum() + (m2 * (e_chain + e_emit)).sum()
    total = e_global + sN * e_aoi + sN * e_frame
    if return_parts:
        return total, dict(e_global=e_global, e_aoi=e_aoi, e_frame=e_frame, emission=emission, Tz=Tz, L=L, q_m=q_m,
                           e_chain=e_chain, e_emit=e_emit, gain=gain, init=init, trans=trans, lamda=lamda, proximity=proximity,
                           background=background, height=height, width=width, x=x, y=y)
    return total


def loss_and_grads(unconstrained, data, ndx, noise, **kw):
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in unconstrained.items()}
    loss = -elbo(leaves, data, ndx, noise, **kw)
    loss.backward()
    return loss.item(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


def z_probs(unconstrained, data: O.OracleData):
    """hmm.py:627-633: forward marginals of the guide's chain, p(z_f = 1) -- computed there by a parallel scan of
    log z_trans (``_sequential_logmatmulexp``), here by the plain recursion."""
    p = to_constrained(unconstrained, data.P, data.dtype)
    zt = p["z_trans"]                                                                    # (Nt,F,C,z',z)
    a = zt[:, 0, :, 0, :]
    out = [a]
    for f in range(1, data.F):
        a = (a[..., :, None] * zt[:, f]).sum(-2)
        out.append(a)
    return torch.stack(out, 1)                                                          # (Nt,F,C,1+S)


@torch.no_grad()
def theta_probs(unconstrained, data: O.OracleData, ndx, noises, z_map, priors=DEFAULT_PRIORS, K=2, S=1):
    """
    hmm.py:541-625: for every particle, p(theta | z = z_MAP, m, x, y) -- the model's theta / m / x / y log-probs with the
    chain state fixed, normalised over theta (the chain's own terms cancel there) -- averaged over m with the guide's
    q(m | z_MAP), then over the particles.  Returns (K, nb, F, C).  ``z_map``: (nb, F, C) long.
    """
    dt, P = data.dtype, data.P
    nb, F, C = len(ndx), data.F, data.C
    half = (P + 1) / 2
    p = to_constrained(unconstrained, P, dt)
    loc = _gather_local(p, ndx)
    g = guide_dists(p, loc, P, priors)
    t = lambda v: torch.as_tensor(v, dtype=dt)
    mcfg = O.m_configs(K, dt)
    M = mcfg.shape[0]
    mp = torch.gather(loc["m_probs"], 0, z_map[None, None].expand(1, K, nb, F, C))[0]           # (K,nb,F,C): m_probs[z_MAP]
    logq_mk = D.Categorical(probs=torch.stack([1 - mp, mp], -1), validate_args=False).logits      # (K,nb,F,C,2)
    q_m = sum(logq_mk[k][..., mcfg[:, k].long()] for k in range(K)).exp()                      # (nb,F,C,M)
    logp_theta = D.Categorical(probs=O.probs_theta(K, dt), validate_args=False).logits[z_map.clamp(0, 1)]   # (nb,F,C,1+K)
    out = torch.zeros(K, nb, F, C, dtype=dt)
    for noise in noises:
        lamda = O.rsample_gamma(*g["lamda"], noise["lamda"])
        proximity = O.rsample_affine_beta(g["proximity"], noise["proximity"])
        size = torch.stack([torch.full_like(proximity, 2.0), ((P + 1) / (2 * proximity)) ** 2 - 1], -1)
        x = O.rsample_affine_beta(g["x"], noise["x"])
        y = O.rsample_affine_beta(g["y"], noise["y"])
        pm = O.probs_m(lamda, K)
        logp_mk = D.Categorical(probs=torch.stack([1 - pm, pm], -1), validate_args=False).logits      # (Q,1+K,K,2)
        xy_prior = [O.AffineBeta(t(0.0), size[s], -half, half) for s in range(2)]
        logp_xy = torch.stack([d.log_prob(x) + d.log_prob(y) for d in xy_prior])                    # (2,K,nb,F,C)
        terms = []
        for th in range(K + 1):
            term = logp_theta[..., th][..., None] + torch.zeros(nb, F, C, M, dtype=dt)
            for k in range(K):
                mk = mcfg[:, k]
                term = term + logp_mk[:, th, k][:, mk.long()][None, None] + mk * logp_xy[int(th == k + 1), k][..., None]
            terms.append(term)
        post = torch.softmax(torch.stack(terms), 0)                                                 # (1+K,nb,F,C,M)
        out += (q_m * post[1:]).sum(-1) / len(noises)
    return out
