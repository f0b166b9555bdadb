"""
Pins the CPU oracle (oracle/cosmos_oracle.py) WITHOUT a GPU.

1. Against the reference's own code: tests/golden/ref_distributions.pt holds outputs of the reference's
   ``distributions/util.py`` and ``distributions/ksmogn.py`` (torch branch) produced by
   tests/golden/make_golden.py in the build container.  Every function of the oracle that has a counterpart
   there is compared entry by entry (fp64, 1e-12).
2. tests/golden/ref_step.pt holds whole SVI iterations produced by the reference's own ``models/cosmos.py``
   (init_parameters, guide, model) and ``models/model.py`` (Model.init, svi.step) run verbatim by
   tests/golden/make_golden_step.py, with the absent pyro / pyroapi packages replaced by the restatement in
   tests/golden/minipyro.py: initial parameters, losses, all 20 gradients and the parameters after the updates.
3. What Pyro itself does (ELBO assembly, implicit reparameterisation gradients, Adam) therefore still rests on
   restatements ("parity unpinned" for Pyro's own machinery); they are cross-checked against code that
   share no code with the oracle: a scalar, loop-per-unit ELBO written with scipy.stats densities after
   models/cosmos.py:170-327 / :342-462; gradients against central differences of that ELBO with every sample
   moved along its quantile (the definition of the pathwise gradient torch's ``rsample`` implements); the
   optimiser against a hand-written Adam recurrence.
"""

import math

import numpy as np
import pytest
import scipy.special as sp
import scipy.stats as st
import torch

from oracle import cosmos_oracle as O
from tests.step_helpers import make_problem

TOL = 1e-12


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-300)
    assert (a - b).abs().max().item() <= tol * scale, ((a - b).abs().max().item(), scale)


# ---------------------------------------------------------------------------------------------------------------------
# 1. golden vectors of the reference's own functions
# ---------------------------------------------------------------------------------------------------------------------
def test_prior_tables_match_reference(golden):
    n = 0
    for name, case in golden["tables"].items():
        if name.startswith("probs_m"):
            close(O.probs_m(case["lamda"], case["K"]), case["probs_m"])           # util.py:94-130
            close(O.truncated_poisson_probs(case["lamda"], case["K"]), case["trunc"])  # util.py:67-91
            n += 1
        elif name.startswith("probs_theta"):
            close(O.probs_theta(int(name[-1])), case)                             # util.py:154-173
            n += 1
    case = golden["tables"]["expand_offtarget"]
    close(O.expand_offtarget(case["probs"]), case["out"])                         # util.py:133-151
    assert n == 8


@pytest.mark.parametrize("name", ["sim_O3", "hist_O16_C2", "small_P6"])
def test_gaussian_spots_and_image_match_reference(golden, name):
    case = golden["ksmogn"][name]
    c = case["inputs"]
    K = c["height"].shape[-1]
    close(O.gaussian_spots(c["height"], c["width"], c["x"], c["y"], c["target_locs"].unsqueeze(-2), c["P"]), case["spots"])
    m_enum = case["m"].reshape(2, 2, 1, 1, 1, K)   # (m_1, m_0, nb, fb, C, K): Pyro's enumeration dims, cosmos.py:419-425
    close(O.gaussian_spots(c["height"], c["width"], c["x"], c["y"], c["target_locs"].unsqueeze(-2), c["P"], m_enum),
          case["spots_m"])
    close(O.ksmogn_image(c["height"], c["width"], c["x"], c["y"], c["target_locs"], c["background"], c["P"], m_enum),
          case["image"])
    # the flat table the oracle enumerates with is the row-major flattening of Pyro's (m_1, m_0) dims
    assert torch.equal(O.m_configs(K, torch.float64), case["m"])


@pytest.mark.parametrize("name", ["sim_O3", "hist_O16_C2", "small_P6"])
def test_ksmogn_log_prob_and_gradients_match_reference(golden, name):
    case = golden["ksmogn"][name]
    c = case["inputs"]
    leaves = {k: c[k].clone().requires_grad_(True) for k in ("height", "width", "x", "y", "background", "gain")}
    logits = torch.distributions.utils.probs_to_logits(c["offset_weights"])      # dataset.py:27-29
    m = case["m"][:, None, None, None, :]
    logp = O.ksmogn_log_prob(leaves["height"], leaves["width"], leaves["x"], leaves["y"], c["target_locs"],
                             leaves["background"], leaves["gain"], c["offset_samples"], logits, c["P"], c["value"], m=m)
    close(logp.detach(), case["log_prob"])
    (case["W"] * logp).sum().backward()
    for k, v in leaves.items():
        close(v.grad, case["grads"][k], 1e-11)
    nom = O.ksmogn_log_prob(c["height"], c["width"], c["x"], c["y"], c["target_locs"], c["background"], c["gain"],
                            c["offset_samples"], logits, c["P"], c["value"])
    close(nom, case["log_prob_no_m"])
    # every spot present == no m argument (ksmogn.py:146-165)
    close(logp.detach()[-1], case["log_prob_no_m"])


def test_pixel_at_or_below_an_offset_is_excluded_not_nan(golden):
    """ksmogn.py:225-236: offsets at or above the pixel value drop out of the mixture (log 0), they do not poison it."""
    case = golden["ksmogn"]["hist_O16_C2"]
    c = case["inputs"]
    assert (c["value"][0, 0, 0, 0, 0] <= c["offset_samples"]).any() and (c["value"][0, 0, 0, 0, 0] > c["offset_samples"]).any()
    assert torch.isfinite(case["log_prob"]).all()


# ---------------------------------------------------------------------------------------------------------------------
# 2. scalar restatement of the ELBO (scipy.stats, one unit at a time)
# ---------------------------------------------------------------------------------------------------------------------
EPS = float(np.finfo(np.float64).eps)


def _clog(p):
    """log of a probability as torch's Categorical / Bernoulli(probs=...) see it: clamped to [eps, 1 - eps]."""
    return math.log(min(max(p, EPS), 1 - EPS))


def _affine_beta_logpdf(v, mean, size, lo, hi):
    a, b = size * (mean - lo) / (hi - lo), size * (hi - mean) / (hi - lo)     # affine_beta.py:33-49
    return st.beta.logpdf((v - lo) / (hi - lo), a, b) - math.log(hi - lo)


def _trunc_poisson(lam, K):
    head = [st.poisson.pmf(k, lam) for k in range(K)]
    return head + [1 - sum(head)]


def _p_m(lam, K, theta, k):
    """util.py:94-130: spot k certainly present when theta points at it, otherwise the expected occupancy of the
    non-specific slots under the truncated Poisson."""
    if theta == k + 1:
        return 1.0
    if theta == 0:
        tp = _trunc_poisson(lam, K)
        return sum(l * tp[l] for l in range(1, K + 1)) / K
    tp = _trunc_poisson(lam, K - 1)
    return sum(l * tp[l] for l in range(1, K)) / (K - 1)


def scalar_elbo(cons, s, data, ndx, fdx, priors, K=2):
    """ELBO of models/cosmos.py:170-327 under the guide :342-462 for GIVEN constrained parameters ``cons`` and guide
    samples ``s``; z and theta summed out in the model, m_k enumerated in the guide, plates scaled by Nt/nb and F/fb,
    masked AOIs dropped.  numpy scalars and loops only."""
    P, C = data.P, data.C
    nb, fb = len(ndx), len(fdx)
    half = (P + 1) / 2
    f = lambda t: np.asarray(t.detach().double())
    c = {k: f(v) for k, v in cons.items()}
    s = {k: f(v) for k, v in s.items()}
    pix, xy = f(data.images), f(data.xy)
    off, offw = f(data.offset_samples), f(data.offset_weights)
    gain, prox = float(s["gain"]), float(s["proximity"])
    # ---- global sites (cosmos.py:170-184 / :342-368)
    e = st.halfnorm.logpdf(gain, scale=priors["gain_std"]) - st.gamma.logpdf(gain, c["gain_loc"] * c["gain_beta"], scale=1 / c["gain_beta"])
    for q in range(C):
        alpha = c["pi_mean"][q] * c["pi_size"][q]
        pv = s["pi"][q] / s["pi"][q].sum()
        e += st.dirichlet.logpdf(pv, np.full(2, 0.5)) - st.dirichlet.logpdf(pv, alpha)
        lam = s["lamda"][q]
        e += st.expon.logpdf(lam, scale=1 / priors["lamda_rate"]) - st.gamma.logpdf(lam, c["lamda_loc"][q] * c["lamda_beta"][q], scale=1 / c["lamda_beta"][q])
    pmax = (P + 1) / math.sqrt(12)
    e += st.expon.logpdf(prox, scale=1 / priors["proximity_rate"]) - _affine_beta_logpdf(prox, float(c["proximity_loc"]), float(c["proximity_size"]), 0.0, pmax)
    size_prior = [2.0, ((P + 1) / (2 * prox)) ** 2 - 1]                      # cosmos.py:185-191
    ii = np.arange(P)
    sN, sF = data.Nt / nb, data.F / fb
    for a, n in enumerate(ndx.tolist()):
        if not bool(data.mask[n]):
            continue
        ont = bool(data.is_ontarget[n])
        for ch in range(C):
            bm, bs = c["background_mean_loc"][n, 0, ch], c["background_std_loc"][n, 0, ch]
            e += sN * (st.halfnorm.logpdf(bm, scale=priors["background_mean_std"]) + st.halfnorm.logpdf(bs, scale=priors["background_std_std"]))
            pz = [1 - s["pi"][ch][1], s["pi"][ch][1]] if ont else [1.0, 0.0]   # expand_offtarget, util.py:133-151
            for b_, fr in enumerate(fdx.tolist()):
                bg = s["background"][a, b_, ch]
                u = st.gamma.logpdf(bg, (bm / bs) ** 2, scale=bs ** 2 / bm)
                u -= st.gamma.logpdf(bg, c["b_loc"][n, fr, ch] * c["b_beta"][n, fr, ch], scale=1 / c["b_beta"][n, fr, ch])
                h = [s["height"][k, a, b_, ch] for k in range(K)]
                w = [s["width"][k, a, b_, ch] for k in range(K)]
                x = [s["x"][k, a, b_, ch] for k in range(K)]
                y = [s["y"][k, a, b_, ch] for k in range(K)]
                tx, ty = xy[n, fr, ch]
                D = pix[n, fr, ch]
                for mbits in range(2 ** K):
                    m = [(mbits >> k) & 1 for k in range(K)]
                    qm = 1.0
                    for k in range(K):
                        p1 = min(max(c["m_probs"][k, n, fr, ch], EPS), 1 - EPS)
                        qm *= p1 if m[k] else 1 - p1
                    # model: sum over z, theta (cosmos.py:242-300)
                    tot = 0.0
                    for z in range(2):
                        for th in range(K + 1):
                            pth = (1.0 if th == 0 else 0.0) if z == 0 else (0.0 if th == 0 else 1.0 / K)
                            lp = _clog(pz[z]) + _clog(pth)
                            for k in range(K):
                                pm1 = min(max(_p_m(s["lamda"][ch], K, th, k), EPS), 1 - EPS)
                                lp += math.log(pm1 if m[k] else 1 - pm1)
                                if m[k]:
                                    sz = size_prior[1 if th == k + 1 else 0]
                                    lp += _affine_beta_logpdf(x[k], 0.0, sz, -half, half) + _affine_beta_logpdf(y[k], 0.0, sz, -half, half)
                            tot += math.exp(lp)
                    v = math.log(tot) - math.log(qm)
                    for k in range(K):
                        if not m[k]:
                            continue
                        v += st.halfnorm.logpdf(h[k], scale=priors["height_std"])
                        v += _affine_beta_logpdf(w[k], 1.5, 2.0, priors["width_min"], priors["width_max"])
                        hl, hb = c["h_loc"][k, n, fr, ch], c["h_beta"][k, n, fr, ch]
                        v -= st.gamma.logpdf(h[k], hl * hb, scale=1 / hb)
                        v -= _affine_beta_logpdf(w[k], c["w_mean"][k, n, fr, ch], c["w_size"][k, n, fr, ch], priors["width_min"], priors["width_max"])
                        v -= _affine_beta_logpdf(x[k], c["x_mean"][k, n, fr, ch], c["size"][k, n, fr, ch], -half, half)
                        v -= _affine_beta_logpdf(y[k], c["y_mean"][k, n, fr, ch], c["size"][k, n, fr, ch], -half, half)
                    # likelihood (ksmogn.py:146-238): rows = y pixel j, columns = x pixel i (util.py:46-48)
                    img = np.full((P, P), bg)
                    for k in range(K):
                        if m[k]:
                            gx = np.exp(-((ii - x[k] - tx) ** 2) / (2 * w[k] ** 2))
                            gy = np.exp(-((ii - y[k] - ty) ** 2) / (2 * w[k] ** 2))
                            img = img + h[k] / (2 * math.pi * w[k] ** 2) * gy[:, None] * gx[None, :]
                    d = D[:, :, None] - off[None, None, :]
                    ok = d > 0
                    lg = np.where(ok, st.gamma.logpdf(np.where(ok, d, 1.0), (img / gain)[:, :, None], scale=gain) + np.log(offw), -np.inf)
                    v += sp.logsumexp(lg, axis=-1).sum()
                    u += qm * v
                e += sN * sF * u
    return float(e)


def _samples_from_parts(parts):
    return {k: parts[k] for k in ("gain", "pi", "lamda", "proximity", "background", "height", "width", "x", "y")}


@pytest.mark.parametrize("cfg", [dict(N=4, F=5, C=1, nb=3, fb=3, seed=0, offsets="sim"),
                                 dict(N=3, F=4, C=2, nb=2, fb=3, seed=3, offsets="hist"),
                                 dict(N=4, F=3, C=1, nb=4, fb=3, seed=5, offsets="sim")])
def test_elbo_matches_the_scalar_restatement(cfg):
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    if cfg["seed"] == 0:
        data.mask[ndx[0]] = False               # a masked AOI contributes nothing (cosmos.py:218-220)
    assert data.is_ontarget[ndx].any() and (not data.is_ontarget[ndx].all() or cfg["N"] < 4)
    total, parts = O.elbo(params, data, ndx, fdx, noise, return_parts=True)
    cons = O.to_constrained(params, data.P, data.dtype)
    ref = scalar_elbo(cons, _samples_from_parts(parts), data, ndx, fdx, O.DEFAULT_PRIORS)
    assert abs(total.item() - ref) <= 1e-10 * abs(ref), (total.item(), ref)


# ---------------------------------------------------------------------------------------------------------------------
# 3. gradients: finite differences of the scalar ELBO with the samples riding their quantiles
# ---------------------------------------------------------------------------------------------------------------------
def _quantile_samples(cons, u, data, ndx, fdx, priors, K=2):
    """Guide samples as functions of the parameters at FIXED quantiles ``u`` (inverse-cdf reparameterisation).
    The pathwise derivative of torch's Gamma/Beta/Dirichlet ``rsample`` is by definition the derivative of this map
    (dx/dalpha = -(dF/dalpha)/f(x)); ATen evaluates it with series/rational approximations."""
    P = data.P
    half = (P + 1) / 2
    f = lambda t: np.asarray(t.detach().double())
    c = {k: f(v) for k, v in cons.items()}
    n, fr = ndx.numpy()[:, None], fdx.numpy()[None, :]
    gam = lambda q, loc, beta: st.gamma.ppf(q, loc * beta, scale=1 / beta)

    def abeta(q, mean, size, lo, hi):
        return lo + (hi - lo) * st.beta.ppf(q, size * (mean - lo) / (hi - lo), size * (hi - mean) / (hi - lo))

    s = {"gain": gam(u["gain"], c["gain_loc"], c["gain_beta"]), "lamda": gam(u["lamda"], c["lamda_loc"], c["lamda_beta"]),
         "proximity": abeta(u["proximity"], c["proximity_loc"], c["proximity_size"], 0.0, (P + 1) / math.sqrt(12))}
    alpha = c["pi_mean"] * c["pi_size"]
    p0 = st.beta.ppf(u["pi"], alpha[:, 0], alpha[:, 1])
    s["pi"] = np.stack([p0, 1 - p0], -1)
    s["background"] = gam(u["background"], c["b_loc"][n, fr], c["b_beta"][n, fr])
    loc = lambda name: c[name][:, n, fr]
    s["height"] = gam(u["height"], loc("h_loc"), loc("h_beta"))
    s["width"] = abeta(u["width"], loc("w_mean"), loc("w_size"), priors["width_min"], priors["width_max"])
    s["x"] = abeta(u["x"], loc("x_mean"), loc("size"), -half, half)
    s["y"] = abeta(u["y"], loc("y_mean"), loc("size"), -half, half)
    return {k: torch.as_tensor(v) for k, v in s.items()}


def _quantiles_of(cons, s, data, ndx, fdx, priors):
    P = data.P
    half = (P + 1) / 2
    f = lambda t: np.asarray(t.detach().double())
    c = {k: f(v) for k, v in cons.items()}
    s = {k: f(v) for k, v in s.items()}
    n, fr = ndx.numpy()[:, None], fdx.numpy()[None, :]
    gam = lambda x, loc, beta: st.gamma.cdf(x, loc * beta, scale=1 / beta)

    def abeta(x, mean, size, lo, hi):
        return st.beta.cdf((x - lo) / (hi - lo), size * (mean - lo) / (hi - lo), size * (hi - mean) / (hi - lo))

    alpha = c["pi_mean"] * c["pi_size"]
    loc = lambda name: c[name][:, n, fr]
    return {"gain": gam(s["gain"], c["gain_loc"], c["gain_beta"]), "lamda": gam(s["lamda"], c["lamda_loc"], c["lamda_beta"]),
            "proximity": abeta(s["proximity"], c["proximity_loc"], c["proximity_size"], 0.0, (P + 1) / math.sqrt(12)),
            "pi": st.beta.cdf(s["pi"][:, 0], alpha[:, 0], alpha[:, 1]),
            "background": gam(s["background"], c["b_loc"][n, fr], c["b_beta"][n, fr]),
            "height": gam(s["height"], loc("h_loc"), loc("h_beta")),
            "width": abeta(s["width"], loc("w_mean"), loc("w_size"), priors["width_min"], priors["width_max"]),
            "x": abeta(s["x"], loc("x_mean"), loc("size"), -half, half),
            "y": abeta(s["y"], loc("y_mean"), loc("size"), -half, half)}


def test_gradients_match_quantile_finite_differences():
    """d(-ELBO)/d(unconstrained parameter) of the oracle (autograd through torch.distributions + ATen's implicit
    reparameterisation gradients) against central differences of the scalar ELBO in which every guide sample follows
    its quantile.  Tolerance 3e-4 of the gradient's size: ATen's gradient approximations are good to ~2e-4 and the
    differences to ~1e-6.  ``proximity_size`` moves both Beta concentrations at a fixed mean: the two implicit
    gradients nearly cancel (1.31e-2 * a/S vs -1.42e-3 * b/S) and ATen's 2e-4 error on each is 5e-3 of what is left --
    checked directly against d ppf / d concentration; the oracle inherits ATen's value on purpose (so does the
    reference)."""
    ds, data, params, ndx, fdx, noise = make_problem(N=3, F=3, C=1, nb=2, fb=2, seed=2, offsets="sim")
    pri = O.DEFAULT_PRIORS
    _, grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    _, parts = O.elbo(params, data, ndx, fdx, noise, return_parts=True)
    cons = O.to_constrained(params, data.P, data.dtype)
    s0 = _samples_from_parts(parts)
    u = _quantiles_of(cons, s0, data, ndx, fdx, pri)
    # the quantile map reproduces the samples it was inverted from
    for k, v in _quantile_samples(cons, u, data, ndx, fdx, pri).items():
        close(v.reshape(s0[k].shape), s0[k], 1e-8)

    def loss_at(p):
        c = O.to_constrained(p, data.P, data.dtype)
        return -scalar_elbo(c, _quantile_samples(c, u, data, ndx, fdx, pri), data, ndx, fdx, pri)

    n0, f0 = int(ndx[0]), int(fdx[1])
    probes = [("gain_loc", ()), ("gain_beta", ()), ("lamda_loc", (0,)), ("lamda_beta", (0,)), ("proximity_loc", ()),
              ("proximity_size", ()), ("pi_mean", (0, 1)), ("pi_size", (0, 0)), ("background_mean_loc", (n0, 0, 0)),
              ("background_std_loc", (n0, 0, 0)), ("b_loc", (n0, f0, 0)), ("b_beta", (n0, f0, 0)), ("m_probs", (0, n0, f0, 0)),
              ("h_loc", (1, n0, f0, 0)), ("h_beta", (0, n0, f0, 0)), ("w_mean", (0, n0, f0, 0)), ("w_size", (1, n0, f0, 0)),
              ("x_mean", (0, n0, f0, 0)), ("y_mean", (1, n0, f0, 0)), ("size", (0, n0, f0, 0))]
    loose = {"proximity_size": 3e-3}
    for name, idx in probes:
        step = 1e-5
        vals = []
        for sign in (+1, -1):
            p = {k: v.clone() for k, v in params.items()}
            p[name][idx] += sign * step
            vals.append(loss_at(p))
        fd = (vals[0] - vals[1]) / (2 * step)
        g = grads[name][idx].item()
        scale = max(grads[name].abs().max().item(), 1e-8)
        assert abs(fd - g) <= loose.get(name, 3e-4) * scale, (name, idx, fd, g, scale)
    # an entry outside the minibatch gets a zero gradient (and still an Adam update: dense optimiser)
    outside = [n for n in range(data.Nt) if n not in ndx.tolist()][0]
    assert grads["h_loc"][:, outside].abs().max().item() == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# 4. the optimiser loop (models/model.py:168-171: Adam(lr, betas=(0.9, 0.999)) on the unconstrained values)
# ---------------------------------------------------------------------------------------------------------------------
def test_svi_step_is_dense_adam_on_the_unconstrained_values():
    ds, data, params, ndx, fdx, noise = make_problem(N=3, F=4, C=1, nb=2, fb=2, seed=1, perturb=False)
    svi = O.OracleSVI(data, lr=0.005, nbatch_size=2, fbatch_size=2)
    start = {k: v.detach().clone() for k, v in svi.params.items()}
    m = {k: torch.zeros_like(v) for k, v in start.items()}
    v2 = {k: torch.zeros_like(v) for k, v in start.items()}
    mine = {k: v.clone() for k, v in start.items()}
    g = torch.Generator().manual_seed(7)
    for t in range(1, 4):
        nd, fd = torch.randperm(3, generator=g)[:2], torch.randperm(4, generator=g)[:2]
        nz = O.draw_noise(mine, data, nd, fd, g)
        loss_ref, grads = O.loss_and_grads(mine, data, nd, fd, nz)
        loss = svi.step(nd, fd, nz)
        assert loss == loss_ref
        for k in mine:
            m[k] = 0.9 * m[k] + 0.1 * grads[k]
            v2[k] = 0.999 * v2[k] + 0.001 * grads[k] ** 2
            mine[k] = mine[k] - 0.005 * (m[k] / (1 - 0.9 ** t)) / ((v2[k] / (1 - 0.999 ** t)).sqrt() + 1e-8)
            close(svi.params[k].detach(), mine[k], 1e-12)
    # frames outside every minibatch moved too (momentum): the update is dense
    assert (svi.params["h_loc"].detach() - start["h_loc"]).abs().min().item() >= 0.0
    assert svi.iter == 3


def test_initial_values_and_constraints_round_trip():
    """cosmos.py:471-598: initial constrained values survive constrained -> unconstrained -> constrained."""
    ds, data, params, ndx, fdx, noise = make_problem(N=3, F=4, C=2, nb=2, fb=2, seed=4, perturb=False)
    init = O.init_constrained(data)
    back = O.to_constrained(O.to_unconstrained(init, data.P, data.dtype), data.P, data.dtype)
    for k in init:
        # pi_mean is initialised at ones (cosmos.py:471-476), which the simplex constraint reads back normalised
        close(back[k], init[k] / init[k].sum(-1, keepdim=True) if k == "pi_mean" else init[k], 1e-12)
    assert init["m_probs"].shape == (2, 3, 4, 2) and init["pi_mean"].shape == (2, 2) and init["b_loc"].shape == (3, 4, 2)
    assert set(init) == set(O.PARAM_NAMES) and len(O.PARAM_NAMES) == 20
    # background initialised at median - mean offset per channel (cosmos.py:519-541)
    close(init["b_loc"][0, 0], data.median - data.offset_mean)


def test_compute_probs_is_a_distribution_and_follows_the_data():
    """cosmos.py:609-672: z_probs sums to one over z, theta_probs[k] <= p(z = 1), off-target AOIs carry (almost) no
    specific binding."""
    ds, data, params, ndx, fdx, noise = make_problem(N=4, F=5, C=1, nb=4, fb=5, seed=0, perturb=False)
    ndx, fdx = torch.arange(4), torch.arange(5)
    g = torch.Generator().manual_seed(0)
    noises = [O.draw_noise(params, data, ndx, fdx, g) for _ in range(3)]
    z, th = O.compute_probs(params, data, ndx, fdx, noises)
    close(z.sum(-1), torch.ones(4, 5, 1), 1e-12)
    close(th.sum(0), z[..., 1], 1e-12)
    off = ~data.is_ontarget
    assert z[off][..., 1].max().item() < 1e-10


# ---------------------------------------------------------------------------------------------------------------------
# 5. the reference's own model / guide / parameter code (cosmos.py, model.py run under tests/golden/minipyro.py)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_steps():
    from pathlib import Path

    return torch.load(Path(__file__).resolve().parent / "golden" / "ref_step.pt", weights_only=False)


def _golden_data(case):
    return O.OracleData(case["images"].double(), case["xy"], case["is_ontarget"], case["mask"], case["offset_samples"],
                        case["offset_weights"])


STEP_CASES = ["c1_initial_point", "c1_perturbed_masked", "c2_hist_offsets", "c1_full_batch"]


@pytest.mark.parametrize("name", STEP_CASES)
def test_initial_parameters_match_reference_init(ref_steps, name):
    """cosmos.init_parameters (cosmos.py:464-598) as stored by pyro.param: names, shapes, unconstrained values."""
    case = ref_steps[name]
    data = _golden_data(case)
    mine = O.to_unconstrained(O.init_constrained(data), data.P, data.dtype)
    assert set(mine) == set(case["init_unconstrained"])
    for k, v in case["init_unconstrained"].items():
        close(mine[k], v.reshape(mine[k].shape) if v.numel() == mine[k].numel() and v.dim() != mine[k].dim() else v, 1e-13)


@pytest.mark.parametrize("name", STEP_CASES)
def test_svi_iterations_match_reference_model_code(ref_steps, name):
    """Every iteration of the reference's ``svi.step()`` (its guide(), model(), Adam) on the recorded minibatches and
    base variates: loss 1e-12, every gradient 1e-8 of the tensor's largest entry (measured <= 7e-9, on ``size`` /
    ``w_size`` whose two Beta-concentration terms nearly cancel, from the one-ulp round trip of the recorded variates),
    parameters after the last update 1e-10."""
    case = ref_steps[name]
    cfg = case["config"]
    data = _golden_data(case)
    svi = O.OracleSVI(data, lr=cfg["lr"], nbatch_size=cfg["nb"], fbatch_size=cfg["fb"])
    with torch.no_grad():
        for k, v in svi.params.items():
            v.copy_(case["start"][k].reshape(v.shape))
    for it, step in enumerate(case["steps"]):
        if cfg["nb"] < cfg["N"]:
            assert len(step["ndx"]) == cfg["nb"] and len(set(step["ndx"].tolist())) == cfg["nb"]
        loss = svi.step(step["ndx"], step["fdx"], step["noise"])
        # Masked AOIs (cosmos.py:218-220): Pyro zeroes the log-probabilities of masked sites but still sums over the
        # values of the enumerated ones, so every masked unit adds the parameter-independent constant
        # 4 configurations x ln(2 * 3 states of (z, theta)) to the reference's reported ELBO.  The oracle (and the
        # kernels) report the ELBO of the unmasked AOIs only; gradients, updates and fits are the same (checked below).
        n_masked = int((~case["mask"][step["ndx"]]).sum())
        const = n_masked * cfg["fb"] * cfg["C"] * 4 * math.log(6) * (cfg["N"] / cfg["nb"]) * (cfg["F"] / cfg["fb"])
        assert abs(loss - const - step["loss"]) <= 1e-12 * abs(step["loss"]), (it, loss, const, step["loss"])
        bad = {}
        for k, g in step["grads"].items():
            ref = g.reshape(svi.last_grads[k].shape)
            err = (svi.last_grads[k] - ref).abs().max().item()
            if err > 1e-8 * max(ref.abs().max().item(), 1e-30):
                bad[k] = err
        assert not bad, (it, bad)
    for k, v in case["final"].items():
        close(svi.params[k].detach(), v.reshape(svi.params[k].shape), 1e-10)
    # Pyro's enumeration layout the kernels' configuration index follows: m_0 at dim -4, m_1 at -5, z at -6, theta at -7
    assert case["steps"][0]["enum_shapes"] == {"z": (2, 1, 1, 1, 1, 1), "theta": (3, 1, 1, 1, 1, 1, 1), "m_k0": (2, 1, 1, 1),
                                               "m_k1": (2, 1, 1, 1, 1)}


def _particles(case):
    p = case["probs"]["particles"]
    return [{k: v[i] for k, v in p.items()} for i in range(p["pi"].shape[0])]


@pytest.mark.parametrize("name", STEP_CASES)
def test_compute_probs_matches_reference_model_code(ref_steps, name):
    """cosmos.compute_probs (cosmos.py:609-672) run verbatim at the parameters after the recorded iterations, 50 guide
    particles: z_probs and theta_probs of the on-target AOIs (the reference leaves off-target AOIs at zero).
    The reference sums Pyro's ``unscaled_log_prob`` of x_k, y_k, which is taken BEFORE the ``m_k > 0`` mask; the oracle
    applies the mask.  The two differ only where p(m_k = 0 | theta = k) = eps enters: below 1e-9 in the probabilities."""
    case = ref_steps[name]
    data = _golden_data(case)
    n_on = case["probs"]["n_on"]
    assert bool(data.is_ontarget[:n_on].all()) and not bool(data.is_ontarget[n_on:].any())
    final = {k: v.reshape(O.init_constrained(data)[k].shape) for k, v in case["final"].items()}
    z, th = O.compute_probs(final, data, torch.arange(n_on), torch.arange(data.F), _particles(case))
    zr, thr = case["probs"]["z_probs"], case["probs"]["theta_probs"]
    assert zr.shape == (data.Nt, data.F, data.C, 2) and thr.shape == (2, data.Nt, data.F, data.C)
    assert (z - zr[:n_on]).abs().max().item() <= 1e-9 and (th - thr[:, :n_on]).abs().max().item() <= 1e-9
    assert zr[n_on:].abs().max().item() == 0 and thr[:, n_on:].abs().max().item() == 0
    assert 0.0 < zr[:n_on, ..., 1].max().item() <= 1.0


def test_c1_hundred_iterations_match_the_reference_run():
    """BASELINE configs[0]: the reference's own 100-iteration fit at N=5 x F=100 (tests/golden/ref_c1_fit.pt), replayed by
    the oracle from the same seed: every loss 1e-12 (measured 1.4e-14), parameters after 100 Adam updates 1e-8
    (measured 2.5e-10)."""
    from tests.step_helpers import golden_c1_fit

    ds, data, case = golden_c1_fit()
    cfg = case["config"]
    svi = O.OracleSVI(data, lr=cfg["lr"], nbatch_size=cfg["nb"], fbatch_size=cfg["fb"])
    ndx, fdx = torch.arange(cfg["N"]), torch.arange(cfg["F"])
    state = torch.get_rng_state()
    try:
        torch.manual_seed(cfg["rng_seed"])
        for it in range(cfg["iters"]):
            noise = O.draw_noise(svi.params, data, ndx, fdx)
            if it == 0:     # the random stream lines up with the reference's guide
                for k, v in case["first_noise"].items():
                    close(noise[k], v.reshape(noise[k].shape), 1e-14)
            loss = svi.step(ndx, fdx, noise)
            assert abs(loss - case["losses"][it].item()) <= 1e-12 * abs(loss), (it, loss, case["losses"][it].item())
    finally:
        torch.set_rng_state(state)
    for k, v in case["final"].items():
        assert (svi.params[k].detach() - v.reshape(svi.params[k].shape)).abs().max().item() <= 1e-8, k
    assert case["losses"][-10:].mean() < 0.9 * case["losses"][:10].mean()       # and it is a fit: the loss went down
