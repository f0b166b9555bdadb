"""
CPU oracle for the cosmos SVI hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker.  The shipped path
(``tapqir_b200``) never routes through it.

What it is: a pure-PyTorch (CPU, fp64 by default) restatement of what one ``svi.step()`` of the
reference's cosmos model computes.  Each function cites the reference ``file:line`` it follows
(paths relative to the reference checkout).  Everything Pyro does around the model (plates,
enumeration, Dice weights, param store, Adam wrapper) lives in an un-vendored third-party
dependency (pyro-ppl >= 1.8.5, ``setup.py:69``) that cannot be installed here; its semantics are
written out explicitly following SURVEY.md Appendix A.

Pinning status
--------------
* ``gaussian_spots``, ``probs_m``, ``truncated_poisson_probs``, ``probs_theta``,
  ``expand_offtarget`` and ``ksmogn_log_prob`` are PINNED: ``tests/golden/make_golden.py`` imports the
  reference's own ``tapqir/distributions/util.py`` and ``ksmogn.py`` (torch branch,
  ``use_pykeops=False``) in the build container and stores their outputs in
  ``tests/golden/ref_distributions.pt``; ``tests/test_oracle.py`` checks this file against them.
* One whole ``svi.step()`` (loss, all 20 gradients, Adam update), the initial parameters and ``compute_probs`` are
  PINNED against the reference's own ``models/cosmos.py`` / ``models/model.py`` run verbatim by
  ``tests/golden/make_golden_step.py`` -> ``tests/golden/ref_step.pt`` (``tests/test_oracle.py``), with the absent
  pyro / pyroapi packages replaced by the restatement ``tests/golden/minipyro.py``.
* Pyro's own machinery (what ``TraceEnum_ELBO`` does with the traces, ``pyro.distributions.AffineBeta``, the param
  store, ``pyro.optim.Adam``) therefore remains UNPINNED ("parity unpinned"): the reference's tests assert only
  ``exit_code == 0`` (``test/test_tapqir.py:91-93``) and Pyro cannot be imported.  It is cross-checked against a
  scalar scipy restatement of the ELBO, quantile finite differences and a hand-written Adam instead.
"""

import itertools
import math
from typing import Dict, Optional

import torch
import torch.distributions as D
from torch.distributions import constraints, transform_to

DEFAULT_PRIORS = {  # models/cosmos.py:55-64
    "background_mean_std": 1000.0,
    "background_std_std": 100.0,
    "lamda_rate": 1.0,
    "height_std": 10000.0,
    "width_min": 0.75,
    "width_max": 2.25,
    "proximity_rate": 1.0,
    "gain_std": 50.0,
}


# ------------------------------------------------------------------------------------------------
# distributions/util.py
# ------------------------------------------------------------------------------------------------
def gaussian_spots(height, width, x, y, target_locs, P, m=None):
    """
    ``m*h/(2 pi w^2) * exp(-((i-x-tx)^2 + (j-y-ty)^2)/(2 w^2))`` on a PxP grid whose LAST axis is
    the x pixel i and second-to-last the y pixel j.  Follows distributions/util.py:15-64
    (``meshgrid(..., indexing="xy")`` at :46-48).  ``height/width/x/y`` are ``(..., K)``,
    ``target_locs`` is ``(..., 1, 2)``; result is ``(..., K, P, P)``.
    """
    grid = torch.arange(P, dtype=height.dtype, device=height.device)
    cx = x + target_locs[..., 0]
    cy = y + target_locs[..., 1]
    w2 = width**2
    ex = -((grid - cx[..., None]) ** 2) / (2 * w2[..., None])  # (..., K, P) along x
    ey = -((grid - cy[..., None]) ** 2) / (2 * w2[..., None])  # (..., K, P) along y
    shape = torch.exp(ey[..., :, None] + ex[..., None, :]) / (2 * math.pi * w2)[..., None, None]
    if m is not None:
        height = m * height
    return height[..., None, None] * shape


def truncated_poisson_probs(lamda, K):
    """distributions/util.py:67-91: Poisson pmf for k<K, remaining mass at k=K."""
    k = torch.arange(K, dtype=lamda.dtype)
    head = torch.exp(torch.xlogy(k, lamda[..., None]) - lamda[..., None] - torch.lgamma(k + 1))
    return torch.cat([head, 1 - head.sum(-1, keepdim=True)], -1)


def probs_m(lamda, K):
    """distributions/util.py:94-130: p(m_k = 1 | theta, lamda), shape lamda.shape + (1+K, K)."""
    out = torch.zeros(lamda.shape + (1 + K, K), dtype=lamda.dtype)
    # theta = other spot: expected fraction of K-1 non-specific slots occupied
    tp = truncated_poisson_probs(lamda, K - 1)
    l = torch.arange(1, K, dtype=lamda.dtype)
    out = out + ((l * tp[..., 1:K]).sum(-1) / (K - 1))[..., None, None]
    # theta = 0: expected fraction of K slots occupied
    tp = truncated_poisson_probs(lamda, K)
    l = torch.arange(1, K + 1, dtype=lamda.dtype)
    row0 = ((l * tp[..., 1 : K + 1]).sum(-1) / K)[..., None].expand(lamda.shape + (K,))
    rows = [row0]
    eye = torch.eye(K, dtype=lamda.dtype)
    for k in range(K):
        # theta = k+1: spot k is certainly present
        rows.append(torch.where(eye[k] > 0, torch.ones_like(row0), out[..., 1 + k, :]))
    return torch.stack(rows, -2)


def expand_offtarget(probs):
    """distributions/util.py:133-151: stack [off-target = (1,0,..,0), on-target = probs] on a new last axis."""
    off = torch.zeros_like(probs)
    off[..., 0] = 1
    return torch.stack([off, probs], -1)


def probs_theta(K, dtype=torch.float64):
    """distributions/util.py:154-173: p(theta | z) rows z=0 -> e_0, z>0 -> uniform over 1..K."""
    out = torch.zeros(2, 1 + K, dtype=dtype)
    out[0, 0] = 1
    out[1, 1:] = 1 / K
    return out


# ------------------------------------------------------------------------------------------------
# distributions/ksmogn.py (torch branch) and affine_beta.py
# ------------------------------------------------------------------------------------------------
def ksmogn_image(height, width, x, y, target_locs, background, P, m=None):
    """distributions/ksmogn.py:146-165: background + sum over spots."""
    spots = gaussian_spots(height, width, x, y, target_locs.unsqueeze(-2), P, m)
    return background[..., None, None] + spots.sum(-3)


def ksmogn_log_prob(height, width, x, y, target_locs, background, gain, offset_samples,
                    offset_logits, P, value, m=None, pixel_weights=None):
    """
    distributions/ksmogn.py:222-238 (``use_pykeops=False`` branch): per pixel
    ``logsumexp_j[ w_j + log 1(D>d_j) + Gamma(D-d_j; image/gain, 1/gain).log_prob ]``, summed over PxP.
    """
    image = ksmogn_image(height, width, x, y, target_locs, background, P, m)
    rate = 1 / gain
    conc = (image * rate)[..., None]
    v = value[..., None]
    ok = v > offset_samples
    yy = torch.where(ok, v - offset_samples, torch.ones((), dtype=v.dtype))
    per_offset = (conc * torch.log(rate) + (conc - 1) * torch.log(yy) - rate * yy
                  - torch.lgamma(conc) + offset_logits + torch.log(ok.to(v.dtype)))
    per_pixel = torch.logsumexp(per_offset, -1)
    if pixel_weights is not None:  # test hook (ones): d/d pixel_weights of a gradient = every pixel's contribution to it
        per_pixel = per_pixel * pixel_weights
    return per_pixel.sum((-1, -2))


class AffineBeta:
    """
    distributions/affine_beta.py:33-49 on top of pyro.distributions.AffineBeta [third party]:
    ``Beta(size (mean-low)/(high-low), size (high-mean)/(high-low))`` mapped onto ``[low, high]``.
    """

    def __init__(self, mean, size, low, high):
        self.low = low
        self.scale = high - low
        self.concentration1 = size * (mean - low) / (high - low)
        self.concentration0 = size * (high - mean) / (high - low)
        self.base = D.Beta(self.concentration1, self.concentration0, validate_args=False)

    def log_prob(self, value):
        return self.base.log_prob((value - self.low) / self.scale) - math.log(abs(self.scale))

    def clamp(self, value):
        # pyro AffineBeta.rsample clamps into [low + eps*scale, high - eps*scale] [third party]
        eps = torch.finfo(value.dtype).eps * self.scale
        lo = torch.as_tensor(self.low + eps, dtype=value.dtype)
        hi = torch.as_tensor(self.low + self.scale - eps, dtype=value.dtype)
        return torch.min(torch.max(value, lo), hi)

    @property
    def mean(self):
        return self.low + self.scale * self.base.mean


# ------------------------------------------------------------------------------------------------
# reparameterised sampling with injectable base variates
# ------------------------------------------------------------------------------------------------
class _ReplayStdGamma(torch.autograd.Function):
    """Value = the given standard-gamma variate; d/d concentration = torch._standard_gamma_grad,
    exactly what ``Gamma.rsample`` -> ``_standard_gamma`` autograd does."""

    @staticmethod
    def forward(ctx, concentration, variate):
        ctx.save_for_backward(concentration, variate)
        return variate.clone()

    @staticmethod
    def backward(ctx, grad_out):
        concentration, variate = ctx.saved_tensors
        return grad_out * torch._standard_gamma_grad(concentration, variate), None


class _ReplayDirichlet(torch.autograd.Function):
    """Value = the given simplex variate; backward = torch/distributions/dirichlet.py
    ``_Dirichlet_backward`` (``torch._dirichlet_grad``)."""

    @staticmethod
    def forward(ctx, concentration, variate):
        ctx.save_for_backward(concentration, variate)
        return variate.clone()

    @staticmethod
    def backward(ctx, grad_out):
        concentration, x = ctx.saved_tensors
        total = concentration.sum(-1, True).expand_as(concentration)
        g = torch._dirichlet_grad(x, concentration, total)
        return g * (grad_out - (x * grad_out).sum(-1, True)), None


def rsample_gamma(concentration, rate, variate):
    """Gamma.rsample: standard gamma / rate, clamped (detached) at finfo.tiny."""
    concentration, rate = torch.broadcast_tensors(concentration, rate)
    value = _ReplayStdGamma.apply(concentration.contiguous(), variate) / rate
    value.detach().clamp_(min=torch.finfo(value.dtype).tiny)
    return value


def rsample_beta01(concentration1, concentration0, variate):
    """Beta.rsample: first component of a 2-Dirichlet rsample."""
    conc = torch.stack(torch.broadcast_tensors(concentration1, concentration0), -1)
    x = torch.stack([variate, 1 - variate], -1)
    return _ReplayDirichlet.apply(conc.contiguous(), x)[..., 0]


def rsample_affine_beta(dist: AffineBeta, variate):
    return dist.clamp(dist.low + dist.scale * rsample_beta01(dist.concentration1, dist.concentration0, variate))


# ------------------------------------------------------------------------------------------------
# data + variational parameters
# ------------------------------------------------------------------------------------------------
class OracleData:
    """Minimal view of utils/dataset.py:40-151 (CosmosDataset + OffsetData) on CPU tensors."""

    def __init__(self, images, xy, is_ontarget, mask=None, offset_samples=None, offset_weights=None,
                 dtype=torch.float64):
        self.dtype = dtype
        self.images = images.to(dtype)
        self.xy = xy.to(dtype)
        self.is_ontarget = is_ontarget.bool()
        self.mask = torch.ones_like(self.is_ontarget) if mask is None else mask.bool()
        self.offset_samples = offset_samples.to(dtype)
        self.offset_weights = offset_weights.to(dtype)
        self.Nt, self.F, self.C, self.P = images.shape[0], images.shape[1], images.shape[2], images.shape[3]

    @property
    def offset_logits(self):  # dataset.py:27-29 (probs_to_logits)
        return torch.distributions.utils.probs_to_logits(self.offset_weights)

    @property
    def offset_mean(self):  # dataset.py:31-33
        return torch.sum(self.offset_samples * self.offset_weights).item()

    @property
    def median(self):  # dataset.py:134-138
        return torch.stack([torch.median(self.images[..., c, :, :]) for c in range(self.C)])


def param_constraints(P, dtype):
    """Constraint of every variational parameter, models/cosmos.py:471-598."""
    eps = torch.finfo(dtype).eps
    half = (P + 1) / 2
    return {
        "pi_mean": constraints.simplex,
        "pi_size": constraints.positive,
        "m_probs": constraints.unit_interval,
        "proximity_loc": constraints.interval(0, (P + 1) / math.sqrt(12) - eps),
        "proximity_size": constraints.greater_than(2.0),
        "lamda_loc": constraints.positive,
        "lamda_beta": constraints.positive,
        "gain_loc": constraints.positive,
        "gain_beta": constraints.positive,
        "background_mean_loc": constraints.positive,
        "background_std_loc": constraints.positive,
        "b_loc": constraints.positive,
        "b_beta": constraints.positive,
        "h_loc": constraints.positive,
        "h_beta": constraints.positive,
        "w_mean": constraints.interval(0.75 + eps, 2.25 - eps),
        "w_size": constraints.greater_than(2.0),
        "x_mean": constraints.interval(-half + eps, half - eps),
        "y_mean": constraints.interval(-half + eps, half - eps),
        "size": constraints.greater_than(2.0),
    }


PARAM_NAMES = list(param_constraints(14, torch.float64).keys())
GLOBAL_PARAMS = ["pi_mean", "pi_size", "proximity_loc", "proximity_size", "lamda_loc", "lamda_beta",
                 "gain_loc", "gain_beta"]
LOCAL_PARAMS = [n for n in PARAM_NAMES if n not in GLOBAL_PARAMS]


def init_constrained(data: OracleData, K=2, S=1):
    """Initial constrained values, models/cosmos.py:471-598."""
    dt = data.dtype
    Q, Nt, F, C = data.C, data.Nt, data.F, data.C
    bg = (data.median - data.offset_mean)
    full = lambda shape, v: torch.full(shape, float(v), dtype=dt)
    return {
        "pi_mean": torch.ones(Q, S + 1, dtype=dt),
        "pi_size": full((Q, 1), 2),
        "m_probs": full((K, Nt, F, Q), 0.5),
        "proximity_loc": torch.tensor(0.5, dtype=dt),
        "proximity_size": torch.tensor(100.0, dtype=dt),
        "lamda_loc": full((Q,), 0.5),
        "lamda_beta": full((Q,), 100),
        "gain_loc": torch.tensor(5.0, dtype=dt),
        "gain_beta": torch.tensor(100.0, dtype=dt),
        "background_mean_loc": bg.expand(Nt, 1, C).clone(),
        "background_std_loc": torch.ones(Nt, 1, C, dtype=dt),
        "b_loc": bg.expand(Nt, F, C).clone(),
        "b_beta": torch.ones(Nt, F, C, dtype=dt),
        "h_loc": full((K, Nt, F, Q), 2000),
        "h_beta": full((K, Nt, F, Q), 0.001),
        "w_mean": full((K, Nt, F, Q), 1.5),
        "w_size": full((K, Nt, F, Q), 100),
        "x_mean": torch.zeros(K, Nt, F, Q, dtype=dt),
        "y_mean": torch.zeros(K, Nt, F, Q, dtype=dt),
        "size": full((K, Nt, F, Q), 200),
    }


def to_unconstrained(constrained: Dict[str, torch.Tensor], P, dtype):
    """What ``pyro.param(name, init, constraint=...)`` stores [third party]: transform_to(c).inv(init)."""
    cons = param_constraints(P, dtype)
    return {k: transform_to(cons[k]).inv(v).detach().clone() for k, v in constrained.items()}


def to_constrained(unconstrained: Dict[str, torch.Tensor], P, dtype):
    cons = param_constraints(P, dtype)
    return {k: transform_to(cons[k])(v) for k, v in unconstrained.items()}


# ------------------------------------------------------------------------------------------------
# guide distributions (models/cosmos.py:329-462)
# ------------------------------------------------------------------------------------------------
def _gather_local(p, ndx, fdx):
    """Vindex gathers of cosmos.py:397-462: (K,Nt,F,Q)->(K,nb,fb,Q), (Nt,F,C)->(nb,fb,C), (Nt,1,C)->(nb,1,C)."""
    n = ndx[:, None]
    f = fdx[None, :]
    out = {}
    for name in ["m_probs", "h_loc", "h_beta", "w_mean", "w_size", "x_mean", "y_mean", "size"]:
        out[name] = p[name][:, n, f, :]
    for name in ["b_loc", "b_beta"]:
        out[name] = p[name][n, f, :]
    for name in ["background_mean_loc", "background_std_loc"]:
        out[name] = p[name][ndx, :, :]
    return out


def guide_dists(p, loc, P, priors):
    """Concentration/rate (or AffineBeta) of every reparameterised guide site."""
    half = (P + 1) / 2
    return {
        "gain": (p["gain_loc"] * p["gain_beta"], p["gain_beta"]),  # cosmos.py:342-348
        "pi": p["pi_mean"] * p["pi_size"],  # :349-352
        "lamda": (p["lamda_loc"] * p["lamda_beta"], p["lamda_beta"]),  # :353-359
        "proximity": AffineBeta(p["proximity_loc"], p["proximity_size"], 0, (P + 1) / math.sqrt(12)),  # :360-368
        "background": (loc["b_loc"] * loc["b_beta"], loc["b_beta"]),  # :408-415
        "height": (loc["h_loc"] * loc["h_beta"], loc["h_beta"]),  # :428-435
        "width": AffineBeta(loc["w_mean"], loc["w_size"], priors["width_min"], priors["width_max"]),  # :436-444
        "x": AffineBeta(loc["x_mean"], loc["size"], -half, half),  # :445-453
        "y": AffineBeta(loc["y_mean"], loc["size"], -half, half),  # :454-462
    }


@torch.no_grad()
def draw_noise(unconstrained, data: OracleData, ndx, fdx, generator=None, priors=DEFAULT_PRIORS):
    """
    Base variates of one guide execution from the same ATen samplers ``rsample`` uses
    (``_standard_gamma``, ``_sample_dirichlet``), in the guide's site order.
    Standard-gamma variates for Gamma sites, (0,1) variates for Beta sites, simplex for ``pi``.
    """
    p = to_constrained(unconstrained, data.P, data.dtype)
    loc = _gather_local(p, ndx, fdx)
    g = guide_dists(p, loc, data.P, priors)
    sg = lambda c: torch._standard_gamma(c.contiguous(), generator=generator)

    def beta01(ab):
        conc = torch.stack(torch.broadcast_tensors(ab.concentration1, ab.concentration0), -1).contiguous()
        return torch._sample_dirichlet(conc, generator=generator)[..., 0]

    noise = {}
    noise["gain"] = sg(g["gain"][0])
    noise["pi"] = torch._sample_dirichlet(g["pi"].contiguous(), generator=generator)
    noise["lamda"] = sg(g["lamda"][0])
    noise["proximity"] = beta01(g["proximity"])
    noise["background"] = sg(g["background"][0])
    K = loc["h_loc"].shape[0]
    hs, ws, xs, ys = [], [], [], []
    for k in range(K):
        hs.append(sg(g["height"][0][k]))
        ws.append(beta01(AffineBeta(loc["w_mean"][k], loc["w_size"][k], priors["width_min"], priors["width_max"])))
        half = (data.P + 1) / 2
        xs.append(beta01(AffineBeta(loc["x_mean"][k], loc["size"][k], -half, half)))
        ys.append(beta01(AffineBeta(loc["y_mean"][k], loc["size"][k], -half, half)))
    noise["height"], noise["width"] = torch.stack(hs), torch.stack(ws)
    noise["x"], noise["y"] = torch.stack(xs), torch.stack(ys)
    return noise


def subsample(size, subsample_size, generator=None):
    """pyro.plate subsampling [third party]: arange if full, else randperm(size)[:subsample_size]."""
    if subsample_size >= size:
        return torch.arange(size)
    return torch.randperm(size, generator=generator)[:subsample_size]


# ------------------------------------------------------------------------------------------------
# ELBO of one step
# ------------------------------------------------------------------------------------------------
def m_configs(K, dtype):
    """Enumeration table: row i has m_k = bit k of i.  (Pyro puts m_0 at dim -4, m_1 at -5,
    cosmos.py:419-425; flattening (m_1, m_0) row-major gives this order.)"""
    return torch.tensor([[(i >> k) & 1 for k in range(K)] for i in range(2**K)], dtype=dtype)


def elbo(unconstrained, data: OracleData, ndx, fdx, noise, priors=DEFAULT_PRIORS, K=2, S=1,
         return_parts=False, plate_sizes=None, unit_weights=None, pixel_weights=None):
    """
    ELBO of one guide+model execution with the given minibatch indices and base variates.
    Follows models/cosmos.py:82-327 (model), :329-462 (guide) and the TraceEnum_ELBO
    semantics of SURVEY.md App. A.3 [third party].  Differentiable w.r.t. ``unconstrained``.
    ``plate_sizes=(Nt, F)``: sizes of the two subsampled plates when ``data`` holds only the gathered
    minibatch block of a larger dataset (tests at BASELINE sizes: the plate scales are the only place
    where the rest of the dataset enters a step, cosmos.py:194-208).  ``unit_weights`` (nb, fb, C), normally absent
    (= ones): per-unit multipliers of the frame-level terms, so that a test can take d/d unit_weights of a gradient
    and read off every unit's contribution to it (conditioning of the cross-unit sums); ``pixel_weights``
    (nb, fb, C, P, P) likewise for the pixels of the likelihood.
    """
    assert S == 1, "cosmos enumerates z in {0, 1} (probs_theta has two rows, util.py:154-173)"
    dt, P = data.dtype, data.P
    Q = C = data.C
    nb, fb = len(ndx), len(fdx)
    plate_n, plate_f = plate_sizes if plate_sizes is not None else (data.Nt, data.F)
    sN, sF = plate_n / nb, plate_f / fb
    half = (P + 1) / 2
    p = to_constrained(unconstrained, P, dt)
    loc = _gather_local(p, ndx, fdx)
    g = guide_dists(p, loc, P, priors)
    t = lambda v: torch.as_tensor(v, dtype=dt)

    # ---- global sites: guide sample, log q, log p ------------------------------------------
    gain = rsample_gamma(*g["gain"], noise["gain"])
    pi = _ReplayDirichlet.apply(g["pi"].contiguous(), noise["pi"])
    lamda = rsample_gamma(*g["lamda"], noise["lamda"])
    proximity = rsample_affine_beta(g["proximity"], noise["proximity"])
    e_global = (
        D.HalfNormal(t(priors["gain_std"])).log_prob(gain) - D.Gamma(*g["gain"]).log_prob(gain)  # cosmos.py:170
        + (D.Dirichlet(torch.ones(Q, S + 1, dtype=dt) / (S + 1)).log_prob(pi)
           - D.Dirichlet(g["pi"]).log_prob(pi)).sum()  # :171-174
        + (D.Exponential(torch.full((Q,), priors["lamda_rate"], dtype=dt)).log_prob(lamda)
           - D.Gamma(*g["lamda"]).log_prob(lamda)).sum()  # :176-181
        + D.Exponential(t(priors["proximity_rate"])).log_prob(proximity) - g["proximity"].log_prob(proximity)  # :182-184
    )
    size = torch.stack([torch.full_like(proximity, 2.0), ((P + 1) / (2 * proximity)) ** 2 - 1], -1)  # :185-191

    # ---- AOI-level sites (Delta guide => log q = 0) ----------------------------------------
    mask = data.mask[ndx].to(dt)[:, None, None]  # (nb,1,1), cosmos.py:218-219
    bm, bs = loc["background_mean_loc"], loc["background_std_loc"]  # (nb,1,C)
    e_aoi = (mask * (D.HalfNormal(t(priors["background_mean_std"])).log_prob(bm)
                     + D.HalfNormal(t(priors["background_std_std"])).log_prob(bs))).sum()  # :221-227

    # ---- frame-level continuous sites ------------------------------------------------------
    background = rsample_gamma(*g["background"], noise["background"])  # (nb,fb,C)
    e_b = D.Gamma((bm / bs) ** 2, bm / bs**2).log_prob(background) - D.Gamma(*g["background"]).log_prob(background)  # :233-239
    height = rsample_gamma(*g["height"], noise["height"])  # (K,nb,fb,C)
    width = rsample_affine_beta(g["width"], noise["width"])
    x = rsample_affine_beta(g["x"], noise["x"])
    y = rsample_affine_beta(g["y"], noise["y"])
    # masked-by-m_k terms that do not depend on (z, theta): cosmos.py:270-282 minus guide :428-462
    spot_terms = (
        D.HalfNormal(t(priors["height_std"])).log_prob(height)
        + AffineBeta(t(1.5), t(2.0), priors["width_min"], priors["width_max"]).log_prob(width)
        - D.Gamma(*g["height"]).log_prob(height) - g["width"].log_prob(width)
        - g["x"].log_prob(x) - g["y"].log_prob(y)
    )  # (K,nb,fb,C)

    # ---- enumerated part -------------------------------------------------------------------
    mcfg = m_configs(K, dt)  # (M,K)
    M = mcfg.shape[0]
    # guide log q(m_k) : Bernoulli(m_probs) (cosmos.py:419-425)
    qm = D.Bernoulli(probs=loc["m_probs"], validate_args=False)
    logq_mk = torch.stack([qm.log_prob(torch.zeros((), dtype=dt)), qm.log_prob(torch.ones((), dtype=dt))])  # (2,K,nb,fb,C)
    logq_m = sum(logq_mk[mcfg[:, k].long(), k] for k in range(K))  # (M,nb,fb,C)
    q_m = logq_m.exp()

    # model: z, theta, m_k | theta, x_k,y_k | theta  (cosmos.py:242-300)
    ont = data.is_ontarget[ndx]
    pi_exp = expand_offtarget(pi)  # (Q,S+1,2)
    pz = pi_exp[:, :, ont.long()].permute(2, 0, 1)[:, None]  # (nb,1,C,S+1)
    logp_z = D.Categorical(probs=pz, validate_args=False).logits  # (nb,1,C,Z)
    logp_theta = D.Categorical(probs=probs_theta(K, dt), validate_args=False).logits  # (2, 1+K) rows = min(z,1)
    pm = probs_m(lamda, K)  # (Q,1+K,K)
    bern = D.Bernoulli(probs=pm, validate_args=False)
    logp_mk = torch.stack([bern.log_prob(torch.zeros((), dtype=dt)), bern.log_prob(torch.ones((), dtype=dt))])  # (2,Q,1+K,K)
    xy_prior = [AffineBeta(t(0.0), size[s], -half, half) for s in range(2)]
    logp_xy = torch.stack([d.log_prob(x) + d.log_prob(y) for d in xy_prior])  # (2,K,nb,fb,C)

    joint = []
    for z in range(S + 1):
        for th in range(K + 1):
            term = logp_z[..., z].expand(nb, fb, C) + logp_theta[min(z, 1), th]  # (nb,fb,C)
            term = term[None].expand(M, nb, fb, C)
            for k in range(K):
                mk = mcfg[:, k]
                spec = int(th == k + 1)
                lm = logp_mk[mk.long(), :, th, k]  # (M,Q)
                term = term + lm[:, None, None, :] + mk[:, None, None, None] * logp_xy[spec, k]
            joint.append(term)
    T = torch.logsumexp(torch.stack(joint), 0)  # (M,nb,fb,C)

    # likelihood for every m config (cosmos.py:310-327)
    stk = lambda v: v.permute(1, 2, 3, 0)  # (K,nb,fb,C)->(nb,fb,C,K)
    target = data.xy[ndx[:, None], fdx[None, :]]  # (nb,fb,C,2)
    obs = data.images[ndx[:, None], fdx[None, :]]  # (nb,fb,C,P,P)
    L = ksmogn_log_prob(stk(height), stk(width), stk(x), stk(y), target, background, gain,
                        data.offset_samples, data.offset_logits, P, obs,
                        m=mcfg[:, None, None, None, :], pixel_weights=pixel_weights)  # (M,nb,fb,C)

    per_config = T + L - logq_m + sum(mcfg[:, k][:, None, None, None] * spot_terms[k] for k in range(K))
    uw = 1.0 if unit_weights is None else unit_weights
    e_frame = (mask * uw * (e_b + (q_m * per_config).sum(0))).sum()
    total = e_global + sN * e_aoi + sN * sF * e_frame
    if return_parts:
        parts = dict(e_global=e_global, e_aoi=e_aoi, e_frame=e_frame, T=T, L=L, q_m=q_m, logq_m=logq_m,
                     spot_terms=spot_terms, e_b=e_b, gain=gain, pi=pi, lamda=lamda, proximity=proximity,
                     background=background, height=height, width=width, x=x, y=y, per_config=per_config)
        return total, parts
    return total


def loss_and_grads(unconstrained, data, ndx, fdx, noise, **kw):
    """``TraceEnum_ELBO.loss_and_grads`` [third party]: loss = -ELBO, dense grads on every parameter."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in unconstrained.items()}
    loss = -elbo(leaves, data, ndx, fdx, noise, **kw)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.item(), grads


class OracleSVI:
    """
    ``Model.init`` + ``Model.run`` loop body (models/model.py:153-186, 210-212) with
    ``pyro.optim.Adam({"lr": lr, "betas": [0.9, 0.999]})`` [third party]: one ``torch.optim.Adam``
    per parameter tensor acting on the unconstrained value, dense over the whole tensor.
    """

    def __init__(self, data: OracleData, K=2, S=1, lr=0.005, nbatch_size=5, fbatch_size=512,
                 priors=DEFAULT_PRIORS, seed=0):
        self.data, self.K, self.S, self.priors = data, K, S, priors
        self.nbatch_size = min(nbatch_size, data.Nt)
        self.fbatch_size = min(fbatch_size, data.F)
        self.params = to_unconstrained(init_constrained(data, K, S), data.P, data.dtype)
        for v in self.params.values():
            v.requires_grad_(True)
        self.optim = {k: torch.optim.Adam([v], lr=lr, betas=(0.9, 0.999)) for k, v in self.params.items()}
        self.generator = torch.Generator().manual_seed(seed)
        self.iter = 0

    def constrained(self):
        with torch.no_grad():
            return to_constrained(self.params, self.data.P, self.data.dtype)

    def step(self, ndx=None, fdx=None, noise=None):
        if ndx is None:
            ndx = subsample(self.data.Nt, self.nbatch_size, self.generator)
        if fdx is None:
            fdx = subsample(self.data.F, self.fbatch_size, self.generator)
        if noise is None:
            noise = draw_noise(self.params, self.data, ndx, fdx, self.generator, self.priors)
        for v in self.params.values():
            v.grad = None
        loss = -elbo(self.params, self.data, ndx, fdx, noise, priors=self.priors, K=self.K, S=self.S)
        loss.backward()
        self.last_grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
                           for k, v in self.params.items()}
        for k, opt in self.optim.items():
            if self.params[k].grad is None:
                self.params[k].grad = torch.zeros_like(self.params[k])
            opt.step()
        self.iter += 1
        return loss.item()


# ------------------------------------------------------------------------------------------------
# posterior of z / theta (models/cosmos.py:609-672), "next" row N1
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def compute_probs(unconstrained, data: OracleData, ndx, fdx, noises, priors=DEFAULT_PRIORS, K=2, S=1):
    """
    z_probs (nb,fb,Q,1+S) and theta_probs (K,nb,fb,Q) for a minibatch, averaged over the given
    list of guide draws (the reference uses 50 particles, cosmos.py:630-667): for every draw,
    log p(z, theta, m, x, y) from the model with the data site hidden (:634-648), normalised over
    (z, theta) (:650), weighted by the guide's q(m) (:651-657), marginalised (:659-667).
    """
    dt, P = data.dtype, data.P
    Q = C = data.C
    nb, fb = len(ndx), len(fdx)
    half = (P + 1) / 2
    t = lambda v: torch.as_tensor(v, dtype=dt)
    p = to_constrained(unconstrained, P, dt)
    loc = _gather_local(p, ndx, fdx)
    g = guide_dists(p, loc, P, priors)
    mcfg = m_configs(K, dt)
    M = mcfg.shape[0]
    qm = D.Bernoulli(probs=loc["m_probs"], validate_args=False)
    logq_mk = torch.stack([qm.log_prob(torch.zeros((), dtype=dt)), qm.log_prob(torch.ones((), dtype=dt))])
    logq_m = sum(logq_mk[mcfg[:, k].long(), k] for k in range(K))  # (M,nb,fb,C)
    ont = data.is_ontarget[ndx]
    z_acc = torch.zeros(nb, fb, C, S + 1, dtype=dt)
    th_acc = torch.zeros(K, nb, fb, C, dtype=dt)
    for noise in noises:
        pi = noise["pi"].to(dt)
        lamda = rsample_gamma(*g["lamda"], noise["lamda"])
        proximity = rsample_affine_beta(g["proximity"], noise["proximity"])
        size = torch.stack([torch.full_like(proximity, 2.0), ((P + 1) / (2 * proximity)) ** 2 - 1], -1)
        x = rsample_affine_beta(g["x"], noise["x"])
        y = rsample_affine_beta(g["y"], noise["y"])
        pz = expand_offtarget(pi)[:, :, ont.long()].permute(2, 0, 1)[:, None]
        logp_z = D.Categorical(probs=pz, validate_args=False).logits
        logp_theta = D.Categorical(probs=probs_theta(K, dt), validate_args=False).logits
        bern = D.Bernoulli(probs=probs_m(lamda, K), validate_args=False)
        logp_mk = torch.stack([bern.log_prob(torch.zeros((), dtype=dt)), bern.log_prob(torch.ones((), dtype=dt))])
        xy_prior = [AffineBeta(t(0.0), size[s], -half, half) for s in range(2)]
        logp_xy = torch.stack([d.log_prob(x) + d.log_prob(y) for d in xy_prior])
        joint = torch.empty(S + 1, K + 1, M, nb, fb, C, dtype=dt)
        for z in range(S + 1):
            for th in range(K + 1):
                term = (logp_z[..., z].expand(nb, fb, C) + logp_theta[min(z, 1), th])[None].expand(M, nb, fb, C)
                for k in range(K):
                    mk = mcfg[:, k]
                    term = term + logp_mk[mk.long(), :, th, k][:, None, None, :] + mk[:, None, None, None] * logp_xy[int(th == k + 1), k]
                joint[z, th] = term
        post = joint - torch.logsumexp(joint.reshape(-1, M, nb, fb, C), 0)  # normalise over (z, theta)
        result = torch.logsumexp(post + logq_m, 2)  # average over m -> (Z, TH, nb, fb, C)
        z_acc += torch.logsumexp(result, 1).exp().permute(1, 2, 3, 0)
        th_acc += torch.logsumexp(result, 0).exp()[1:]
    return z_acc / len(noises), th_acc / len(noises)
