"""
Synthetic cosmos data with the reference's sampling recipe (tapqir/utils/simulate.py:12-138).

The reference drives ``pyro.infer.Predictive`` over the *model* with the global variables fixed to
``params``; written out, that is (models/cosmos.py:242-327 with simulate.py:39-58,92-105):

* the first ``N // 2`` AOIs are on-target, target location is the patch centre ``(P-1)/2``;
* ``z ~ Bernoulli(pi)`` on-target, ``0`` off-target; ``theta = 0`` if ``z = 0`` else uniform on 1..K;
* ``m_k ~ Bernoulli(probs_m(lamda, K)[theta, k])`` (distributions/util.py:94-130);
* ``height, width, background`` fixed; ``x_k, y_k ~ AffineBeta(0, size, -(P+1)/2, (P+1)/2)`` with
  ``size = ((P+1)/(2 proximity))^2 - 1`` for the target-specific spot and ``2`` (uniform) otherwise;
* pixels ``floor(Gamma(image/gain, 1/gain) + offset)`` with three identical offset bins.

Random numbers come from this module's own seeded ``torch.Generator`` (the reference's Pyro RNG
stream cannot be reproduced), so datasets agree with the reference in distribution, not bit-wise.
Used for benchmark inputs and tests; not on the timed hot path.
"""

import math

import numpy as np
import torch

from tapqir_b200.distributions.util import probs_m
from tapqir_b200.utils.dataset import CosmosDataset

TEST_PARAMS = {  # test/test_tapqir.py:25-40
    "pi": 0.15,
    "width": 1.4,
    "gain": 7.0,
    "lamda": 0.15,
    "proximity": 0.2,
    "offset": 90.0,
    "height": 3000.0,
    "background": 150.0,
}


def simulate(*args, **kwargs) -> CosmosDataset:
    """
    ``simulate(model, N, F, C=1, P=14, seed=0, params={})`` -- the reference's positional signature
    (tapqir/utils/simulate.py:12-20; ``model`` may be a model instance, its registry name or None: the recipe below IS
    the cosmos / cosmos+hmm generative model, selected by ``params``) -- or, without the leading model,
    ``simulate(N, F, ...)``.  See :func:`simulate_dataset` for the remaining keywords.
    """
    if args and not isinstance(args[0], int):
        args = args[1:]
    elif "model" in kwargs:
        kwargs.pop("model")
    if len(args) > 2:      # reference order of the optional positionals: C, P, seed, params
        for name, value in zip(("C", "P", "seed", "params"), args[2:]):
            kwargs[name] = value
        args = args[:2]
    return simulate_dataset(*args, **kwargs)


def simulate_dataset(N: int, F: int, C: int = 1, P: int = 14, K: int = 2, seed: int = 0, params: dict = None,
                     device="cpu", aoi_chunk: int = 64, offset_samples=None, offset_weights=None) -> CosmosDataset:
    """
    Draw a dataset of ``N`` AOIs (half on-target) x ``F`` frames x ``C`` channels of PxP patches.

    ``params`` with ``"kon"`` and ``"koff"`` selects the kinetic (hidden-Markov) recipe for ``z`` instead of the
    time-independent ``"pi"`` one.  ``offset_samples / offset_weights`` default to the reference's three equal bins at
    ``params["offset"]`` (simulate.py:92,103); pass a histogram for the secondary realism runs.
    Images are returned on the CPU as float32 with integer values (simulate.py:122 floors them).
    """
    prm = dict(TEST_PARAMS)
    prm.update(params or {})
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=dev)

    if offset_samples is None:
        offset_samples = torch.full((3,), float(prm["offset"]), dtype=torch.float64)
        offset_weights = torch.ones(3, dtype=torch.float64) / 3
    off_s = offset_samples.to(**f64)
    off_cdf = torch.cumsum(offset_weights.to(**f64), 0)
    off_cdf[-1] = 1.0

    pm = probs_m(torch.full((C,), float(prm["lamda"]), dtype=torch.float64), K).to(dev)  # (C,1+K,K)
    size_spec = ((P + 1) / (2 * prm["proximity"])) ** 2 - 1
    half = (P + 1) / 2
    gain = float(prm["gain"])

    is_ontarget = torch.zeros(N, dtype=torch.bool)
    is_ontarget[: N // 2] = True
    images = torch.empty(N, F, C, P, P, dtype=torch.float32)
    z_all = torch.zeros(N, F, C, dtype=torch.int64)
    grid = torch.arange(P, **f64)
    centre = (P - 1) / 2

    for lo in range(0, N, aoi_chunk):
        hi = min(lo + aoi_chunk, N)
        n = hi - lo
        ont = is_ontarget[lo:hi].to(dev)[:, None, None]
        u = lambda *s: torch.rand(*s, generator=gen, **f64)
        if "kon" in prm and "koff" in prm:
            # kinetic simulation (simulate.py:66-90): two-state Markov chain with init = stationary distribution
            # [koff, kon] / (kon + koff) and trans = [[1 - kon, kon], [koff, 1 - koff]]
            kon, koff = float(prm["kon"]), float(prm["koff"])
            draws = u(n, F, C)
            z = torch.zeros(n, F, C, dtype=torch.bool, device=dev)
            z[:, 0] = draws[:, 0] < kon / (kon + koff)
            for f in range(1, F):
                p_on = torch.where(z[:, f - 1], torch.full((), 1.0 - koff, **f64), torch.full((), kon, **f64))
                z[:, f] = draws[:, f] < p_on
            z = z & ont
        else:
            z = (u(n, F, C) < prm["pi"]) & ont
        theta = torch.where(z, 1 + torch.floor(u(n, F, C) * K).clamp(max=K - 1).long(), torch.zeros((), dtype=torch.long, device=dev))
        image = torch.full((n, F, C, P, P), float(prm["background"]), **f64)
        cdx = torch.arange(C, device=dev)[None, None, :].expand(n, F, C)
        for k in range(K):
            m_k = u(n, F, C) < pm[cdx, theta, k]
            conc = torch.where(theta == k + 1, torch.full((), size_spec / 2, **f64), torch.ones((), **f64))
            xy = []
            for _ in range(2):  # Beta(c, c) = g1 / (g1 + g2)
                g1 = torch._standard_gamma(conc.expand(n, F, C).contiguous(), generator=gen)
                g2 = torch._standard_gamma(conc.expand(n, F, C).contiguous(), generator=gen)
                xy.append(-half + 2 * half * g1 / (g1 + g2))
            w2 = prm["width"] ** 2
            gx = torch.exp(-((grid - (xy[0] + centre)[..., None]) ** 2) / (2 * w2))  # (n,F,C,P) along x
            gy = torch.exp(-((grid - (xy[1] + centre)[..., None]) ** 2) / (2 * w2))
            amp = m_k.to(torch.float64) * prm["height"] / (2 * math.pi * w2)
            image = image + amp[..., None, None] * gy[..., :, None] * gx[..., None, :]
        noise = torch._standard_gamma((image / gain).contiguous(), generator=gen) * gain
        odx = torch.searchsorted(off_cdf, u(n, F, C, P, P).reshape(-1)).clamp(max=len(off_s) - 1)
        pixels = torch.floor(noise + off_s[odx].reshape(n, F, C, P, P))
        images[lo:hi] = pixels.to(torch.float32).cpu()
        z_all[lo:hi] = z.long().cpu()

    n_on = N // 2
    labels = np.zeros((n_on, F, C), dtype=[("aoi", int), ("frame", int), ("z", int)])
    labels["aoi"] = np.arange(n_on).reshape(-1, 1, 1)
    labels["frame"] = np.arange(F).reshape(-1, 1)
    labels["z"] = z_all[:n_on].numpy()
    return CosmosDataset(
        images,
        torch.full((N, F, C, 2), centre, dtype=torch.float64),
        is_ontarget,
        labels=labels,
        offset_samples=offset_samples.clone().double(),  # the reference simulates under default dtype double
        offset_weights=offset_weights.clone().double(),
        device="cpu",
        name=f"simulated_N{N}_F{F}_C{C}_seed{seed}",
    )
