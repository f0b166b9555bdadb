"""
tests/golden/minipyro.py stands in for Pyro when tests/golden/make_golden_step.py runs the reference's model code; the
goldens are only as good as that stand-in.  These tests check it, on CPU, against answers known independently: the
indexing identities of ``Vindex``, closed-form ELBOs of small discrete models (enumeration in the model, in the guide,
in both; plates with subsampling; masks), the clamp of ``AffineBeta.rsample``, and the SVI / Adam loop.
"""

import math
import sys
from pathlib import Path

import pytest
import scipy.stats as st
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import minipyro as mp  # noqa: E402  (used directly; ``install()`` is NOT called, sys.modules stays untouched)


@pytest.fixture(autouse=True)
def clean():
    mp.clear_param_store()
    assert not mp._STACK
    yield
    assert not mp._STACK


def test_vindex_broadcasts_indices_on_the_left_and_keeps_slices_on_the_right():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 4, 3, generator=g)
    i, j = torch.tensor([[0], [3], [4]]), torch.tensor([1, 2])                 # (3,1) and (2,)
    out = mp.Vindex(x)[i, :, j]                                                # -> (3, 2, 4)
    assert out.shape == (3, 2, 4)
    for a in range(3):
        for b in range(2):
            assert torch.equal(out[a, b], x[i[a, 0], :, j[b]])
    # leading Ellipsis: batch dimensions of the tensor itself line up with the indices' batch dimensions
    y = torch.randn(3, 1, 6, 2, generator=g)                                   # batch (3,1), event (6,2)
    k = torch.tensor([5, 0])                                                   # (2,)
    out = mp.Vindex(y)[..., k, :]
    assert out.shape == (3, 2, 2)
    for a in range(3):
        for b in range(2):
            assert torch.equal(out[a, b], y[a, 0, k[b], :])
    # plain cases fall through to ordinary indexing
    assert torch.equal(mp.Vindex(x)[2], x[2]) and torch.equal(mp.Vindex(x)[1, 2, 0], x[1, 2, 0])


def _elbo(model, guide, nesting=1):
    return mp.TraceEnum_ELBO(max_plate_nesting=nesting).differentiable_elbo(model, guide)


def test_model_side_enumeration_is_the_log_marginal():
    p, mu, x = torch.tensor([0.2, 0.5, 0.3]), torch.tensor([-1.0, 0.5, 2.0]), torch.tensor(0.7)

    def model():
        z = mp.sample("z", mp.Categorical(p), infer={"enumerate": "parallel"})
        mp.sample("x", mp._wrap(torch.distributions.Normal)(mu[z], 1.0), obs=x)

    ref = math.log(sum(p[k].item() * st.norm.pdf(x.item(), mu[k].item(), 1.0) for k in range(3)))
    assert abs(_elbo(model, lambda: None).item() - ref) < 1e-6


def test_guide_side_enumeration_is_the_exact_expectation_with_plate_scale_and_gradient():
    """A plate of 6 with 3 subsampled; guide enumerates m ~ Bernoulli(q_n); model m ~ Bernoulli(0.3), x_n ~ N(m, 1)."""
    xs = torch.tensor([0.1, 1.2, -0.4, 0.9, 2.0, 0.3])
    logit = torch.tensor([0.3, -0.2, 0.8, 0.0, 1.5, -1.0], requires_grad=True)
    idx = torch.tensor([4, 0, 3])

    def guide():
        with mp.plate("n", 6, subsample=idx, dim=-1) as i:
            mp.sample("m", mp.Bernoulli(torch.sigmoid(logit[i])), infer={"enumerate": "parallel"})

    def model():
        with mp.plate("n", 6, subsample=idx, dim=-1) as i:
            m = mp.sample("m", mp.Bernoulli(torch.tensor(0.3)))
            mp.sample("x", mp._wrap(torch.distributions.Normal)(m, 1.0), obs=xs[i])

    elbo = _elbo(model, guide)
    q = torch.sigmoid(logit[idx])
    ref = 0.0
    for n in range(3):
        for m, qm in ((0.0, 1 - q[n]), (1.0, q[n])):
            lp = math.log(0.3 if m else 0.7) + st.norm.logpdf(xs[idx[n]].item(), m, 1.0)
            ref = ref + qm * (lp - torch.log(qm))
    ref = 2.0 * ref                                                            # plate scale 6 / 3
    assert abs(elbo.item() - ref.item()) < 1e-6
    (g,) = torch.autograd.grad(elbo, logit, retain_graph=True)
    (gr,) = torch.autograd.grad(ref, logit)
    assert torch.allclose(g, gr, atol=1e-6) and g[[1, 2, 5]].abs().max().item() == 0.0


def test_both_sides_enumerated_and_a_masked_dependent_site():
    """The structure of cosmos in miniature: guide enumerates m, model enumerates z; a site masked by m > 0."""
    q, pz = torch.tensor(0.35), torch.tensor([0.6, 0.4])
    pm = torch.tensor([0.1, 0.8])                       # p(m = 1 | z)
    x, h = torch.tensor(0.4), torch.tensor(1.7)
    Normal = mp._wrap(torch.distributions.Normal)

    def guide():
        m = mp.sample("m", mp.Bernoulli(q), infer={"enumerate": "parallel"})
        with mp.mask(mask=m > 0):
            mp.sample("h", mp.Delta(h))

    def model():
        z = mp.sample("z", mp.Categorical(pz), infer={"enumerate": "parallel"})
        m = mp.sample("m", mp.Bernoulli(mp.Vindex(pm)[z]))
        with mp.mask(mask=m > 0):
            hh = mp.sample("h", Normal(1.0, 2.0))
        mp.sample("x", Normal(m * hh, 1.0), obs=x)

    ref = 0.0
    for m, qm in ((0, 1 - q.item()), (1, q.item())):
        inner = sum(pz[z].item() * (pm[z].item() if m else 1 - pm[z].item()) for z in range(2))
        v = math.log(inner) + st.norm.logpdf(x.item(), m * h.item(), 1.0) - math.log(qm)
        if m:
            v += st.norm.logpdf(h.item(), 1.0, 2.0)
        ref += qm * v
    assert abs(_elbo(model, guide, nesting=0).item() - ref) < 1e-6


def test_masked_enumerated_sites_leave_the_log_cardinality():
    """What the goldens' masked-AOI constant rests on: a mask zeroes log-probabilities, enumeration still sums over the
    values, so a fully masked unit contributes (number of guide configurations) x log(number of model states)."""
    def guide():
        with mp.mask(mask=torch.tensor(False)):
            mp.sample("m", mp.Bernoulli(torch.tensor(0.3)), infer={"enumerate": "parallel"})

    def model():
        with mp.mask(mask=torch.tensor(False)):
            z = mp.sample("z", mp.Categorical(torch.tensor([0.2, 0.3, 0.5])), infer={"enumerate": "parallel"})
            mp.sample("m", mp.Bernoulli(torch.tensor([0.1, 0.5, 0.9])[z]))

    assert abs(_elbo(model, guide, nesting=0).item() - 2 * math.log(3)) < 1e-6


def test_affine_beta_density_and_clamped_rsample():
    torch.set_default_dtype(torch.float64)               # as the reference does (model.py:82-91); python-number bounds follow it
    try:
        _affine_beta_checks()
    finally:
        torch.set_default_dtype(torch.float32)


def _affine_beta_checks():
    d = mp.AffineBeta(torch.tensor(3.0, dtype=torch.float64), torch.tensor(5.0, dtype=torch.float64), -7.5, 15.0)
    v = torch.tensor(1.25, dtype=torch.float64)
    ref = st.beta.logpdf((1.25 + 7.5) / 15.0, 3.0, 5.0) - math.log(15.0)
    assert abs(d.log_prob(v).item() - ref) < 1e-12
    assert abs(d.mean.item() - (-7.5 + 15.0 * 3 / 8)) < 1e-12
    tiny = mp.AffineBeta(torch.tensor(1e-3, dtype=torch.float64), torch.tensor(1e-3, dtype=torch.float64), 0.0, 2.0)
    torch.manual_seed(0)
    s = tiny.rsample((2000,))                           # mass piles up at both ends: the clamp keeps the support open
    eps = torch.finfo(torch.float64).eps * 2.0
    assert s.min().item() >= eps and s.max().item() <= 2.0 - eps and (s.min().item() == eps or s.max().item() == 2.0 - eps)


def test_param_store_and_svi_step_are_adam_on_the_unconstrained_value():
    from torch.distributions import constraints

    data = torch.tensor([1.5, 2.5, 2.0])
    Normal = mp._wrap(torch.distributions.Normal)

    def model():
        s = mp.param("scale", lambda: torch.tensor(2.0), constraint=constraints.positive)
        with mp.plate("d", 3, dim=-1):
            mp.sample("x", Normal(2.0, s), obs=data)

    svi = mp.SVI(model, lambda: None, mp.Adam({"lr": 0.1, "betas": [0.9, 0.999]}), mp.TraceEnum_ELBO(max_plate_nesting=1))
    u0 = math.log(2.0)                                   # what the store holds: the unconstrained value
    s = torch.tensor(u0, requires_grad=True)
    loss0 = -Normal(2.0, s.exp()).log_prob(data).sum()
    (g,) = torch.autograd.grad(loss0, s)
    loss = svi.step()
    assert abs(loss - loss0.item()) < 1e-6
    u1 = mp.get_param_store().unconstrained()["scale"].item()
    assert abs(u1 - (u0 - 0.1 * math.copysign(1.0, g.item()))) < 1e-6      # first Adam step: lr * sign(gradient)
    assert abs(dict(mp.get_param_store().items())["scale"].item() - math.exp(u1)) < 1e-6


def test_predictive_fixes_given_sites_unconditions_observations_and_samples_the_rest():
    Normal = mp._wrap(torch.distributions.Normal)

    def model():
        a = mp.sample("a", Normal(0.0, 1.0))
        with mp.plate("n", 1000, dim=-1):
            return mp.sample("x", Normal(a, 0.1), obs=torch.zeros(1000))

    torch.manual_seed(0)
    out = mp.Predictive(mp.uncondition(model), posterior_samples={"a": torch.tensor([5.0, -3.0])})()
    assert set(out) == {"x"} and out["x"].shape == (2, 1000)                 # 'a' was given; 'x' is sampled, not its obs
    assert abs(out["x"][0].mean().item() - 5.0) < 0.02 and abs(out["x"][1].mean().item() + 3.0) < 0.02
    assert abs(out["x"][0].std().item() - 0.1) < 0.01
