"""
GPU parity of one full cosmos SVI step (all kernels through the C ABI) against the fp64 oracle in
replay mode: explicit minibatch indices + injected base variates ("RNG parity").
Tolerances: fp64 kernels -> rounding level; fp32 kernels -> north-star 1e-5 (loss and gradients,
gradients relative to the largest entry of each tensor).
"""

import pytest
import torch

from oracle import cosmos_oracle as O
from tests.step_helpers import compare_grads, flat_inputs, make_problem

pytestmark = pytest.mark.gpu


def make_engine(ds, data, params, nb, fb, dtype, **kw):
    from tapqir_b200.models.engine import CosmosEngine

    store = ds.device_store("cuda", dtype)
    eng = CosmosEngine(store, data.Nt, data.F, data.C, data.P, O.DEFAULT_PRIORS, dtype=dtype, nbatch_size=nb,
                       fbatch_size=fb, **kw)
    eng.load_unconstrained(params)
    return eng


def replay_args(eng, data, params, ndx, fdx, noise, dtype):
    _, gl, _, _, lnoise, gnoise = flat_inputs(data, params, noise, dtype)
    # explicit indices even for a full batch: the oracle's minibatch order is a permutation
    n, f = ndx.to(torch.int32).cuda(), fdx.to(torch.int32).cuda()
    return dict(ndx=n, fdx=f, local_noise=lnoise.cuda(), global_noise=gnoise.cuda())


CONFIGS = [
    dict(N=4, F=6, C=1, nb=3, fb=4, seed=0),
    dict(N=4, F=5, C=2, nb=4, fb=3, seed=1),
    dict(N=3, F=4, C=1, nb=3, fb=4, seed=2, offsets="hist"),
    dict(N=5, F=100, C=1, nb=5, fb=100, seed=3, perturb=False),   # BASELINE config 1 shape, full batch
    dict(N=6, F=70, C=1, nb=4, fb=33, seed=4),
]


@pytest.mark.parametrize("cfg", CONFIGS)
# fp64 gradients: 1e-8 (the paired Rice expansion of the Beta reparameterisation gradient is
# ill-conditioned near x = mean, evaluation order moves it at the 1e-9 level; see test_hostcheck_math)
@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)])
def test_step_loss_and_grads_match_oracle(cfg, dtype, ltol, gtol):
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    if dtype == torch.float32:  # both sides see the same fp32-rounded parameters and variates
        params = {k: v.float().double() for k, v in params.items()}
        noise = {k: v.float().double() for k, v in noise.items()}
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    eng = make_engine(ds, data, params, cfg["nb"], cfg["fb"], dtype)
    loss = eng.step(update=False, **replay_args(eng, data, params, ndx, fdx, noise, dtype)).item()
    assert abs(loss - ref_loss) <= ltol * abs(ref_loss)
    bad = compare_grads(eng.named_grads(), ref_grads, gtol)
    assert not bad, bad


def test_masked_aoi_contributes_nothing():
    ds, data, params, ndx, fdx, noise = make_problem(N=4, F=5, nb=4, fb=5, seed=4)
    ds.mask[1] = False
    data.mask[1] = False
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    eng = make_engine(ds, data, params, 4, 5, torch.float64)
    loss = eng.step(update=False, **replay_args(eng, data, params, ndx, fdx, noise, torch.float64)).item()
    assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
    assert not compare_grads(eng.named_grads(), ref_grads, 1e-9)
    assert eng.named_grads()["b_loc"][1].abs().max().item() == 0


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-5)])
def test_adam_trajectory_matches_oracle(dtype, tol):
    """Several replayed steps with dense Adam (model.py:168-171): the parameters stay on the
    oracle's trajectory.  Entries outside the minibatch move too (momentum), SURVEY fact 5."""
    cfg = dict(N=4, F=8, C=1, nb=2, fb=5, seed=7)
    ds, data, params, _, _, _ = make_problem(**cfg)
    if dtype == torch.float32:
        params = {k: v.float().double() for k, v in params.items()}
    svi = O.OracleSVI(data, nbatch_size=2, fbatch_size=5)
    for k in svi.params:
        svi.params[k].data.copy_(params[k])
    eng = make_engine(ds, data, params, 2, 5, dtype)
    g = torch.Generator().manual_seed(11)
    for it in range(6):
        ndx = torch.randperm(data.Nt, generator=g)[:2]
        fdx = torch.randperm(data.F, generator=g)[:5]
        cur = {k: v.detach().clone() for k, v in svi.params.items()}
        noise = O.draw_noise(cur, data, ndx, fdx, g)
        if dtype == torch.float32:
            noise = {k: v.float().double() for k, v in noise.items()}
        ref_loss = svi.step(ndx, fdx, noise)
        loss = eng.step(**replay_args(eng, data, cur, ndx, fdx, noise, dtype)).item()
        assert abs(loss - ref_loss) <= max(tol, 1e-6) * abs(ref_loss), it
    assert eng.iteration == 6
    ours = eng.named_unconstrained()
    for k, v in svi.params.items():
        err = (ours[k].double().cpu().reshape(v.shape) - v.detach()).abs().max().item()
        assert err <= tol * max(1.0, v.detach().abs().max().item()), (k, err)


def test_device_rng_step_runs_and_is_reproducible():
    """Production mode: indices and variates drawn on the device (Philox keyed by seed and step).
    Same seed -> bit-identical trajectory; different seed -> different samples; loss decreases."""
    cfg = dict(N=6, F=40, C=1, nb=4, fb=16, seed=5, perturb=False)
    ds, data, params, _, _, _ = make_problem(**cfg)
    runs = []
    for seed in (3, 3, 4):
        eng = make_engine(ds, data, params, 4, 16, torch.float32, seed=seed)
        losses = [eng.step().item() for _ in range(40)]
        runs.append((losses, eng.lparams.clone()))
    assert runs[0][0] == runs[1][0] and torch.equal(runs[0][1], runs[1][1])
    assert runs[0][0] != runs[2][0]
    assert all(torch.isfinite(torch.tensor(r[0])).all() for r in runs)
    first, last = sum(runs[0][0][:10]) / 10, sum(runs[0][0][-10:]) / 10
    assert last < first


def test_device_sampler_matches_guide_distribution():
    """In-kernel Marsaglia-Tsang / Beta sampling: moments of the guide samples at the initial
    parameters (b ~ Gamma(b_loc b_beta, b_beta), x ~ AffineBeta(0, 200, -7.5, 7.5), ...)."""
    ds, data, params, _, _, _ = make_problem(N=8, F=200, C=1, nb=8, fb=200, seed=6, perturb=False)
    eng = make_engine(ds, data, params, 8, 200, torch.float32, seed=1)
    eng.step(update=False)
    S = eng.samples.double().cpu()
    c = O.to_constrained(params, data.P, data.dtype)
    b_loc, b_beta = c["b_loc"][0, 0, 0].item(), c["b_beta"][0, 0, 0].item()
    n = S.shape[1]
    assert abs(S[0].mean().item() - b_loc) < 5 * (b_loc / b_beta / n) ** 0.5
    assert abs(S[0].var().item() - b_loc / b_beta) < 0.15 * b_loc / b_beta
    # height ~ Gamma(2000 * 0.001, 0.001): mean 2000, var 2e6
    assert abs(S[1].mean().item() - 2000) < 5 * (2e6 / n) ** 0.5
    # width ~ AffineBeta(1.5, 100, .75, 2.25): var = scale^2 * m(1-m)/(size+1)
    assert abs(S[3].mean().item() - 1.5) < 5 * (1.5**2 * 0.25 / 101 / n) ** 0.5
    assert abs(S[3].var().item() - 1.5**2 * 0.25 / 101) < 0.15 * 1.5**2 * 0.25 / 101
    # x ~ AffineBeta(0, 200, -7.5, 7.5)
    assert abs(S[5].mean().item()) < 5 * (15**2 * 0.25 / 201 / n) ** 0.5
    assert abs(S[5].var().item() - 15**2 * 0.25 / 201) < 0.15 * 15**2 * 0.25 / 201
