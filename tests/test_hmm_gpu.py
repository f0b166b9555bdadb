"""
GPU parity of one SVI step of the hmm variant (BASELINE config 5; reference tapqir/models/hmm.py) against
oracle/hmm_oracle.py in replay mode: loss and every gradient, fp64 kernels at rounding level, fp32 kernels at the
north-star tolerance (1e-5 of each tensor's largest entry).
"""

import pytest
import torch

from oracle import cosmos_oracle as O
from oracle import hmm_oracle as H
from tapqir_b200.models import layout as L
from tapqir_b200.utils.simulate import simulate
from tests.step_helpers import compare_grads

pytestmark = pytest.mark.gpu


def make_problem(N, F, C, nb, seed, perturb=0.3):
    ds = simulate(N, F, C=C, seed=seed)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    g = torch.Generator().manual_seed(seed + 100)
    params = H.to_unconstrained(H.init_constrained(data), data.P, data.dtype)
    for v in params.values():
        v.add_(perturb * torch.randn(v.shape, generator=g, dtype=v.dtype))
    ndx = torch.randperm(N, generator=g)[:nb]
    return ds, data, params, ndx, g


def make_engine(ds, data, params, nb, dtype):
    from tapqir_b200.models.hmm_engine import HmmEngine

    store = ds.device_store("cuda", dtype)
    eng = HmmEngine(store, data.Nt, data.F, data.C, data.P, O.DEFAULT_PRIORS, dtype=dtype, nbatch_size=nb)
    eng.load_unconstrained(params)
    return eng


@pytest.mark.parametrize("cfg", [dict(N=4, F=6, C=1, nb=3, seed=0), dict(N=3, F=9, C=2, nb=3, seed=1),
                                 dict(N=5, F=40, C=1, nb=4, seed=2)])
@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)])
def test_hmm_step_loss_and_grads_match_oracle(cfg, dtype, ltol, gtol):
    ds, data, params, ndx, g = make_problem(**cfg)
    if dtype == torch.float32:
        params = {k: v.float().double() for k, v in params.items()}
    noise = H.draw_noise(params, data, ndx, g)
    if dtype == torch.float32:
        noise = {k: v.float().double() for k, v in noise.items()}
    ref_loss, ref_grads = H.loss_and_grads(params, data, ndx, noise)
    eng = make_engine(ds, data, params, cfg["nb"], dtype)
    lnoise = L.pack_local_noise(noise, dtype, "cuda")
    gnoise = eng.gl.pack_noise(noise).cuda()
    loss = eng.step(update=False, ndx=ndx.to(torch.int32).cuda(), local_noise=lnoise, global_noise=gnoise).item()
    assert abs(loss - ref_loss) <= ltol * abs(ref_loss)
    bad = compare_grads(eng.named_grads(), ref_grads, gtol)
    assert not bad, bad


def test_hmm_z_probs_match_oracle_and_default_step_runs():
    ds, data, params, ndx, g = make_problem(N=4, F=12, C=1, nb=4, seed=5)
    eng = make_engine(ds, data, params, 4, torch.float32)
    zp = eng.z_probs().cpu().double()
    ref = H.z_probs({k: v.float().double() for k, v in params.items()}, data)
    assert (zp - ref).abs().max().item() < 1e-6
    # device-drawn variates + Adam, CUDA-graph replay from the third call on
    losses = [eng.step().item() for _ in range(5)]
    assert all(torch.isfinite(torch.tensor(losses)))
    assert eng.iteration == 5


def test_hmm_model_api_fit_checkpoint_resume(tmp_path):
    """`tapqir fit --model cosmos+hmm` surface: registry key, parameter names / shapes (hmm.py:419-467), posterior
    properties, checkpoint written and resumed."""
    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save

    data = simulate(3, 8, C=1, P=14, seed=0)
    save(data, tmp_path)
    model = models["cosmos+hmm"](device="cuda", dtype="float")
    model.load(tmp_path)
    model.init(lr=0.005, nbatch_size=3)
    model.run(1, progress_bar=lambda it: it)   # the checkpoint of iteration 0 holds the parameters after this step
    assert model.iter == 1 and model.iter_loss == model.iter_loss
    ckpt = torch.load(tmp_path / ".tapqir" / "cosmos+hmm_model.tpqr", weights_only=False)
    p = ckpt["params"]["params"]
    assert p["m_probs"].shape == (2, 2, 3, 8, 1) and p["z_trans"].shape == (3, 8, 1, 2, 2)
    assert p["init_mean"].shape == (1, 2) and p["trans_mean"].shape == (1, 2, 2) and p["trans_size"].shape == (1, 2, 1)
    assert "pi_mean" not in p and set(ckpt["optimizer"]) == set(p)
    zp = model.z_probs
    assert zp.shape == (3, 8, 1, 2) and torch.allclose(zp.sum(-1), torch.ones(3, 8, 1), atol=1e-6)
    assert model.m_probs.shape == (2, 3, 8, 1) and model.z_map.shape == (3, 8, 1)
    zs = model.z_sample(64)
    assert zs.shape == (64, data.N, 8, 1) and abs(zs.float().mean().item() - zp[: data.N, :, :, 1].mean().item()) < 0.15
    th = model.theta_probs
    assert th.shape == (2, 3, 8, 1) and bool((th >= 0).all()) and bool((th.sum(0) <= 1 + 1e-5).all())
    again = models["cosmos+hmm"](device="cuda", dtype="float")
    again.load(tmp_path)
    again.init(lr=0.005, nbatch_size=3)
    assert again.iter == ckpt["iter"]
    a, b = model.engine.named_unconstrained(), again.engine.named_unconstrained()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(model.engine.lm, again.engine.lm) and torch.equal(model.engine.gv, again.engine.gv)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_hmm_theta_probs_match_oracle(dtype, tol):
    """theta_probs given z_MAP (hmm.py:541-625) for 3 replayed guide particles."""
    ds, data, params, _, g = make_problem(N=4, F=7, C=1, nb=4, seed=8)
    if dtype == torch.float32:
        params = {k: v.float().double() for k, v in params.items()}
    ndx = torch.arange(3)
    noises = [H.draw_noise(params, data, ndx, g) for _ in range(3)]
    if dtype == torch.float32:
        noises = [{k: v.float().double() for k, v in n.items()} for n in noises]
    z_map = H.z_probs(params, data)[ndx].argmax(-1)
    ref = H.theta_probs(params, data, ndx, noises, z_map)
    eng = make_engine(ds, data, params, 4, dtype)
    out = eng.compute_theta_probs(z_map.cuda(), aoi_count=3, particles=3,
                                  local_noise=[L.pack_local_noise(n, dtype, "cuda") for n in noises],
                                  global_noise=[eng.gl.pack_noise(n).cuda() for n in noises])
    assert out.shape == (2, 3, 7, 1)
    assert (out.cpu().double() - ref).abs().max().item() < tol
