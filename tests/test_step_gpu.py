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


def make_engine(ds, data, params, nb, fb, dtype, merge_offsets=True, **kw):
    from tapqir_b200.models.engine import CosmosEngine

    store = ds.device_store("cuda", dtype, merge_offsets=merge_offsets)
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
    # AOI sizes other than 14 (generic sweep of the likelihood kernel, P-dependent ranges and constraints) with launch sizes
    # that are not a multiple of a warp's four patches, and a single-AOI, single-frame minibatch
    dict(N=3, F=5, C=1, nb=3, fb=5, seed=5, P=10),
    dict(N=2, F=3, C=2, nb=1, fb=3, seed=6, P=17),
    dict(N=3, F=4, C=1, nb=1, fb=1, seed=7, f64_only=True),
]


# fp64 gradients: 1e-8 (the paired Rice expansion of the Beta reparameterisation gradient is
# ill-conditioned near x = mean, evaluation order moves it at the 1e-9 level; see test_hostcheck_math)
PRECISIONS = [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)]
# `f64_only`: a single-unit minibatch leaves each gradient tensor with one unit's entries as its scale: d loss / d m_probs is
# q1 q0 x (differences of the four log-likelihoods, |L| ~ 3e3), and fp32 rounding of L alone is 2e-4 absolute -- 3e-5 to 1e-4
# of a gradient of 20-70 (the host build of the same arithmetic shows the same); in a real minibatch the tensor's scale is
# ~1e3.  The launch-shape path that case is here for does not depend on the dtype.
STEP_CASES = [pytest.param({k: v for k, v in cfg.items() if k != "f64_only"}, dtype, ltol, gtol,
                           id=f"cfg{i}-{'f64' if dtype == torch.float64 else 'f32'}")
              for i, cfg in enumerate(CONFIGS) for dtype, ltol, gtol in PRECISIONS
              if not (cfg.get("f64_only") and dtype == torch.float32)]


@pytest.mark.parametrize("cfg,dtype,ltol,gtol", STEP_CASES)
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
    eng.keep_intermediates = True   # the fused step kernel keeps the samples in shared memory unless asked
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


@pytest.mark.parametrize("cfg", [dict(N=4, F=6, C=1, nb=3, fb=4, seed=0), dict(N=4, F=5, C=2, nb=4, fb=3, seed=1)])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_compute_probs_matches_oracle(cfg, dtype, tol):
    """Row N1: z / theta posteriors for given guide draws (cosmos.py:609-672)."""
    ds, data, params, ndx, fdx, _ = make_problem(**cfg)
    if dtype == torch.float32:
        params = {k: v.float().double() for k, v in params.items()}
    g = torch.Generator().manual_seed(5)
    noises = [O.draw_noise(params, data, ndx, fdx, g) for _ in range(3)]
    if dtype == torch.float32:
        noises = [{k: v.float().double() for k, v in n.items()} for n in noises]
    z_ref, th_ref = O.compute_probs(params, data, ndx, fdx, noises)
    eng = make_engine(ds, data, params, cfg["nb"], cfg["fb"], dtype)
    flat = [flat_inputs(data, params, n, dtype) for n in noises]
    z, th = eng.compute_probs(particles=3, ndx=ndx.to(torch.int32).cuda(), fdx=fdx.to(torch.int32).cuda(),
                              local_noise=[f[4].cuda() for f in flat], global_noise=[f[5].cuda() for f in flat])
    assert (z.double().cpu() - z_ref).abs().max().item() < tol
    assert (th.double().cpu() - th_ref).abs().max().item() < tol
    assert torch.allclose(z.double().cpu().sum(-1), torch.ones_like(z_ref[..., 0]), atol=1e-5)


def test_c1_replayed_fit_tracks_oracle():
    """
    BASELINE config 1 (N=5 AOIs x F=100 frames, full batch, 100 SVI iterations) in replay mode: the
    fp32 engine is fed the oracle's base variates at every iteration.  After the fixed 100 iterations
    the loss agrees to 1e-5, every variational parameter to 1e-4 (of its largest entry) and the fitted
    posteriors z_probs / theta_probs (same particles) to 1e-4 absolute.
    """
    from tapqir_b200.models.cosmos import cosmos
    from tapqir_b200.utils.simulate import simulate

    ds = simulate(5, 100, seed=0)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    svi = O.OracleSVI(data, nbatch_size=5, fbatch_size=100, seed=0)
    model = cosmos(device="cuda", dtype="float")
    model.data = ds
    model.init(lr=0.005, nbatch_size=5, fbatch_size=100, seed=0)
    eng = model.engine
    g = torch.Generator().manual_seed(0)
    ndx, fdx = torch.arange(5), torch.arange(100)
    for it in range(100):
        cur = {k: v.detach().clone() for k, v in svi.params.items()}
        noise = {k: v.float().double() for k, v in O.draw_noise(cur, data, ndx, fdx, g).items()}
        ref_loss = svi.step(ndx, fdx, noise)
        _, _, _, _, lnoise, gnoise = flat_inputs(data, cur, noise, torch.float32)
        loss = eng.step(local_noise=lnoise.cuda(), global_noise=gnoise.cuda()).item()  # identity indices
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    final = {k: v.detach().clone() for k, v in svi.params.items()}
    ours = eng.named_unconstrained()
    for k, v in final.items():
        err = (ours[k].double().cpu().reshape(v.shape) - v).abs().max().item()
        assert err <= 1e-4 * max(1.0, v.abs().max().item()), (k, err)
    # fitted posteriors with identical particles
    on = torch.arange(int(data.is_ontarget.sum()))
    noises = [{k: v.float().double() for k, v in O.draw_noise(final, data, on, fdx, g).items()} for _ in range(5)]
    z_ref, th_ref = O.compute_probs(final, data, on, fdx, noises)
    flat = [flat_inputs(data, final, n, torch.float32) for n in noises]
    z, th = eng.compute_probs(particles=5, ndx=on.to(torch.int32).cuda(), fdx=fdx.to(torch.int32).cuda(),
                              local_noise=[f[4].cuda() for f in flat], global_noise=[f[5].cuda() for f in flat])
    assert (z.double().cpu() - z_ref).abs().max().item() < 1e-4
    assert (th.double().cpu() - th_ref).abs().max().item() < 1e-4


def test_c1_fit_matches_oracle_in_distribution():
    """
    Same configuration with independent random streams (torch CPU generator in the oracle, in-kernel
    Philox here): agreement is statistical.  One global draw per step makes single trajectories noisy
    (a few % after 100 iterations), so the MEAN over 6 engine seeds is compared with the mean over 2
    oracle seeds: |difference| <= max(3 %, 3 sample standard deviations of the engine runs).
    """
    from tapqir_b200.models.cosmos import cosmos
    from tapqir_b200.utils.simulate import simulate

    ds = simulate(5, 100, seed=0)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    names = ("gain_loc", "lamda_loc", "proximity_loc", "pi_mean", "gain_beta", "lamda_beta", "pi_size", "proximity_size")
    local_names = ("b_loc", "b_beta", "h_loc", "w_mean", "w_size", "size", "background_mean_loc", "m_probs")

    def summary(get):
        out = {f"{n}[{i}]": v for n in names for i, v in enumerate(get(n).double().cpu().reshape(-1).tolist())}
        out.update({f"mean({n})": get(n).double().cpu().mean().item() for n in local_names})
        return out

    refs = []
    for seed in (0, 1):
        svi = O.OracleSVI(data, nbatch_size=5, fbatch_size=100, seed=seed)
        losses = [svi.step() for _ in range(100)]
        c = svi.constrained()
        refs.append(dict(summary(lambda n: c[n]), loss=sum(losses[-10:]) / 10))
    runs = []
    for seed in range(6):
        model = cosmos(device="cuda", dtype="float")
        model.data = ds
        model.init(lr=0.005, nbatch_size=5, fbatch_size=100, seed=seed)
        losses = [model.step().item() for _ in range(100)]
        assert model.engine.iteration == 100
        runs.append(dict(summary(model.param), loss=sum(losses[-10:]) / 10))
    bad = {}
    for key in refs[0]:
        ref = sum(r[key] for r in refs) / len(refs)
        vals = torch.tensor([r[key] for r in runs], dtype=torch.float64)
        tol = max(0.03 * abs(ref), 3 * vals.std().item(), 1e-3)
        if abs(vals.mean().item() - ref) > tol:
            bad[key] = (vals.mean().item(), ref, tol)
    assert not bad, bad
    # posterior summaries exist and are consistent
    z = model.z_probs
    assert z.shape == (5, 100, 1, 2) and float(z[2:].abs().max()) == 0  # off-target AOIs stay 0
    assert torch.allclose(z[:2].sum(-1), torch.ones(2, 100, 1), atol=1e-4)


@pytest.mark.parametrize("C", [1, 2])
def test_split_global_reverse_mode_equals_one_shot(C):
    """tq_cosmos_globals_prepare + tq_cosmos_globals_finish (what the engine runs) against the one-shot
    tq_cosmos_globals_grad on the accumulators of the same step."""
    import ctypes

    from tapqir_b200 import _lib

    cfg = dict(N=4, F=6, C=C, nb=4, fb=6, seed=11)
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    eng = make_engine(ds, data, params, 4, 6, torch.float64)
    loss = eng.step(update=False, **replay_args(eng, data, params, ndx, fdx, noise, torch.float64)).item()
    split = eng.ggrads.clone()
    one = torch.zeros_like(eng.ggrads)
    parts = torch.zeros(2 + 2 * 4, dtype=torch.float64, device="cuda")
    loss1 = torch.zeros(1, dtype=torch.float64, device="cuda")
    p = _lib.ptr
    _lib.check(eng.lib.tq_cosmos_globals_grad(eng.code, eng.C, p(eng.gparams), ctypes.byref(eng.mc), p(eng.gstate), p(eng.acc),
                                              eng.sN, eng.sF, p(one), p(parts), p(loss1), _lib.stream_ptr(eng.device)),
               "tq_cosmos_globals_grad")
    torch.cuda.synchronize()
    assert abs(loss - loss1.item()) <= 1e-13 * abs(loss)
    assert (split - one).abs().max().item() <= 1e-11 * one.abs().max().item()


@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)])
def test_step_with_unmerged_offset_bins(dtype, ltol, gtol):
    """The simulator's three identical offset bins kept as three (merge_offsets=False): the O = 3 kernels."""
    cfg = dict(N=5, F=40, C=1, nb=5, fb=40, seed=13)
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    if dtype == torch.float32:
        params = {k: v.float().double() for k, v in params.items()}
        noise = {k: v.float().double() for k, v in noise.items()}
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    eng = make_engine(ds, data, params, cfg["nb"], cfg["fb"], dtype, merge_offsets=False)
    assert eng.store.offset_samples.numel() == 3
    loss = eng.step(update=False, **replay_args(eng, data, params, ndx, fdx, noise, dtype)).item()
    assert abs(loss - ref_loss) <= ltol * abs(ref_loss)
    assert not compare_grads(eng.named_grads(), ref_grads, gtol)
    merged = make_engine(ds, data, params, cfg["nb"], cfg["fb"], dtype)
    assert merged.store.offset_samples.numel() == 1

