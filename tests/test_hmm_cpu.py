"""
CPU: the hmm oracle's chain algebra against brute-force enumeration of every z path (tiny F), the host-side
layouts of the hmm variant, and -- at the end -- oracle and host build of the hmm kernels' arithmetic against golden SVI
iterations of the reference's own models/hmm.py (tests/golden/ref_step_hmm.pt).
"""

import itertools

import torch

from oracle import cosmos_oracle as O
from oracle import hmm_oracle as H
from tapqir_b200.models import layout as L
from tapqir_b200.utils.simulate import simulate


def small_problem(N=3, F=3, C=1, seed=0):
    ds = simulate(N, F, C=C, seed=seed)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    g = torch.Generator().manual_seed(seed + 5)
    params = H.to_unconstrained(H.init_constrained(data), data.P, data.dtype)
    for v in params.values():
        v.add_(0.4 * torch.randn(v.shape, generator=g, dtype=v.dtype))
    ndx = torch.arange(N)
    noise = H.draw_noise(params, data, ndx, g)
    return ds, data, params, ndx, noise


def test_forward_recursion_equals_path_enumeration():
    """sum over all 2^F paths of q(path) [log p(path, emissions) - log q(path)] == the forward-recursion ELBO."""
    ds, data, params, ndx, noise = small_problem(N=2, F=3)
    total, parts = H.elbo(params, data, ndx, noise, return_parts=True)
    p = H.to_constrained(params, data.P, data.dtype)
    zt = p["z_trans"][ndx]                                                    # (nb,F,C,z',z)
    logq = torch.distributions.Categorical(probs=zt, validate_args=False).logits
    ont = data.is_ontarget[ndx].long()
    logp_init = torch.distributions.Categorical(probs=O.expand_offtarget(parts["init"])[:, :, ont].permute(2, 0, 1),
                                                validate_args=False).logits
    logp_trans = torch.distributions.Categorical(probs=O.expand_offtarget(parts["trans"])[:, :, :, ont].permute(3, 0, 1, 2),
                                                 validate_args=False).logits
    em = parts["emission"]                                                    # (z,nb,F,C)
    nb, F, C = len(ndx), data.F, data.C
    brute = torch.zeros(nb, C, dtype=data.dtype)
    for path in itertools.product(range(2), repeat=F):
        lq = torch.zeros(nb, C, dtype=data.dtype)
        lp = torch.zeros(nb, C, dtype=data.dtype)
        prev = 0
        for f, z in enumerate(path):
            lq = lq + logq[:, f, :, prev, z]
            lp = lp + (logp_init[:, :, z] if f == 0 else logp_trans[:, :, prev, z]) + em[z, :, f, :]
            prev = z
        brute = brute + lq.exp() * (lp - lq)
    assert torch.allclose(brute, parts["e_chain"] + parts["e_emit"], rtol=1e-12, atol=1e-9)


def test_z_probs_are_the_forward_marginals():
    ds, data, params, ndx, noise = small_problem(N=2, F=4)
    zp = H.z_probs(params, data)
    assert zp.shape == (2, 4, 1, 2)
    assert torch.allclose(zp.sum(-1), torch.ones(2, 4, 1, dtype=zp.dtype))


def test_hmm_layouts_round_trip():
    ll, gl = L.HmmLocalLayout(3, 4, 2), L.HmmGlobalLayout(2)
    assert gl.numel == 4 + 11 * 2 and gl.noise_numel == 2 + 7 * 2
    g = torch.Generator().manual_seed(0)
    named = {"m_probs": torch.randn(2, 2, 3, 4, 2, generator=g), "z_trans": torch.randn(3, 4, 2, 2, 2, generator=g)}
    flat = torch.zeros(ll.numel)
    for n, s in ll.shapes.items():
        if n not in ("m_probs_z0", "m_probs_z1", "z_trans"):
            named[n] = torch.randn(s, generator=g)
    ll.load_named(flat, named)
    back = ll.named(flat)
    for n, t in named.items():
        assert torch.equal(back[n], t), n
    # z_trans sits behind the cosmos layout and m_probs[z = 1]
    assert ll.offsets["z_trans"] == ll.std_numel + 2 * 3 * 4 * 2


# ---- the kernels' arithmetic (csrc/cosmos_hmm.cuh + the hmm global sites) compiled for the host, against the oracle ----
import ctypes

import pytest

from tests import hostcheck
from tests.step_helpers import compare_grads


def host_hmm_step(hc, data, params, ndx, noise, dtype):
    ll, gl = L.HmmLocalLayout(data.Nt, data.F, data.C), L.HmmGlobalLayout(data.C)
    lparams = torch.zeros(ll.numel, dtype=dtype)
    ll.load_named(lparams, params)
    gparams = gl.pack({k: params[k] for k in gl.shapes}, dtype=torch.float64)
    lnoise = L.pack_local_noise(noise, dtype, "cpu")
    gnoise = gl.pack_noise(noise)
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, data.P, torch.float64)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    ndx32 = ndx.to(torch.int32).contiguous()
    pixels, xy = data.images.to(dtype).contiguous(), data.xy.to(dtype).contiguous()
    ont, mask = data.is_ontarget.to(torch.uint8).contiguous(), data.mask.to(torch.uint8).contiguous()
    off_s, off_w = data.offset_samples.to(dtype).contiguous(), data.offset_logits.to(dtype).contiguous()
    lgrads = torch.empty_like(lparams)
    ggrads = torch.zeros(gl.numel, dtype=torch.float64)
    fn = hc.hc_hmm_step_f64 if dtype == torch.float64 else hc.hc_hmm_step_f32
    fn.restype = ctypes.c_double
    loss = fn(len(ndx), data.Nt, data.F, data.C, data.P, off_s.numel(), p(ndx32), p(pixels), p(xy), p(ont), p(mask), p(off_s),
              p(off_w), ctypes.byref(mc), ctypes.c_double(data.Nt / len(ndx)), p(lparams), p(gparams), p(lnoise), p(gnoise),
              p(lgrads), p(ggrads))
    grads = dict(ll.named(lgrads))
    grads.update(gl.views(ggrads))
    return loss, grads


@pytest.mark.parametrize("cfg", [dict(N=3, F=5, C=1, seed=0), dict(N=3, F=4, C=2, seed=1)])
@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)])
def test_hmm_step_arithmetic_matches_oracle_on_the_host(cfg, dtype, ltol, gtol):
    hc = hostcheck.load()
    ds, data, params, ndx, noise = small_problem(**cfg)
    ndx = torch.tensor([2, 0])
    if dtype == torch.float32:
        params = {k: v.float().double() for k, v in params.items()}
    g = torch.Generator().manual_seed(3)
    noise = H.draw_noise(params, data, ndx, g)
    if dtype == torch.float32:
        noise = {k: v.float().double() for k, v in noise.items()}
    ref_loss, ref_grads = H.loss_and_grads(params, data, ndx, noise)
    loss, grads = host_hmm_step(hc, data, params, ndx, noise, dtype)
    assert abs(loss - ref_loss) <= ltol * abs(ref_loss)
    bad = compare_grads(grads, ref_grads, gtol)
    assert not bad, bad


# ---- golden SVI iterations of the reference's own hmm.py (sequential form) ---------------------------------------------
HMM_GOLDEN = ["hmm_c1", "hmm_c2_initial_point"]


def hmm_golden_case(name):
    """tests/golden/ref_step_hmm.pt: ``models/hmm.py`` (init_parameters, guide, model) run verbatim with ``vectorized=False``
    -- the reference's own pyro.markov + TraceEnum_ELBO form of the model -- by tests/golden/make_golden_step.py under
    tests/golden/minipyro.py, which evaluates the expectation over the guide's enumerated chain and spot presences by
    brute force over every enumerated dimension (no forward recursion, no closed form)."""
    from pathlib import Path

    from tapqir_b200.utils.dataset import CosmosDataset

    case = torch.load(Path(__file__).resolve().parent / "golden" / "ref_step_hmm.pt", weights_only=False)[name]
    ds = CosmosDataset(case["images"].to(torch.float32), case["xy"], case["is_ontarget"], case["mask"].clone(), None,
                       case["offset_samples"], case["offset_weights"])
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    return ds, data, case


def adam_update(p, grads, m, v, t, lr=0.005):
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) on every tensor (model.py:168-171)."""
    for k in p:
        m[k] = 0.9 * m[k] + 0.1 * grads[k]
        v[k] = 0.999 * v[k] + 0.001 * grads[k] ** 2
        p[k] = p[k] - lr * (m[k] / (1 - 0.9 ** t)) / ((v[k] / (1 - 0.999 ** t)).sqrt() + 1e-8)


@pytest.mark.parametrize("name", HMM_GOLDEN)
def test_hmm_oracle_matches_reference_model_code(name):
    """Initial parameters (hmm.py:419-467), then every recorded iteration: the closed form of SURVEY App. B.2 (forward
    recursion) against the reference code's brute-force value -- loss 1e-12, gradients 1e-8 of each tensor's largest entry
    (measured <= 7e-10), parameters after the last Adam update 1e-10."""
    ds, data, case = hmm_golden_case(name)
    init = H.to_unconstrained(H.init_constrained(data), data.P, data.dtype)
    assert set(init) == set(case["init_unconstrained"])
    for k, v in case["init_unconstrained"].items():
        assert (init[k] - v.reshape(init[k].shape)).abs().max().item() <= 1e-13 * max(1.0, v.abs().max().item()), k
    p = {k: v.reshape(init[k].shape).clone() for k, v in case["start"].items()}
    m, v2 = {k: torch.zeros_like(x) for k, x in p.items()}, {k: torch.zeros_like(x) for k, x in p.items()}
    for t, step in enumerate(case["steps"], 1):
        loss, grads = H.loss_and_grads(p, data, step["ndx"], step["noise"])
        assert abs(loss - step["loss"]) <= 1e-12 * abs(step["loss"]), (t, loss, step["loss"])
        ref_grads = {k: g.reshape(p[k].shape) for k, g in step["grads"].items()}
        bad = compare_grads(grads, ref_grads, 1e-8)
        assert not bad, (t, bad)
        adam_update(p, grads, m, v2, t)
    for k, v in case["final"].items():
        assert (p[k] - v.reshape(p[k].shape)).abs().max().item() <= 1e-10 * max(1.0, v.abs().max().item()), k


@pytest.mark.parametrize("name", HMM_GOLDEN)
@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)])
def test_hmm_host_arithmetic_matches_reference_model_code(name, dtype, ltol, gtol):
    """The hmm kernels' arithmetic (host build) against the same file, at the reference's parameters of every iteration;
    fp32: nothing rounded on the reference side (global gradients on the scale of their parameter pair)."""
    hc = hostcheck.load()
    ds, data, case = hmm_golden_case(name)
    shapes = {k: v.shape for k, v in H.init_constrained(data).items()}
    p = {k: v.reshape(shapes[k]).clone() for k, v in case["start"].items()}
    m, v2 = {k: torch.zeros_like(x) for k, x in p.items()}, {k: torch.zeros_like(x) for k, x in p.items()}
    for t, step in enumerate(case["steps"], 1):
        loss, grads = host_hmm_step(hc, data, p, step["ndx"], step["noise"], dtype)
        assert abs(loss - step["loss"]) <= ltol * abs(step["loss"]), (t, loss, step["loss"])
        ref_grads = {k: g.reshape(shapes[k]) for k, g in step["grads"].items()}
        if dtype == torch.float64:
            bad = compare_grads(grads, ref_grads, gtol)
        else:
            bad = compare_grads(grads, ref_grads, gtol, names=[k for k in ref_grads if k not in H.GLOBAL_PARAMS])
            bad.update(hmm_global_grads(grads, ref_grads, gtol))
        assert not bad, (t, bad)
        _, og = H.loss_and_grads(p, data, step["ndx"], step["noise"])
        adam_update(p, og, m, v2, t)


HMM_GLOBAL_PAIRS = [("gain_loc", "gain_beta"), ("lamda_loc", "lamda_beta"), ("proximity_loc", "proximity_size"),
                    ("init_mean", "init_size"), ("trans_mean", "trans_size")]


def hmm_global_grads(ours, ref, tol):
    """Global gradients against the largest entry of their distribution's parameter pair (tests/step_helpers.py)."""
    bad = {}
    for pair in HMM_GLOBAL_PAIRS:
        scale = max(ref[k].double().abs().max().item() for k in pair)
        for k in pair:
            err = (ours[k].double().cpu().reshape(ref[k].shape) - ref[k].double()).abs().max().item()
            if not err <= tol * scale:
                bad[k] = err / scale
    return bad


@pytest.mark.parametrize("name", HMM_GOLDEN + ["hmm_zprobs_only"])
def test_hmm_z_probs_match_reference_model_code(name):
    """hmm.z_probs (hmm.py:627-633, the reference's own parallel scan ``_sequential_logmatmulexp`` :480-533) at the
    recorded final parameters, chains of 3, 4 and 23 frames: the oracle's plain forward recursion, 1e-13."""
    ds, data, case = hmm_golden_case(name)
    shapes = {k: v.shape for k, v in H.init_constrained(data).items()}
    final = {k: v.reshape(shapes[k]) for k, v in case["final"].items()}
    zp = H.z_probs(final, data)
    assert zp.shape == case["z_probs"].shape == (data.Nt, data.F, data.C, 2)
    assert (zp - case["z_probs"]).abs().max().item() <= 1e-13
