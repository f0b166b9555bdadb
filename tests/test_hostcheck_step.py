"""
CPU: the kernel pipeline's arithmetic (csrc/cosmos_local.cuh, cosmos_globals.cuh, ksmogn_core.cuh
compiled for the host) against the fp64 oracle: loss and all 20 gradients of one SVI step in
replay mode (explicit minibatch indices + injected base variates).
"""

import pytest
import torch

from oracle import cosmos_oracle as O
from tests import hostcheck
from tapqir_b200.models import layout as L
from tests.step_helpers import check_global_grads, compare_grads, host_step, make_problem


@pytest.mark.parametrize("cfg", [
    dict(N=4, F=6, C=1, nb=3, fb=4, seed=0),
    dict(N=4, F=5, C=2, nb=4, fb=3, seed=1),
    dict(N=3, F=4, C=1, nb=3, fb=4, seed=2, offsets="hist"),
    dict(N=4, F=6, C=1, nb=2, fb=6, seed=3, perturb=False),
    dict(N=3, F=5, C=1, nb=3, fb=5, seed=5, P=10),   # another AOI size: the +-(P+1)/2 position range, the proximity constraint
    dict(N=2, F=3, C=1, nb=1, fb=3, seed=6, P=17),
])
def test_step_f64_matches_oracle(cfg):
    hc = hostcheck.load()
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    loss, grads, _ = host_step(hc, data, params, ndx, fdx, noise, torch.float64)
    assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
    bad = compare_grads(grads, ref_grads, 1e-9)
    assert not bad, bad


def test_step_respects_aoi_mask():
    hc = hostcheck.load()
    ds, data, params, ndx, fdx, noise = make_problem(N=4, F=5, nb=4, fb=5, seed=4)
    data.mask[1] = False
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    loss, grads, _ = host_step(hc, data, params, ndx, fdx, noise, torch.float64)
    assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
    assert not compare_grads(grads, ref_grads, 1e-9)
    assert grads["b_loc"][1].abs().max() == 0


@pytest.mark.parametrize("cfg", [dict(N=4, F=6, C=1, nb=3, fb=4, seed=0), dict(N=4, F=5, C=2, nb=4, fb=3, seed=1)])
def test_step_f32_within_tolerance(cfg):
    """fp32 arithmetic (fp32 parameters, variates and pixels) vs the fp64 oracle fed the same
    fp32-rounded inputs: loss 1e-6, gradients 1e-5 of each tensor's largest entry (north-star)."""
    hc = hostcheck.load()
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    params = {k: v.float().double() for k, v in params.items()}
    noise = {k: v.float().double() for k, v in noise.items()}
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    loss, grads, _ = host_step(hc, data, params, ndx, fdx, noise, torch.float32)
    assert abs(loss - ref_loss) <= 1e-6 * abs(ref_loss)
    bad = compare_grads(grads, ref_grads, 1e-5)
    assert not bad, bad


@pytest.mark.parametrize("name", ["c1_initial_point", "c1_perturbed_masked", "c2_hist_offsets", "c1_full_batch"])
def test_step_f64_matches_reference_model_code(name):
    """The kernels' arithmetic (host build) against tests/golden/ref_step.pt directly -- losses and gradients the
    reference's own cosmos.py guide()/model() produced (tests/golden/make_golden_step.py) -- for every recorded
    iteration, at the parameters the reference had at that iteration (replayed here with the oracle's Adam, which
    tests/test_oracle.py pins to the same file)."""
    from tests.step_helpers import golden_step_case, masked_loss_constant

    hc = hostcheck.load()
    ds, data, case = golden_step_case(name)
    cfg = case["config"]
    svi = O.OracleSVI(data, lr=cfg["lr"], nbatch_size=cfg["nb"], fbatch_size=cfg["fb"])
    with torch.no_grad():
        for k, v in svi.params.items():
            v.copy_(case["start"][k].reshape(v.shape))
    for step in case["steps"]:
        params = {k: v.detach().clone() for k, v in svi.params.items()}
        loss, grads, _ = host_step(hc, data, params, step["ndx"], step["fdx"], step["noise"], torch.float64)
        ref_loss = step["loss"] + masked_loss_constant(case, step)
        assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
        ref_grads = {k: g.reshape(params[k].shape) for k, g in step["grads"].items()}
        bad = compare_grads(grads, ref_grads, 1e-8)
        assert not bad, bad
        svi.step(step["ndx"], step["fdx"], step["noise"])


@pytest.mark.parametrize("name", ["c1_initial_point", "c1_perturbed_masked", "c2_hist_offsets", "c1_full_batch"])
def test_step_f32_within_north_star_of_reference_model_code(name):
    """The fp32 production arithmetic against the reference's own fp64 numbers (tests/golden/ref_step.pt), nothing
    rounded on the reference side: loss 1e-6, gradients of the AOI-local tensors 1e-5 of each tensor's largest entry at
    every iteration; global gradients 1e-5 of their own magnitude, or of their forward-error scale where the entry is
    ill-conditioned (step_helpers.check_global_grads; the global parameters are float64 on the device as here)."""
    from tests.step_helpers import golden_step_case, masked_loss_constant

    hc = hostcheck.load()
    ds, data, case = golden_step_case(name)
    cfg = case["config"]
    svi = O.OracleSVI(data, lr=cfg["lr"], nbatch_size=cfg["nb"], fbatch_size=cfg["fb"])
    with torch.no_grad():
        for k, v in svi.params.items():
            v.copy_(case["start"][k].reshape(v.shape))
    for step in case["steps"]:
        params = {k: v.detach().clone() for k, v in svi.params.items()}
        loss, grads, _ = host_step(hc, data, params, step["ndx"], step["fdx"], step["noise"], torch.float32)
        ref_loss = step["loss"] + masked_loss_constant(case, step)
        assert abs(loss - ref_loss) <= 1e-6 * abs(ref_loss)
        ref_grads = {k: g.reshape(params[k].shape) for k, g in step["grads"].items()}
        bad = compare_grads(grads, ref_grads, 1e-5, names=L.LOCAL_NAMES)
        # global gradients: 1e-5 of their OWN magnitude where nothing cancels, 1e-5 of their forward-error scale elsewhere
        bad.update(check_global_grads(grads, ref_grads, params, data, step["ndx"], step["fdx"], step["noise"]))
        assert not bad, bad
        svi.step(step["ndx"], step["fdx"], step["noise"])


def test_c1_hundred_iterations_of_the_kernel_arithmetic_match_the_reference_run():
    """BASELINE configs[0] (tests/golden/ref_c1_fit.pt: the reference's own 100-iteration fit): the fp64 host build of the
    kernels with a dense Adam follows it from the same seed -- every loss 1e-11, final parameters 1e-8; the fp32
    production arithmetic, evaluated every tenth iteration at that trajectory's parameters on identical fp32-rounded
    inputs, stays within loss 1e-6 / gradients 1e-5 of the fp64 one (global gradients: step_helpers.check_global_grads)."""
    from tests.step_helpers import golden_c1_fit

    hc = hostcheck.load()
    ds, data, case = golden_c1_fit()
    cfg = case["config"]
    p = O.to_unconstrained(O.init_constrained(data), data.P, data.dtype)
    m, v2 = {k: torch.zeros_like(x) for k, x in p.items()}, {k: torch.zeros_like(x) for k, x in p.items()}
    ndx, fdx = torch.arange(cfg["N"]), torch.arange(cfg["F"])
    state = torch.get_rng_state()
    try:
        torch.manual_seed(cfg["rng_seed"])
        for t in range(1, cfg["iters"] + 1):
            noise = O.draw_noise(p, data, ndx, fdx)
            loss, grads, _ = host_step(hc, data, p, ndx, fdx, noise, torch.float64)
            assert abs(loss - case["losses"][t - 1].item()) <= 1e-11 * abs(loss), (t, loss)
            if t % 10 == 1:
                # identical (fp32-representable) inputs on both sides, as the north star words it
                pr, nr = {k: v.float().double() for k, v in p.items()}, {k: v.float().double() for k, v in noise.items()}
                loss64, grads64, _ = host_step(hc, data, pr, ndx, fdx, nr, torch.float64)
                loss32, grads32, _ = host_step(hc, data, pr, ndx, fdx, nr, torch.float32)
                assert abs(loss32 - loss64) <= 1e-6 * abs(loss64), t
                ref = {k: g.reshape(p[k].shape) for k, g in grads64.items()}
                bad = compare_grads(grads32, ref, 1e-5, names=L.LOCAL_NAMES)
                bad.update(check_global_grads(grads32, ref, pr, data, ndx, fdx, nr))
                assert not bad, (t, bad)
            for k in p:
                g = grads[k].reshape(p[k].shape)
                m[k] = 0.9 * m[k] + 0.1 * g
                v2[k] = 0.999 * v2[k] + 0.001 * g * g
                p[k] = p[k] - cfg["lr"] * (m[k] / (1 - 0.9 ** t)) / ((v2[k] / (1 - 0.999 ** t)).sqrt() + 1e-8)
    finally:
        torch.set_rng_state(state)
    for k, v in case["final"].items():
        assert (p[k] - v.reshape(p[k].shape)).abs().max().item() <= 1e-8, k
