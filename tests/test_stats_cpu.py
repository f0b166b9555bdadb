"""
CPU: the credible intervals of ``compute_params`` (row N2) against the reference's own ``cosmos.compute_params`` /
``stats.torch_to_scipy_dist`` (cosmos.py:711-784, stats.py:262-293), run verbatim by tests/golden/make_golden_step.py on
the parameters at the end of the recorded fits.
"""

import pytest
import torch

from oracle import cosmos_oracle as O
from oracle import hmm_oracle as H
from tapqir_b200.utils.stats import credible_intervals, guide_scipy_dist
from tests.step_helpers import golden_c1_fit


def test_cosmos_credible_intervals_match_reference_compute_params():
    ds, data, case = golden_c1_fit()
    shapes = {k: v.shape for k, v in O.init_constrained(data).items()}
    cons = O.to_constrained({k: v.reshape(shapes[k]) for k, v in case["final"].items()}, data.P, data.dtype)
    value = lambda name: cons[name].detach().double()
    names = ["gain", "pi", "lamda", "proximity", "background", "height", "width", "x", "y"]      # cosmos.py:66-76
    assert list(case["ci"]) == names
    ours = credible_intervals(names, value, data.P, O.DEFAULT_PRIORS, case["CI"])
    for name in names:
        for stat in ("LL", "UL", "Mean"):
            ref = case["ci"][name][stat].double()
            got = ours[name][stat].double().reshape(ref.shape)
            assert (got - ref).abs().max().item() <= 1e-10 * max(1.0, ref.abs().max().item()), (name, stat)
        assert bool((ours[name]["LL"] <= ours[name]["UL"]).all())


@pytest.mark.parametrize("name", ["hmm_c1", "hmm_c2_initial_point"])
def test_hmm_init_and_trans_intervals_match_reference(name):
    """``init`` / ``trans`` (hmm's ci_params, hmm.py:70-81): Dirichlet sites summarised by their Beta marginals."""
    from tests.test_hmm_cpu import hmm_golden_case

    ds, data, case = hmm_golden_case(name)
    shapes = {k: v.shape for k, v in H.init_constrained(data).items()}
    cons = H.to_constrained({k: v.reshape(shapes[k]) for k, v in case["final"].items()}, data.P, data.dtype)
    value = lambda n: cons[n].detach().double()
    ours = credible_intervals(["init", "trans"], value, data.P, O.DEFAULT_PRIORS, case["CI"])
    for site in ("init", "trans"):
        for stat in ("LL", "UL", "Mean"):
            ref = case["ci"][site][stat].double()
            assert (ours[site][stat].double() - ref).abs().max().item() <= 1e-12, (site, stat)
    assert ours["trans"]["Mean"].shape == (data.C, 2, 2) and ours["init"]["Mean"].shape == (data.C, 2)


def test_unknown_latent_is_refused():
    with pytest.raises(NotImplementedError):
        guide_scipy_dist("alpha", lambda n: torch.ones(1), 14, O.DEFAULT_PRIORS)      # crosstalk: out of scope
