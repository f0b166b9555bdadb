"""
CPU: the credible intervals of ``compute_params`` (row N2) against the reference's own ``cosmos.compute_params`` /
``stats.torch_to_scipy_dist`` (cosmos.py:711-784, stats.py:262-293), run verbatim by tests/golden/make_golden_step.py on
the parameters at the end of the recorded fits.  The product evaluates the inverse CDFs on the device
(csrc/stats.cu); here the HOST build of the same arithmetic (csrc/stats_math.cuh through tests/hostcheck) is pinned to
those goldens and to scipy over the parameter ranges the guides visit.  The CUDA kernels: tests/test_stats_gpu.py.
"""

import ctypes

import numpy as np
import pytest
import torch

from oracle import cosmos_oracle as O
from oracle import hmm_oracle as H
from tapqir_b200.utils.stats import guide_family
from tests import hostcheck
from tests.step_helpers import golden_c1_fit


def host_credible_intervals(ci_params, value, P, priors, CI):
    """utils/stats.py::credible_intervals with the host build of the interval kernels in place of the CUDA ones."""
    hc = hostcheck.load()
    out = {}
    for name in ci_params:
        family, p1, p2, low, scale, mean = guide_family(name, value, P, priors)
        shape = p1.shape
        a = p1.double().reshape(-1).contiguous()
        b = p2.double().expand(shape).reshape(-1).contiguous()
        lo, hi = torch.empty_like(a), torch.empty_like(a)
        fn = hc.hc_gamma_interval if family == "gamma" else hc.hc_beta_interval
        fn(ctypes.c_int64(a.numel()), ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_double(CI),
           ctypes.c_void_p(lo.data_ptr()), ctypes.c_void_p(hi.data_ptr()))
        out[name] = {"LL": (low + scale * lo).reshape(shape), "UL": (low + scale * hi).reshape(shape), "Mean": mean.double()}
    return out


def test_cosmos_credible_intervals_match_reference_compute_params():
    ds, data, case = golden_c1_fit()
    shapes = {k: v.shape for k, v in O.init_constrained(data).items()}
    cons = O.to_constrained({k: v.reshape(shapes[k]) for k, v in case["final"].items()}, data.P, data.dtype)
    value = lambda name: cons[name].detach().double()
    names = ["gain", "pi", "lamda", "proximity", "background", "height", "width", "x", "y"]      # cosmos.py:66-76
    assert list(case["ci"]) == names
    ours = host_credible_intervals(names, value, data.P, O.DEFAULT_PRIORS, case["CI"])
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
    ours = host_credible_intervals(["init", "trans"], value, data.P, O.DEFAULT_PRIORS, case["CI"])
    for site in ("init", "trans"):
        for stat in ("LL", "UL", "Mean"):
            ref = case["ci"][site][stat].double()
            assert (ours[site][stat].double() - ref).abs().max().item() <= 1e-11, (site, stat)
    assert ours["trans"]["Mean"].shape == (data.C, 2, 2) and ours["init"]["Mean"].shape == (data.C, 2)


@pytest.mark.parametrize("ci", [0.5, 0.95, 0.999])
def test_inverse_cdfs_match_scipy_over_the_guides_ranges(ci):
    """Gamma concentrations 0.05 .. 3e5 (height guides of absent spots relax to ~1; backgrounds reach 1e4), Beta
    concentrations 0.3 .. 2e4 each (sizes 2 .. 4e4): quantiles within 1e-9 of scipy's."""
    import scipy.stats as st

    hc = hostcheck.load()
    rng = np.random.default_rng(int(ci * 1000))
    n = 3000
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    conc, rate = np.exp(rng.uniform(np.log(0.05), np.log(3e5), n)), np.exp(rng.uniform(-7, 5, n))
    lo, hi = np.empty(n), np.empty(n)
    hc.hc_gamma_interval(ctypes.c_int64(n), p(conc), p(rate), ctypes.c_double(ci), p(lo), p(hi))
    L, U = st.gamma(conc, scale=1 / rate).interval(ci)
    assert np.max(np.abs(lo - L) / L) < 1e-9 and np.max(np.abs(hi - U) / U) < 1e-9
    c1, c0 = np.exp(rng.uniform(np.log(0.3), np.log(2e4), n)), np.exp(rng.uniform(np.log(0.3), np.log(2e4), n))
    hc.hc_beta_interval(ctypes.c_int64(n), p(c1), p(c0), ctypes.c_double(ci), p(lo), p(hi))
    L, U = st.beta(c1, c0).interval(ci)
    assert np.max(np.abs(lo - L) / L) < 1e-9 and np.max(np.abs(hi - U) / U) < 1e-9
    assert np.max(np.abs((1 - hi) - (1 - U)) / (1 - U)) < 1e-7   # the upper end measured from 1


def test_unknown_latent_is_refused():
    with pytest.raises(NotImplementedError):
        guide_family("alpha", lambda n: torch.ones(1), 14, O.DEFAULT_PRIORS)      # crosstalk: out of scope
