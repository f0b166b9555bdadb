"""
CPU checks of the exact arithmetic the kernels execute (csrc/*.cuh compiled for the host by
tests/hostcheck) against (a) the reference's own KSMOGN/gaussian_spots outputs stored in
tests/golden and (b) torch's ATen reparameterisation-gradient functions.
"""

import ctypes

import numpy as np
import pytest
import torch

from tests import hostcheck


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def run_host_ksmogn(case, dtype, fast_oc=None, merge=False):
    hc = hostcheck.load()
    i = case["inputs"]
    K, P = 2, i["P"]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    batch = tuple(i["background"].shape)
    U = int(np.prod(batch))
    spot = lambda t: t.reshape(U, K).t().contiguous().to(tdt)
    h, w, x, y = spot(i["height"]), spot(i["width"]), spot(i["x"]), spot(i["y"])
    b = i["background"].reshape(U).contiguous().to(tdt)
    tgt = i["target_locs"].reshape(U, 2).contiguous().to(tdt)
    val = i["value"].reshape(U, P, P).contiguous().to(tdt)
    off_s = i["offset_samples"].to(tdt).contiguous()
    off_w = torch.distributions.utils.probs_to_logits(i["offset_weights"]).to(tdt).contiguous()
    if merge:  # identical bins merged, as CosmosDataset.device_store does
        from tapqir_b200.utils.dataset import merge_offset_support

        off_s, off_w = (t.contiguous() for t in merge_offset_support(off_s, off_w))
    mcfg = case["m"].to(tdt).contiguous()
    NM = mcfg.shape[0]
    W = case["W"].reshape(NM, U).contiguous().to(tdt)
    logp = torch.empty(NM, U, dtype=tdt)
    g_h, g_w, g_x, g_y = (torch.empty(K, U, dtype=tdt) for _ in range(4))
    g_b, g_rate = torch.empty(U, dtype=tdt), torch.empty(U, dtype=tdt)
    cf = ctypes.c_double if dtype == "f64" else ctypes.c_float
    if fast_oc is not None:
        hc.hc_ksmogn_fast_f32(ctypes.c_int64(U), P, off_s.numel(), fast_oc, _p(h), _p(w), _p(x), _p(y), _p(b),
                              cf(i["gain"].item()), _p(W), _p(val), _p(tgt), _p(off_s), _p(off_w), _p(logp), _p(g_h),
                              _p(g_w), _p(g_x), _p(g_y), _p(g_b), _p(g_rate))
    else:
        fn = getattr(hc, f"hc_ksmogn_{dtype}")
        fn(ctypes.c_int64(U), P, off_s.numel(), NM, _p(h), _p(w), _p(x), _p(y), _p(b), cf(i["gain"].item()), _p(mcfg),
           _p(W), _p(val), _p(tgt), _p(off_s), _p(off_w), _p(logp), _p(g_h), _p(g_w), _p(g_x), _p(g_y), _p(g_b), _p(g_rate))
    back = lambda t: t.t().reshape(batch + (K,)).double()
    gain = i["gain"].item()
    return dict(log_prob=logp.reshape((NM,) + batch).double(), height=back(g_h), width=back(g_w), x=back(g_x), y=back(g_y),
                background=g_b.reshape(batch).double(), gain=-(g_rate.double().sum()) / gain**2)


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("name", ["sim_O3", "hist_O16_C2", "small_P6"])
def test_pixel_math_f64_matches_reference(golden, name):
    case = golden["ksmogn"][name]
    out = run_host_ksmogn(case, "f64")
    assert relerr(out["log_prob"], case["log_prob"]) < 1e-12
    for k in ("height", "width", "x", "y", "background", "gain"):
        assert relerr(out[k], case["grads"][k]) < 1e-10, k


@pytest.mark.parametrize("name", ["sim_O3", "hist_O16_C2", "small_P6"])
def test_pixel_math_f32_within_tolerance(golden, name):
    # fp32 arithmetic against the fp64 reference: north-star tolerance 1e-5 relative (to the
    # largest entry of each tensor)
    case = golden["ksmogn"][name]
    out = run_host_ksmogn(case, "f32")
    assert relerr(out["log_prob"], case["log_prob"]) < 1e-5
    for k in ("height", "width", "x", "y", "background", "gain"):
        assert relerr(out[k], case["grads"][k]) < 1e-5, k


@pytest.mark.parametrize("name,oc", [("sim_O3", 3), ("sim_O3", 0), ("hist_O16_C2", 0), ("small_P6", 0)])
def test_fast_fp32_form_within_tolerance(golden, name, oc):
    """ksmogn_fast.cuh (Stirling lgamma/digamma, base-2 log-sum-exp, cached offsets) vs the reference."""
    case = golden["ksmogn"][name]
    out = run_host_ksmogn(case, "f32", fast_oc=oc)
    assert relerr(out["log_prob"], case["log_prob"]) < 1e-5
    for k in ("height", "width", "x", "y", "background", "gain"):
        assert relerr(out[k], case["grads"][k]) < 1e-5, k


def test_fast_fp32_single_bin_form_within_tolerance(golden):
    """The simulator's three identical offset bins merged to one: the packed single-bin form
    (ksmogn_fast.cuh::pixel_pair_single_bin, closed-form spot-free configuration) vs the reference's 3-bin result."""
    case = golden["ksmogn"]["sim_O3"]
    out = run_host_ksmogn(case, "f32", fast_oc=1, merge=True)
    assert relerr(out["log_prob"], case["log_prob"]) < 1e-5
    for k in ("height", "width", "x", "y", "background", "gain"):
        assert relerr(out[k], case["grads"][k]) < 1e-5, k


def test_digamma_matches_torch():
    hc = hostcheck.load()
    hc.hc_digamma_f64.restype = ctypes.c_double
    hc.hc_digamma_f64.argtypes = [ctypes.c_double]
    xs = np.concatenate([np.geomspace(1e-3, 1e4, 200), [10.0, 1.0, 2.0]])
    ours = np.array([hc.hc_digamma_f64(float(v)) for v in xs])
    ref = torch.digamma(torch.tensor(xs)).numpy()
    np.testing.assert_allclose(ours, ref, rtol=1e-12, atol=1e-13)


def test_std_gamma_grad_matches_aten():
    hc = hostcheck.load()
    hc.hc_std_gamma_grad_f64.restype = ctypes.c_double
    hc.hc_std_gamma_grad_f64.argtypes = [ctypes.c_double, ctypes.c_double]
    rng = np.random.default_rng(0)
    alphas = np.concatenate([rng.uniform(0.05, 8, 300), rng.uniform(8, 500, 300)])
    xs = np.concatenate([rng.gamma(alphas[:300]) + 1e-6, rng.gamma(alphas[300:])])
    # include points inside the Taylor patch around x = alpha and the small-x series
    alphas = np.concatenate([alphas, [50.0, 50.0, 3.0, 0.5]])
    xs = np.concatenate([xs, [50.0, 52.0, 0.3, 0.01]])
    ref = torch._standard_gamma_grad(torch.tensor(alphas), torch.tensor(xs)).numpy()
    ours = np.array([hc.hc_std_gamma_grad_f64(float(a), float(x)) for a, x in zip(alphas, xs)])
    np.testing.assert_allclose(ours, ref, rtol=1e-11, atol=1e-14)


def test_beta_grad_matches_aten():
    hc = hostcheck.load()
    hc.hc_beta_grad_f64.restype = ctypes.c_double
    hc.hc_beta_grad_f64.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(0.2, 6, 300), rng.uniform(6, 800, 300)])
    b = np.concatenate([rng.uniform(0.2, 6, 300), rng.uniform(6, 800, 300)])
    x = np.clip(rng.beta(a, b), 1e-6, 1 - 1e-6)
    a = np.concatenate([a, [50.0, 1.0, 700.0]])
    b = np.concatenate([b, [50.0, 1.0, 700.0]])
    x = np.concatenate([x, [0.5, 0.3, 0.5001]])
    conc = torch.tensor(np.stack([a, b], -1))
    xv = torch.tensor(np.stack([x, 1 - x], -1))
    ref = torch._dirichlet_grad(xv, conc, conc.sum(-1, True).expand_as(conc)).numpy()
    ours0 = np.array([hc.hc_beta_grad_f64(float(xi), float(ai), float(ai + bi)) for xi, ai, bi in zip(x, a, b)])
    ours1 = np.array([hc.hc_beta_grad_f64(float(1 - xi), float(bi), float(ai + bi)) for xi, ai, bi in zip(x, a, b)])
    np.testing.assert_allclose(ours0, ref[:, 0], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(ours1, ref[:, 1], rtol=1e-10, atol=1e-14)


def test_beta_grad_pair_matches_aten():
    """The shared-subexpression pair used by site_eval == two ATen _dirichlet_grad evaluations."""
    hc = hostcheck.load()
    rng = np.random.default_rng(2)
    a = np.concatenate([rng.uniform(6, 1500, 500), rng.uniform(0.3, 6, 100)])
    b = np.concatenate([rng.uniform(6, 1500, 500), rng.uniform(0.3, 900, 100)])
    x = np.clip(rng.beta(a, b), 1e-6, 1 - 1e-6)
    a, b, x = np.concatenate([a, [50.0, 700.0]]), np.concatenate([b, [50.0, 700.0]]), np.concatenate([x, [0.5, 0.5003]])
    conc = torch.tensor(np.stack([a, b], -1))
    xv = torch.tensor(np.stack([x, 1 - x], -1))
    ref = torch._dirichlet_grad(xv, conc, conc.sum(-1, True).expand_as(conc)).numpy()
    g1, g0 = ctypes.c_double(), ctypes.c_double()
    ours = []
    for xi, ai, bi in zip(x, a, b):
        hc.hc_beta_grad_pair_f64(ctypes.c_double(xi), ctypes.c_double(ai), ctypes.c_double(bi), ctypes.byref(g1), ctypes.byref(g0))
        ours.append((g1.value, g0.value))
    # mathematically identical to ATen; the expansion is ill-conditioned near x = mean, so a different
    # evaluation order moves the result at the 1e-9 level
    np.testing.assert_allclose(np.array(ours), ref, rtol=1e-7, atol=1e-14)


def test_lgamma_pos_matches_scipy():
    from scipy.special import gammaln

    hc = hostcheck.load()
    hc.hc_lgamma_pos.restype = ctypes.c_double
    hc.hc_lgamma_pos.argtypes = [ctypes.c_double]
    xs = np.concatenate([np.geomspace(1e-3, 1e5, 400), [1.0, 2.0, 10.0, 0.5]])
    ours = np.array([hc.hc_lgamma_pos(float(v)) for v in xs])
    np.testing.assert_allclose(ours, gammaln(xs), rtol=1e-13, atol=2e-14)


def test_gamma_sampler_moments():
    hc = hostcheck.load()
    n = 20000
    out = np.empty(n)
    for alpha in (0.3, 2.0, 150.0):
        hc.hc_sample_gamma_f64(ctypes.c_uint64(7), ctypes.c_uint64(3), ctypes.c_double(alpha), n,
                               out.ctypes.data_as(ctypes.c_void_p))
        assert abs(out.mean() - alpha) < 5 * np.sqrt(alpha / n)
        assert abs(out.var() - alpha) < 0.1 * alpha


def test_normal_uniform_pairs_of_a_philox_block():
    """Both Box-Muller branches and the two uniforms of a block: standard moments, no correlation between any two of the
    four, Kolmogorov-Smirnov against the normal / uniform laws."""
    from scipy import stats

    hc = hostcheck.load()
    n = 200000
    out = np.empty((n, 4), dtype=np.float32)
    hc.hc_normal_uniform_pairs(ctypes.c_uint64(5), ctypes.c_uint64(9), n, out.ctypes.data_as(ctypes.c_void_p))
    x = out.astype(np.float64)
    for j in (0, 2):
        assert abs(x[:, j].mean()) < 5 / np.sqrt(n) and abs(x[:, j].var() - 1) < 5 * np.sqrt(2 / n)
        assert abs(stats.kurtosis(x[:, j])) < 0.05
        assert stats.kstest(x[:, j], "norm").pvalue > 1e-3
    for j in (1, 3):
        assert x[:, j].min() > 0 and x[:, j].max() <= 1
        assert stats.kstest(x[:, j], "uniform").pvalue > 1e-3
    c = np.corrcoef(x.T)
    assert np.abs(c - np.eye(4)).max() < 5 / np.sqrt(n)
    # ... and of the squares (Box-Muller's two branches share their radius: independent only if the angle is uniform)
    c2 = np.corrcoef((x[:, [0, 2]] ** 2).T)
    assert abs(c2[0, 1]) < 5 / np.sqrt(n)


def test_gamma_draws_that_share_their_trials():
    """Two gamma draws fed from one GammaTrials (second draw starts on the first block's spare pair): right marginals
    (KS against scipy), uncorrelated, and their ratio g1 / (g1 + g2) is Beta(c1, c0)."""
    from scipy import stats

    hc = hostcheck.load()
    n = 100000
    out = np.empty((n, 2))
    for c1, c0 in ((50.0, 50.0), (100.0, 100.0), (2.5, 7.0), (0.6, 1.7), (12.0, 0.8)):
        hc.hc_sample_gamma_pair_f32(ctypes.c_uint64(11), ctypes.c_uint64(2), ctypes.c_float(c1), ctypes.c_float(c0), n,
                                    out.ctypes.data_as(ctypes.c_void_p))
        assert stats.kstest(out[:, 0], "gamma", args=(c1,)).pvalue > 1e-3, (c1, c0)
        assert stats.kstest(out[:, 1], "gamma", args=(c0,)).pvalue > 1e-3, (c1, c0)
        assert abs(np.corrcoef(out.T)[0, 1]) < 5 / np.sqrt(n), (c1, c0)
        assert stats.kstest(out[:, 0] / out.sum(1), "beta", args=(c1, c0)).pvalue > 1e-3, (c1, c0)
