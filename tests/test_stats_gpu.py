"""
GPU: post-fit statistics on the device (row N2, csrc/stats.cu) -- the credible intervals of ``compute_params`` against the
reference's own ``cosmos.compute_params`` output (tests/golden/ref_c1_fit.pt / ref_step_hmm.pt: scipy inverse CDFs through
``torch_to_scipy_dist``, cosmos.py:711-784, stats.py:262-293) and against scipy over the ranges the guides visit;
``snr_and_chi2`` on the store's own uint16 pixels against the reference formula (stats.py:29-86).
"""

import numpy as np
import pytest
import torch

from oracle import cosmos_oracle as O
from tapqir_b200.utils.stats import credible_intervals, snr_and_chi2
from tests.step_helpers import golden_c1_fit

pytestmark = pytest.mark.gpu


def test_device_credible_intervals_match_reference_compute_params():
    ds, data, case = golden_c1_fit()
    shapes = {k: v.shape for k, v in O.init_constrained(data).items()}
    cons = O.to_constrained({k: v.reshape(shapes[k]) for k, v in case["final"].items()}, data.P, data.dtype)
    value = lambda name: cons[name].detach().double()
    names = list(case["ci"])
    ours = credible_intervals(names, value, data.P, O.DEFAULT_PRIORS, case["CI"], "cuda")
    for name in names:
        for stat in ("LL", "UL", "Mean"):
            ref = case["ci"][name][stat].double()
            got = ours[name][stat].double().reshape(ref.shape)
            assert (got - ref).abs().max().item() <= 1e-9 * max(1.0, ref.abs().max().item()), (name, stat)


@pytest.mark.parametrize("ci", [0.5, 0.95])
def test_device_inverse_cdfs_match_scipy(ci):
    import ctypes

    import scipy.stats as st

    from tapqir_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(3)
    n = 20000
    conc, rate = np.exp(rng.uniform(np.log(0.05), np.log(3e5), n)), np.exp(rng.uniform(-7, 5, n))
    c1, c0 = np.exp(rng.uniform(np.log(0.3), np.log(2e4), n)), np.exp(rng.uniform(np.log(0.3), np.log(2e4), n))
    dev = lambda a: torch.from_numpy(a).cuda()
    for fn, a, b, dist in ((lib.tq_gamma_interval, conc, rate, st.gamma(conc, scale=1 / rate)), (lib.tq_beta_interval, c1, c0, st.beta(c1, c0))):
        ta, tb = dev(a), dev(b)
        lo, hi = torch.empty_like(ta), torch.empty_like(ta)
        _lib.check(fn(n, _lib.ptr(ta), _lib.ptr(tb), ctypes.c_double(ci), _lib.ptr(lo), _lib.ptr(hi), _lib.stream_ptr()))
        L, U = dist.interval(ci)
        assert np.max(np.abs(lo.cpu().numpy() - L) / L) < 1e-9 and np.max(np.abs(hi.cpu().numpy() - U) / U) < 1e-9


def test_snr_and_chi2_on_the_resident_uint16_pixels():
    from tapqir_b200.utils.simulate import simulate

    ds = simulate(5, 40, C=1, seed=9)
    store = ds.device_store("cuda", torch.float32)
    assert store.pixels.dtype == torch.uint16
    g = torch.Generator().manual_seed(1)
    K, Nt, F, Q, P = 2, 5, 40, 1, 14
    r = lambda *s: torch.rand(*s, generator=g)
    h, w = 1000 + 2000 * r(K, Nt, F, Q), 1.2 + r(K, Nt, F, Q)
    x, y, b = 2 * r(K, Nt, F, Q) - 1, 2 * r(K, Nt, F, Q) - 1, 100 + 80 * r(Nt, F, Q)
    snr, chi2 = snr_and_chi2(store.pixels, h.cuda(), w.cuda(), x.cuda(), y.cuda(), store.xy, b.cuda(), 7.0, 90.0, 0.5, P, None)
    ga = O.gaussian_spots(h.double()[..., None], w.double()[..., None], x.double()[..., None], y.double()[..., None],
                          ds.xy.double()[None, :, :, :, None, :], P)[..., 0, :, :]            # (K, Nt, F, Q, P, P)
    D = ds.images.double()
    sig = ((D - b.double()[..., None, None] - 90.0) * (ga / h.double()[..., None, None])).sum((-1, -2))
    ref_snr = sig / (0.5 + b.double() * 7.0).sqrt()
    ideal = b.double()[..., None, None] + ga.sum(0)
    ref_chi2 = ((D - ideal - 90.0) ** 2 / ideal).mean((-1, -2))
    assert snr.shape == (K, Nt, F, Q) and chi2.shape == (Nt, F, Q)
    assert torch.allclose(snr.double().cpu(), ref_snr, rtol=2e-4, atol=1e-3) and torch.allclose(chi2.double().cpu(), ref_chi2, rtol=2e-4)
