"""
CPU: the fp32 production form of the guide sites (csrc/cosmos_sites_fast.cuh, compiled for the host)
against the double-precision form (csrc/cosmos_local.cuh::site_eval, itself pinned to the oracle by
test_hostcheck_step.py).  Every output -- sample, log q, d log q / d sample and the four
reparameterisation-map entries, plus the background prior record -- over the regimes the cosmos guide
visits: small / medium / large Gamma concentrations (ATen's Taylor, rational and Rice branches) and
Beta sample sizes from 2.1 (absent spots: the guide relaxes to the flat prior; ATen's small-x series and rational
branches, plain fp32) to 6e4 (Rice expansion and its Taylor patch, reformulated).

Tolerance: 5e-6 of the largest entry of each output over the batch (the north-star asks 1e-5 for the
gradients these feed); measured ~1e-6 or better.
"""

import ctypes

import numpy as np
import pytest
import torch

from oracle import cosmos_oracle as O
from tapqir_b200.models import layout as L
from tests import hostcheck

NOUT = 1 + 6 + 4
NAMES = ["sample", "LQ", "DQ", "A0", "B0", "A1", "B1", "EX_LP", "EX_DP", "EX_GBM", "EX_GBS"]


def _run(hc, mc, s, u0, u1, ubm, ubs, var):
    n = len(u0)
    ref, fast, status = np.zeros((n, NOUT)), np.zeros((n, NOUT)), np.zeros(n, dtype=int)
    buf = (ctypes.c_double * NOUT)()
    f32 = lambda v: float(np.float32(v))
    for i in range(n):
        a, b, c, d = f32(u0[i]), f32(u1[i]), f32(ubm[i]), f32(ubs[i])
        hc.hc_site_eval_f64(s, ctypes.c_double(a), ctypes.c_double(b), ctypes.c_double(c), ctypes.c_double(d),
                            ctypes.byref(mc), ctypes.c_double(var[i]), buf)
        ref[i] = buf[:]
        status[i] = hc.hc_site_eval_fast(s, ctypes.c_float(a), ctypes.c_float(b), ctypes.c_float(c), ctypes.c_float(d),
                                         ctypes.byref(mc), ctypes.c_double(var[i]), buf)
        fast[i] = buf[:]
    return ref, fast, status


def _check(ref, fast, status, nout, min_fast, tol=5e-6):
    ok = status == 0
    assert ok.mean() >= min_fast, f"only {ok.mean():.2f} of the sites took the fp32 path"
    bad = {}
    for j in range(nout):
        r, f = ref[ok, j], fast[ok, j]
        rel = np.abs(f - r).max() / np.abs(r).max()
        if not rel < tol:
            bad[NAMES[j]] = rel
    assert not bad, bad


@pytest.mark.parametrize("site", [0, 1])   # background (with its prior record), height
@pytest.mark.parametrize("lconc", [(-1.0, 2.0), (2.0, 3.0), (3.0, 8.0), (8.0, 12.0)])
def test_gamma_sites(site, lconc):
    hc = hostcheck.load()
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
    g = torch.Generator().manual_seed(int(10 * lconc[0]) + site + 50)
    torch.manual_seed(int(10 * lconc[0]) + site + 51)   # the torch.distributions draws below use the global generator
    n = 1500
    lc = torch.empty(n).uniform_(*lconc, generator=g).double()
    u1 = torch.empty(n).uniform_(-7, 1, generator=g).double()
    u0 = lc - u1
    conc = torch.exp(u0.float().double() + u1.float().double())
    var = torch.distributions.Gamma(conc, torch.ones_like(conc)).sample().float().double().clamp_min(1e-30)
    ubm = u0 + 0.1 * torch.randn(n, generator=g).double()
    ubs = ubm - torch.empty(n).uniform_(0.5, 3.0, generator=g).double()
    ref, fast, status = _run(hc, mc, site, u0.numpy(), u1.numpy(), ubm.numpy(), ubs.numpy(), var.numpy())
    _check(ref, fast, status, NOUT if site == 0 else 7, 0.99)


@pytest.mark.parametrize("site", [3, 5, 8])   # width, x, y
@pytest.mark.parametrize("lsize,min_fast", [((-2.0, 1.0), 0.95), ((1.0, 3.0), 0.85), ((3.0, 5.0), 0.95), ((5.0, 8.0), 0.99),
                                            ((8.0, 11.0), 0.99)])
def test_beta_sites(site, lsize, min_fast):
    hc = hostcheck.load()
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
    g = torch.Generator().manual_seed(int(10 * lsize[0]) + site)
    torch.manual_seed(int(10 * lsize[0]) + site + 1)     # the torch.distributions draws below use the global generator
    n = 1500
    u1 = torch.empty(n).uniform_(*lsize, generator=g).double()
    u0 = torch.empty(n).uniform_(-1.5, 1.5, generator=g).double()
    S = 2 + torch.exp(u1.float().double())
    m1 = torch.sigmoid(u0.float().double())
    var = torch.distributions.Beta(S * m1, S * (1 - m1)).sample().float().double()
    # a slice of the batch sits in / next to the Taylor patch around the mean, where ATen's form cancels worst
    k = n // 5
    sd = torch.sqrt(m1 * (1 - m1) / (S + 1))
    var[:k] = (m1[:k] + sd[:k] * torch.empty(k).uniform_(-0.3, 0.3, generator=g).double()).float().double()
    z = np.zeros(n)
    ref, fast, status = _run(hc, mc, site, u0.numpy(), u1.numpy(), z, z, var.numpy())
    # the plain-fp32 tier (sizes below ~22) keeps the textbook cancellation in d sample / d size (A1): 1e-5, the north-star bound
    _check(ref, fast, status, 7, min_fast, tol=1e-5 if lsize[1] <= 3.0 else 5e-6)


def test_out_of_regime_falls_back():
    """Tiny concentrations, samples on the clamps and fp32 reference conventions are left to the double form."""
    hc = hostcheck.load()
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
    buf = (ctypes.c_double * NOUT)()
    call = lambda s, a, b, v, m=mc: hc.hc_site_eval_fast(s, ctypes.c_float(a), ctypes.c_float(b), ctypes.c_float(0), ctypes.c_float(0),
                                                         ctypes.byref(m), ctypes.c_double(v), buf)
    assert call(5, 0.0, 1.0, 0.5) == 0          # Beta with c1 = c0 = 2.4: the small-concentration fp32 tier
    assert call(5, 0.0, 6.0, 1e-9) != 0         # far tail of a concentrated Beta (size 405 > 64)
    assert call(5, 0.0, 1.0, 1e-35) != 0        # draw that underflows fp32
    assert call(1, -9.0, 2.0, 1.0) != 0         # Gamma concentration < e^-4
    assert call(5, 0.0, 6.0, 0.5) == 0
    mc32 = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float32)
    assert call(5, 0.0, 6.0, 0.5, mc32) != 0


def test_rng_draws_through_fast_path_have_the_right_moments():
    hc = hostcheck.load()
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
    n = 40000
    out = np.zeros(n)
    p = out.ctypes.data_as(ctypes.c_void_p)
    # height: Gamma(loc * beta, beta) with loc = 3000, beta = 0.02
    hc.hc_site_draws_fast(1, ctypes.c_float(np.log(3000.0)), ctypes.c_float(np.log(0.02)), ctypes.byref(mc), ctypes.c_uint64(11), n, p)
    assert abs(out.mean() - 3000.0) < 5 * np.sqrt(3000.0 / 0.02 / n)
    assert abs(out.var() / (3000.0 / 0.02) - 1.0) < 0.05
    # x: AffineBeta(mean 1.5, size 400) on [-7.5, 7.5]
    m1 = (1.5 + 7.5) / 15.0
    hc.hc_site_draws_fast(5, ctypes.c_float(np.log(m1 / (1 - m1))), ctypes.c_float(np.log(398.0)), ctypes.byref(mc), ctypes.c_uint64(12), n, p)
    var = 15.0 ** 2 * m1 * (1 - m1) / 401.0
    assert abs(out.mean() - 1.5) < 5 * np.sqrt(var / n)
    assert abs(out.var() / var - 1.0) < 0.05


@pytest.mark.parametrize("site", [3, 6])   # width, x
@pytest.mark.parametrize("regime", ["tails", "tails_large", "lopsided", "lopsided_far", "small_series"])
def test_beta_sites_outside_the_bulk(site, regime):
    """What a TRAINED model sends outside the reformulated bulk regime (profiles/site_regimes.py): draws in the tails of
    moderately concentrated guides (series in double on one side, Rice on the other), guides pushed against an edge of
    their interval (one concentration <= 6 with total > 64) and the small-x series with beta x >= 2."""
    hc = hostcheck.load()
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
    seed = {"tails": 1, "tails_large": 2, "lopsided": 3, "lopsided_far": 4, "small_series": 5}[regime] * 10 + site
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed + 1000)
    n = 3000
    U = lambda lo, hi: torch.empty(n).uniform_(lo, hi, generator=g).double()
    if regime in ("tails", "tails_large"):
        u1 = U(2.5, 4.1) if regime == "tails" else U(4.2, 8.0)
        u0 = U(-1.5, 1.5)
        widen = 0.12 if regime == "tails" else 0.03          # draws from a much flatter Beta: mostly tails of the guide
    elif regime in ("lopsided", "lopsided_far"):
        u1 = U(2.7, 9.0) if regime == "lopsided" else U(4.2, 9.0)
        u0 = (U(2.0, 5.0) if regime == "lopsided" else U(5.0, 9.0)) * torch.where(U(0, 1) < 0.5, -1.0, 1.0)
        widen = 1.0
    else:
        u1 = U(0.5, 4.1)
        u0 = U(-2.5, 2.5)
        widen = 1.0
    S = 2 + torch.exp(u1.float().double())
    m1 = torch.sigmoid(u0.float().double())
    var = torch.distributions.Beta(widen * S * m1, widen * S * (1 - m1)).sample().float().double().clamp(1e-7, 1 - 1e-7)   # closer to a bound the float sample is clamped anyway
    if regime == "tails_large":
        t = U(0.05, 2.6) / S                                 # total x (1 - x) around / below ATen's 2.5 and 0.75 boundaries
        var = torch.where(U(0, 1) < 0.5, t, 1 - t).float().double()
    z = np.zeros(n)
    ref, fast, status = _run(hc, mc, site, u0.numpy(), u1.numpy(), z, z, var.numpy())
    x, c1, c0 = var.numpy(), (S * m1).numpy(), (S * (1 - m1)).numpy()
    bulk = (c1 > 6) & (c0 > 6) & (S.numpy() * x * (1 - x) >= 2.5)
    assert (~bulk).mean() > 0.25, "the batch is meant to sit outside the bulk regime"
    ok = status == 0
    sane = np.minimum(c1, c0) > 0.3        # below that the draws underflow fp32 (x ~ u^(1/c)): the double form's business
    assert ok[sane].mean() >= 0.995, f"only {ok[sane].mean():.3f} of the sites took the fp32 path"
    # entry by entry here (a tail draw's log q is tens of times the typical one), relative to the entry or -- for entries
    # that pass through zero -- to the batch's typical size: 1e-5, the north-star bound.  Three outputs are differences
    # that cancel in ANY fp32 evaluation from fp32 parameters (d log q / d sample ~ (c1 - 1)/x - (c0 - 1)/y at c1 -> 1;
    # d sample / d size = m1 dx/dc1 + m0 dx/dc0, ten- to thirty-fold for lopsided guides; likewise d log q / d size):
    # for those, 1e-5 on all but a few percent of the entries and 5e-5 on the rest.
    bad = {}
    keep = ok & ~bulk & sane
    for j in range(7):
        r, f = ref[keep, j], fast[keep, j]
        err = np.abs(f - r) / np.maximum(np.abs(r), np.median(np.abs(r)))
        if NAMES[j] == "LQ":
            # a sum of lgamma-sized terms (tens) that feeds only the ELBO value, where it is one of ~10 per unit next to
            # a likelihood term of order -1000: judged against that sum's scale
            err = np.abs(f - r) / np.maximum(np.abs(r), 10.0)
        loose = NAMES[j] in ("DQ", "A1", "B1")
        if not (err.max() < (5e-5 if loose else 1e-5) and np.mean(err > 1e-5) < 0.03):
            i = err.argmax()
            bad[NAMES[j]] = (err.max(), float(np.mean(err > 1e-5)), r[i], f[i], c1[keep][i], c0[keep][i], x[keep][i])
    assert not bad, bad


def test_two_pass_form_equals_the_one_call_form():
    """site_fast_kernel evaluates the bulk regime in a first pass and replays the other sites, compacted by regime class, in
    a second (site_eval_fast_t<1> / <2>): same draws, same outputs as the one-call form, and every class is exercised."""
    hc = hostcheck.load()
    mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
    hc.hc_site_eval_fast_split.restype = ctypes.c_int
    hc.hc_site_eval_fast_rng.restype = ctypes.c_int
    g = np.random.default_rng(0)
    a, b = (ctypes.c_double * NOUT)(), (ctypes.c_double * NOUT)()
    va, vb, cls = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    seen = {"gamma": set(), "beta": set()}
    for i in range(24000):
        s = int(g.integers(0, 9))
        if s < 3:
            lc = g.uniform(-1.0, 9.0); u1 = g.uniform(-7, 1); u0 = lc - u1
            ubm = u0 + 0.1 * g.normal(); ubs = ubm - g.uniform(0.5, 3.0)
        else:
            u0 = g.uniform(-6, 6) if i % 3 == 0 else g.uniform(-1.5, 1.5); u1 = g.uniform(-1, 9); ubm = ubs = 0.0
        args = (s, ctypes.c_float(u0), ctypes.c_float(u1), ctypes.c_float(ubm), ctypes.c_float(ubs), ctypes.byref(mc), ctypes.c_uint64(i))
        st2 = hc.hc_site_eval_fast_split(*args, a, ctypes.byref(cls), ctypes.byref(va))
        st1 = hc.hc_site_eval_fast_rng(*args, b, ctypes.byref(vb))
        assert st1 == st2 and va.value == vb.value, (i, s, st1, st2)
        if st1 == 0:
            assert list(a) == list(b), (i, s, cls.value, list(a), list(b))
        seen["gamma" if s < 3 else "beta"].add(cls.value)
    # classes (cosmos_sites_fast.cuh): Gamma = bits (x < 0.8, conc > 8, conc > 10); Beta = density regime * 16 + the
    # branches of the two beta_grad_tierb calls -- every density regime and every branch must have been replayed
    assert seen["gamma"] >= {-1, 0, 1, 2}, seen   # (x < 0.8 at concentration > 8 is a far-tail draw: not in 24000 tries)
    beta = {c for c in seen["beta"] if c >= 0}
    assert {c // 16 for c in beta} == {0, 1, 2, 3} and {(c // 4) % 4 for c in beta} == {0, 1, 2, 3} and -1 in seen["beta"], seen
    assert max(beta) < 64
