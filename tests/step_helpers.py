"""Shared by the step parity tests: builds a small simulated problem, oracle params/noise, and the
flat buffers the kernels (or the host emulation) take."""

import ctypes

import torch

from oracle import cosmos_oracle as O
from tapqir_b200.models import layout as L
from tapqir_b200.utils.simulate import simulate


def make_problem(N=4, F=6, C=1, nb=3, fb=4, seed=0, perturb=True, offsets="sim", dtype=torch.float64, P=14):
    kw = {} if P == 14 else {"P": P}   # AOI size (the reference's `--aoi-size`; 14 is the default of glimpse and simulate)
    if offsets == "hist":
        s = torch.arange(80.0, 96.0)
        w = torch.exp(-0.5 * ((s - 90) / 3) ** 2) + 1e-3
        kw = dict(offset_samples=s, offset_weights=w / w.sum())
    ds = simulate(N, F, C=C, seed=seed, **kw)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights, dtype=dtype)
    g = torch.Generator().manual_seed(seed + 100)
    params = O.to_unconstrained(O.init_constrained(data), data.P, data.dtype)
    if perturb:  # move away from the symmetric initial point so every gradient path is exercised
        for k, v in params.items():
            v.add_(0.3 * torch.randn(v.shape, generator=g, dtype=v.dtype))
    ndx = torch.randperm(N, generator=g)[:nb]
    fdx = torch.randperm(F, generator=g)[:fb]
    noise = O.draw_noise(params, data, ndx, fdx, g)
    return ds, data, params, ndx, fdx, noise


def flat_inputs(data, params, noise, dtype):
    ll, gl = L.LocalLayout(data.Nt, data.F, data.C), L.GlobalLayout(data.C)
    lparams = ll.pack({k: params[k] for k in L.LOCAL_NAMES}, dtype=dtype)
    gparams = gl.pack({k: params[k] for k in L.GLOBAL_NAMES}, dtype=torch.float64)
    lnoise = L.pack_local_noise(noise, dtype, "cpu")
    gnoise = gl.pack_noise(noise)
    return ll, gl, lparams, gparams, lnoise, gnoise


def host_step(hc, data, params, ndx, fdx, noise, dtype=torch.float64, priors=O.DEFAULT_PRIORS, ref_dtype=torch.float64):
    """Run tests/hostcheck's CPU emulation of the kernel pipeline; returns loss and named grads."""
    ll, gl, lparams, gparams, lnoise, gnoise = flat_inputs(data, params, noise, dtype)
    mc = L.ModelConst.make(priors, data.P, ref_dtype)
    assert hc.hc_sizeof_model_const() == ctypes.sizeof(mc)
    nb, fb = len(ndx), len(fdx)
    U = nb * fb * data.C
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    ndx32, fdx32 = ndx.to(torch.int32).contiguous(), fdx.to(torch.int32).contiguous()
    pixels = data.images.to(dtype).contiguous()
    xy = data.xy.to(dtype).contiguous()
    ont, mask = data.is_ontarget.to(torch.uint8).contiguous(), data.mask.to(torch.uint8).contiguous()
    off_s, off_w = data.offset_samples.to(dtype).contiguous(), data.offset_logits.to(dtype).contiguous()
    lgrads = torch.empty_like(lparams)
    ggrads = torch.empty(gl.numel, dtype=torch.float64)
    acc = torch.empty(data.C * L.NACC, dtype=torch.float64)
    samples = torch.empty(L.NSAMP, U, dtype=dtype)
    fn = hc.hc_cosmos_step_f64 if dtype == torch.float64 else hc.hc_cosmos_step_f32
    fn.restype = ctypes.c_double
    loss = fn(nb, fb, data.Nt, data.F, data.C, data.P, off_s.numel(), p(ndx32), p(fdx32), p(pixels), p(xy), p(ont), p(mask),
              p(off_s), p(off_w), ctypes.byref(mc), ctypes.c_double(data.Nt / nb), ctypes.c_double(data.F / fb),
              p(lparams), p(gparams), p(lnoise), p(gnoise), p(lgrads), p(ggrads), p(acc), p(samples))
    grads = dict(ll.views(lgrads))
    grads.update(gl.views(ggrads))
    return loss, grads, samples


def compare_grads(ours, ref, tol, names=None):
    """max |ours - ref| / max |ref| per parameter; returns the worst offenders for the message."""
    bad = {}
    for k in names or ref.keys():
        r = ref[k].double()
        denom = r.abs().max().item()
        err = (ours[k].double().cpu().reshape(r.shape) - r).abs().max().item()
        rel = err / denom if denom > 0 else err
        if not rel < tol:
            bad[k] = rel
    return bad


def golden_step_case(name):
    """A case of tests/golden/ref_step.pt (SVI iterations of the reference's own cosmos.py / model.py, produced by
    tests/golden/make_golden_step.py): returns (dataset, oracle data view, case dict)."""
    from pathlib import Path

    from tapqir_b200.utils.dataset import CosmosDataset

    case = torch.load(Path(__file__).resolve().parent / "golden" / "ref_step.pt", weights_only=False)[name]
    ds = CosmosDataset(case["images"].to(torch.float32), case["xy"], case["is_ontarget"], case["mask"].clone(), None,
                       case["offset_samples"], case["offset_weights"])
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    return ds, data, case


def masked_loss_constant(case, step):
    """What the reference's reported loss lacks relative to ours for masked AOIs in the minibatch: Pyro still sums over
    the values of enumerated sites whose log-probabilities the mask zeroed, 4 configurations x ln 6 states per unit
    (parameter-independent; tests/test_oracle.py)."""
    import math

    cfg = case["config"]
    n_masked = int((~case["mask"][step["ndx"]]).sum())
    return n_masked * cfg["fb"] * cfg["C"] * 4 * math.log(6) * (cfg["N"] / cfg["nb"]) * (cfg["F"] / cfg["fb"])


GLOBAL_PAIRS = [("gain_loc", "gain_beta"), ("lamda_loc", "lamda_beta"), ("proximity_loc", "proximity_size"), ("pi_mean", "pi_size")]


def compare_global_grads(ours, ref, tol, running_scale=None):
    """Global gradients are one to four numbers each; measured against themselves they have no scale when they pass
    through zero.  The two parameters of one guide distribution (mean-like, concentration-like) share their terms: the
    gradient of the concentration-like one is the difference of two terms of the mean-like one's size (moving it keeps
    the mean), e.g. 58 against 4472 for ``gain`` on the test data.  Errors are therefore measured against the largest
    entry of the PAIR -- the conditioning of the quantity, as the largest entry of a tensor is for the local ones."""
    bad = {}
    for pair in GLOBAL_PAIRS:
        scale = max(ref[k].double().abs().max().item() for k in pair)
        if running_scale is not None:
            # along a fit the global gradients go to zero (gain: 118 k at the first iteration, 24 k at the 91st) while
            # the fp32 error of the sums they are made of stays put (0.2-0.5 for gain at C1); what Adam divides by is the
            # running scale of the gradient, so that is what the error is measured against
            scale = running_scale[pair] = max(scale, running_scale.get(pair, 0.0))
        for k in pair:
            r = ref[k].double()
            err = (ours[k].double().cpu().reshape(r.shape) - r).abs().max().item()
            if not err <= tol * scale:
                bad[k] = err / scale
    return bad


def golden_c1_fit():
    """tests/golden/ref_c1_fit.pt: BASELINE configs[0] (N=5 x F=100, full batch, 100 SVI iterations) run by the reference's
    own cosmos.py / model.py (tests/golden/make_golden_step.py::run_c1_fit).  Only losses and final parameters are stored:
    with a full batch the guide's variates are the only consumers of torch's random stream, in the guide's site order, so
    ``torch.manual_seed(rng_seed)`` + ``oracle.draw_noise`` at the current parameters reproduces every draw."""
    from pathlib import Path

    from tapqir_b200.utils.dataset import CosmosDataset

    case = torch.load(Path(__file__).resolve().parent / "golden" / "ref_c1_fit.pt", weights_only=False)
    ds = CosmosDataset(case["images"].to(torch.float32), case["xy"], case["is_ontarget"], case["mask"].clone(), None,
                       case["offset_samples"], case["offset_weights"])
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    return ds, data, case


def global_grad_conditioning(params, data, ndx, fdx, noise, **kw):
    """
    Forward-error scale of every global gradient entry, from the fp64 oracle.  With s the guide samples of the global
    sites (gain, pi, lamda, proximity):

        d loss / d theta_i = sum_s J_is G_s + e_i,     G_s = d loss / d s = sum_units g_us + r_s

    ``J`` (pathwise d sample / d theta), ``e`` (explicit part) and ``r`` (prior / guide terms of the sample itself) come
    from the fp64 global-site code; the per-unit contributions ``g_us`` -- for the gain, sums over the pixels of a
    patch of terms  img ln(y/img) + img - y + gain/2  whose expectation is ZERO at the true gain -- are what the fp32
    kernels deliver, and they carry both signs: as a fit converges G_gain goes to zero while its terms do not (C1 fit:
    118 k at iteration 1, 24 k at iteration 91, terms unchanged).  The scale a relative tolerance is measured against is
    therefore  sum_s |J_is| (sum over pixels / units of |terms of G_s| + |r_s|) + |e_i|,  and its ratio to
    |d loss / d theta_i| is the entry's condition number (1 = nothing cancels).  The terms are read off as
    d G_s / d pixel_weights and d G_s / d unit_weights (one double-backward pass per sample component).
    Returns name -> (scale tensor, worst condition number).
    """
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    uw = torch.ones(len(ndx), len(fdx), data.C, dtype=data.dtype, requires_grad=True)
    pw = torch.ones(len(ndx), len(fdx), data.C, data.P, data.P, dtype=data.dtype, requires_grad=True)
    total, parts = O.elbo(leaves, data, ndx, fdx, noise, return_parts=True, unit_weights=uw, pixel_weights=pw, **kw)
    loss = -total
    gl = [leaves[k] for k in L.GLOBAL_NAMES]
    samples = [parts[s] for s in ("gain", "pi", "lamda", "proximity")]
    G = torch.autograd.grad(loss, samples, create_graph=True)
    chain = [torch.zeros_like(t) for t in gl]
    chain_abs = [torch.zeros_like(t) for t in gl]
    for smp, Gs in zip(samples, G):
        for j in range(Gs.numel()):
            Gj = Gs.reshape(-1)[j]
            g_units, g_pix = torch.autograd.grad(Gj, [uw, pw], retain_graph=True, allow_unused=True)
            g_units = torch.zeros_like(uw) if g_units is None else g_units
            g_pix = torch.zeros_like(pw) if g_pix is None else g_pix
            # pixel terms, what a unit adds beyond its pixels, and what the sample's own prior / guide terms add
            g_abs = g_pix.abs().sum() + (g_units - g_pix.sum((-1, -2))).abs().sum() + (Gj.detach() - g_units.sum()).abs()
            Jj = torch.autograd.grad(smp.reshape(-1)[j], gl, retain_graph=True, allow_unused=True)
            for i, J in enumerate(Jj):
                if J is not None:
                    chain[i] += J * Gj.detach()
                    chain_abs[i] += J.abs() * g_abs
    full = torch.autograd.grad(loss, gl)
    out = {}
    for k, f, c, ca in zip(L.GLOBAL_NAMES, full, chain, chain_abs):
        scale = ca + (f - c).abs()
        out[k] = (scale, (scale / f.abs().clamp_min(1e-300)).max().item())
    return out


def compare_global_grads_conditioned(ours, ref, cond, tol):
    """|ours - ref| <= tol * (|J^T G| + |e|) entry by entry (see global_grad_conditioning); returns offenders as
    name -> (error / scale, error / |ref|, condition number)."""
    bad = {}
    for k, (scale, kappa) in cond.items():
        r = ref[k].double()
        err = (ours[k].double().cpu().reshape(r.shape) - r).abs()
        if not bool((err <= tol * scale.reshape(r.shape)).all()):
            bad[k] = ((err / scale.reshape(r.shape)).max().item(), (err / r.abs()).max().item(), kappa)
    return bad


def check_global_grads(ours, ref, params, data, ndx, fdx, noise, tol=1e-5, **kw):
    """North-star check of the global gradients: ``tol`` of each entry's forward-error scale
    (global_grad_conditioning; implies ``tol`` x condition number relative to the entry itself), and ``tol`` relative
    to the tensor's own largest entry wherever nothing cancels (condition number <= 1.5).  Returns the offenders
    (empty = pass)."""
    cond = global_grad_conditioning(params, data, ndx, fdx, noise, **kw)
    bad = {k: ("conditioned",) + v for k, v in compare_global_grads_conditioned(ours, ref, cond, tol).items()}
    well = [k for k, (_, kappa) in cond.items() if kappa <= 1.5]
    if well:
        bad.update({k: ("self-max", v) for k, v in compare_grads(ours, ref, tol, names=well).items()})
    return bad


def m_probs_grad_scale(params, data, ndx, fdx, noise, plate_sizes=None, **kw):
    """
    Forward-error scale of d loss / d m_probs (unconstrained, logit u_k):  s sum_m q(m) |m_k - p_k| |E(m)|  with E(m) the
    per-configuration ELBO term (likelihood ~ -1e3 + priors - log q) and s the plate scales x mask.  The gradient itself
    is  s sum_m q(m) (m_k - p_k) E(m): a q-weighted DIFFERENCE of the E(m), which goes to zero as the fit converges
    (p -> sigmoid of that difference) while the E(m) -- each delivered to the loss tolerance, 1e-6 -- do not.
    Returns a (K, nb, fb, C) tensor.
    """
    with torch.no_grad():
        _, parts = O.elbo(params, data, ndx, fdx, noise, return_parts=True, plate_sizes=plate_sizes, **kw)
        p = O.to_constrained(params, data.P, data.dtype)["m_probs"][:, ndx[:, None], fdx[None, :]]   # (K, nb, fb, C)
        Nt, F = plate_sizes if plate_sizes is not None else (data.Nt, data.F)
        s = (Nt / len(ndx)) * (F / len(fdx)) * data.mask[ndx].to(data.dtype)[:, None, None]
        mcfg = O.m_configs(p.shape[0], data.dtype)
        E, q = parts["per_config"].abs(), parts["q_m"]
        return torch.stack([s * sum(q[m] * (mcfg[m, k] - p[k]).abs() * E[m] for m in range(mcfg.shape[0]))
                            for k in range(p.shape[0])])
