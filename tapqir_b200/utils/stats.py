"""
Post-fit statistics with the reference's outputs (tapqir/utils/stats.py:29-293, row N2):
credible intervals of every guide distribution, SNR and chi2 per patch, classification metrics
against simulation labels, and the files ``<name>_params.tpqr`` / ``.mat`` / ``<name>_summary.csv``.

The two heavy parts run on the device (csrc/stats.cu): the inverse CDFs behind the credible intervals over whole
(K, Nt, F, Q) arrays, and SNR / chi2 in one fused launch over the resident pixels; the classification metrics against
simulation labels stay scipy / sklearn on a few thousand labels.  The rastergram PNGs of
stats.py:110-128 need matplotlib (not a dependency here) and are skipped, as the reference does
when the ``CI`` environment variable is set.
"""

import logging
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

logger = logging.getLogger(__name__)


def quantile(samples: torch.Tensor, q: float) -> torch.Tensor:
    """pyro.ops.stats.quantile [third party]: linear interpolation between order statistics."""
    s, _ = samples.flatten().double().sort()
    pos = q * (s.numel() - 1)
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    return s[lo] + (s[hi] - s[lo]) * (pos - lo)


def hpdi(samples: torch.Tensor, prob: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """pyro.ops.stats.hpdi [third party]: narrowest interval holding ``prob`` of the samples."""
    s, _ = samples.flatten().double().sort()
    n = s.numel()
    mass = int(prob * n)
    widths = s[mass:] - s[: n - mass]
    i = int(torch.argmin(widths))
    return s[i], s[i + mass]


def guide_family(name, value, P, priors):
    """
    ``(family, p1, p2, low, scale, mean)`` of the guide of the latent ``name`` (``model.ci_params``) from the constrained
    variational parameters ``value(param_name) -> double tensor``: family "gamma" (p1 = concentration, p2 = rate) or
    "beta" (p1, p2 = concentrations on [0, 1]; the interval is ``low + scale * quantile``).  Reference: the branches of
    ``compute_params`` (cosmos.py:713-772) with ``torch_to_scipy_dist`` (stats.py:262-293): Gamma(loc * beta, rate beta);
    Dirichlet(mean * size) summarised by its Beta marginals (``pi``; hmm's ``init`` and ``trans``); AffineBeta as a
    located and scaled Beta.
    """
    import math

    half = (P + 1) / 2

    def gamma(loc, beta):
        return "gamma", loc * beta, beta, 0.0, 1.0, loc

    def affine_beta(mean, size, lo, hi):
        return "beta", size * (mean - lo) / (hi - lo), size * (hi - mean) / (hi - lo), lo, hi - lo, mean

    if name in ("pi", "init", "trans"):
        conc = value(f"{name}_mean") * value(f"{name}_size")
        total = conc.sum(-1, keepdim=True)
        return "beta", conc, total - conc, 0.0, 1.0, conc / total
    if name == "gain":
        return gamma(value("gain_loc"), value("gain_beta"))
    if name == "lamda":
        return gamma(value("lamda_loc"), value("lamda_beta"))
    if name == "proximity":
        return affine_beta(value("proximity_loc"), value("proximity_size"), 0.0, (P + 1) / math.sqrt(12))
    if name == "background":
        return gamma(value("b_loc"), value("b_beta"))
    if name == "height":
        return gamma(value("h_loc"), value("h_beta"))
    if name == "width":
        return affine_beta(value("w_mean"), value("w_size"), priors["width_min"], priors["width_max"])
    if name == "x":
        return affine_beta(value("x_mean"), value("size"), -half, half)
    if name == "y":
        return affine_beta(value("y_mean"), value("size"), -half, half)
    raise NotImplementedError(f"no credible interval for '{name}'")


def credible_intervals(ci_params, value, P, priors, CI, device):
    """
    ``{name: {"LL", "UL", "Mean"}}`` (CPU double tensors) for every latent in ``ci_params`` (cosmos.py:773-778): the
    inverse CDFs of the Gamma / Beta guides evaluated on ``device`` over whole parameter arrays
    (``tq_gamma_interval`` / ``tq_beta_interval``, csrc/stats.cu) -- the reference goes through scipy element by element
    on the CPU (stats.py:262-293), 90 M inverse incomplete gamma / beta evaluations at 1000 AOIs x 5000 frames.
    """
    from tapqir_b200 import _lib

    lib = _lib.load()
    device = torch.device(device)
    out = {}
    with torch.cuda.device(device):
        st = _lib.stream_ptr(device)
        for name in ci_params:
            family, p1, p2, low, scale, mean = guide_family(name, value, P, priors)
            shape = p1.shape
            a = p1.to(device=device, dtype=torch.float64).reshape(-1).contiguous()
            b = p2.to(device=device, dtype=torch.float64).expand(shape).reshape(-1).contiguous()
            lo, hi = torch.empty_like(a), torch.empty_like(a)
            fn = lib.tq_gamma_interval if family == "gamma" else lib.tq_beta_interval
            _lib.check(fn(a.numel(), _lib.ptr(a), _lib.ptr(b), float(CI), _lib.ptr(lo), _lib.ptr(hi), st), f"interval({name})")
            out[name] = {"LL": (low + scale * lo).reshape(shape).cpu(), "UL": (low + scale * hi).reshape(shape).cpu(),
                         "Mean": mean.detach().double().cpu()}
    return out


def snr_and_chi2(data, height, width, x, y, target_locs, background, gain, offset_mean, offset_var, P,
                 theta_probs) -> Tuple[torch.Tensor, torch.Tensor]:
    r"""
    Signal-to-noise ratio and chi2 of the fitted spots (reference: stats.py:29-86), one fused kernel over all patches
    (``tq_snr_chi2``: separable spot factors, no (K, n, F, Q, P, P) temporaries).

    .. math:: \text{SNR}_{knf} = \frac{\sum_{ij} (D_{nfij} - b_{nf} - \mu_{offset})\,\mathcal N_{ij}}
              {\sqrt{\sigma^2_{offset} + b_{nf}\, g}}

    ``data``: CUDA pixels ``(..., P, P)``, uint16 or float32 (the device store's own buffer works); spot parameters
    ``(K, ...)`` and ``background (...)`` CUDA tensors, spots stacked along the FIRST axis; ``target_locs (..., 2)``.
    Returns ``snr (K, ...)`` and ``chi2 (...)`` as float32 CUDA tensors.
    """
    from tapqir_b200 import _lib

    lib = _lib.load()
    dev = data.device
    batch = tuple(background.shape)
    U = int(np.prod(batch)) if batch else 1
    K = height.shape[0]
    if K != 2:
        raise NotImplementedError("tq_snr_chi2 is built for K = 2")
    if data.dtype not in (torch.uint16, torch.float32):
        data = data.to(torch.float32)
    spot = lambda t: t.to(device=dev, dtype=torch.float32).reshape(K, U).contiguous()
    pix = data.reshape(U, P, P).contiguous()
    xy = target_locs.to(device=dev, dtype=torch.float32).reshape(U, 2).contiguous()
    b = background.to(device=dev, dtype=torch.float32).reshape(U).contiguous()
    h, w, xs, ys = spot(height), spot(width), spot(x), spot(y)
    snr = torch.empty(K, U, dtype=torch.float32, device=dev)
    chi2 = torch.empty(U, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.tq_snr_chi2(U, int(P), _lib.pix_code(pix.dtype), _lib.ptr(pix), _lib.ptr(xy), _lib.ptr(h), _lib.ptr(w),
                                   _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(b), float(gain), float(offset_mean), float(offset_var),
                                   _lib.ptr(snr), _lib.ptr(chi2), _lib.stream_ptr(dev)), "tq_snr_chi2")
    return snr.reshape((K,) + batch), chi2.reshape(batch)


def save_stats(model, path, CI=0.95, save_matlab=False):
    """Reference: stats.py:89-259."""
    import pandas as pd
    from sklearn.metrics import confusion_matrix, matthews_corrcoef, precision_score, recall_score

    global_params = model._global_params
    ll, ul = f"{int(100 * CI)}% LL", f"{int(100 * CI)}% UL"
    summary = pd.DataFrame(index=global_params, columns=["Mean", ll, ul], dtype=object)
    logger.info("- credible intervals & spot probabilities")
    ci_stats = model.compute_params(CI)
    for param in global_params:
        scalar = ci_stats[param]["Mean"].ndim == 0
        for col, key in (("Mean", "Mean"), (ll, "LL"), (ul, "UL")):
            summary.loc[param, col] = ci_stats[param][key].item() if scalar else ci_stats[param][key].tolist()

    theta_mask = ci_stats["theta_probs"] > 0.5
    hmax = np.percentile(ci_stats["height"]["Mean"][theta_mask], 99) if theta_mask.sum() else 1
    ci_stats["height"]["vmin"], ci_stats["height"]["vmax"] = -0.03 * hmax, 1.3 * hmax
    ci_stats["width"]["vmin"], ci_stats["width"]["vmax"] = 0.5, 2.5
    ci_stats["x"]["vmin"], ci_stats["x"]["vmax"] = -9, 9
    ci_stats["y"]["vmin"], ci_stats["y"]["vmax"] = -9, 9
    bmax = np.percentile(ci_stats["background"]["Mean"].flatten(), 99)
    ci_stats["background"]["vmin"], ci_stats["background"]["vmax"] = -0.03 * bmax, 1.3 * bmax
    if model.data.time1 is not None:
        ci_stats["time1"] = model.data.time1
    if model.data.ttb is not None:
        ci_stats["ttb"] = model.data.ttb
    model.params = ci_stats

    logger.info("- SNR and Chi2-test")
    data, dev = model.data, model.device
    Nt, F, Q, K = data.Nt, data.F, model.Q, model.K
    # one launch over the pixels already resident on the device (the reference loops over AOIs in Python, stats.py:193-215)
    store = model.engine.store if model.engine is not None else data.device_store(dev, torch.float32)
    to = lambda t: torch.as_tensor(t).to(device=dev, dtype=torch.float32)
    snr, chi2 = snr_and_chi2(
        store.pixels, to(ci_stats["height"]["Mean"]), to(ci_stats["width"]["Mean"]), to(ci_stats["x"]["Mean"]),
        to(ci_stats["y"]["Mean"]), store.xy, to(ci_stats["background"]["Mean"]), float(ci_stats["gain"]["Mean"]),
        data.offset.mean, data.offset.var, data.P, None)
    snr, chi2 = snr.cpu(), chi2.cpu()
    for q in range(Q):
        masked = snr[..., q][ci_stats["theta_probs"][..., q] > 0.5]
        summary.loc[f"SNR_{q}", "Mean"] = masked.mean().item() if masked.numel() else float("nan")
    cmax = quantile(chi2, 0.99).item()
    ci_stats["chi2"] = {"values": chi2, "vmin": -0.03 * cmax, "vmax": 1.3 * cmax}

    if data.labels is not None:
        pred = model.z_map[data.is_ontarget].cpu().numpy().ravel()
        true = data.labels["z"][: data.N].ravel()
        with np.errstate(divide="ignore", invalid="ignore"):
            summary.loc["MCC", "Mean"] = matthews_corrcoef(true, pred)
        summary.loc["Recall", "Mean"] = recall_score(true, pred, zero_division=0)
        summary.loc["Precision", "Mean"] = precision_score(true, pred, zero_division=0)
        (summary.loc["TN", "Mean"], summary.loc["FP", "Mean"], summary.loc["FN", "Mean"],
         summary.loc["TP", "Mean"]) = confusion_matrix(true, pred, labels=(0, 1)).ravel()
        mask = torch.from_numpy(data.labels["z"][: data.N]) > 0
        samples = torch.masked_select(model.z_probs[data.is_ontarget].argmax(dim=-1).cpu(), mask)
        if len(samples):
            z_ll, z_ul = hpdi(samples, CI)
            summary.loc["p(specific)", "Mean"] = quantile(samples, 0.5).item()
            summary.loc["p(specific)", ll], summary.loc["p(specific)", ul] = z_ll.item(), z_ul.item()
        else:
            summary.loc["p(specific)", ["Mean", ll, ul]] = 0.0
    model.summary = summary

    if path is not None:
        path = Path(path)
        torch.save(ci_stats, path / f"{model.name}_params.tpqr")
        logger.info(f"Parameters were saved in {path / f'{model.name}_params.tpqr'}")
        if save_matlab:
            from scipy.io import savemat

            mat = {}
            for param, field in ci_stats.items():
                if isinstance(field, dict):
                    mat[param] = {k: np.asarray(v) for k, v in field.items()}
                else:
                    mat[param] = np.asarray(field)
            savemat(path / f"{model.name}_params.mat", mat)
        summary.to_csv(path / f"{model.name}_summary.csv")
        logger.info(f"Summary statistics were saved in {path / f'{model.name}_summary.csv'}")
