"""
Generates tests/golden/ref_*.pt by running the REFERENCE's own code in the build container.

Run as:  python tests/golden/make_golden.py   (needs /root/reference; does not run on the GPU box)

What can be imported from the reference here:
* tapqir/distributions/util.py      -- needs only torch: imported verbatim.
* tapqir/distributions/ksmogn.py    -- imports pykeops and pyro.distributions.TorchDistribution at
  module scope (ksmogn.py:9-12).  Both packages are absent and not installable, so the two import
  names are satisfied by empty stand-ins (a ``Genred`` placeholder that is never called because
  ``use_pykeops=False``; ``TorchDistribution = torch.distributions.Distribution``).  The arithmetic
  executed is the reference's own torch branch (ksmogn.py:222-238) and util.gaussian_spots.
Nothing from the reference is copied into the repository: only the numeric outputs are stored.
"""

import importlib.util
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_reference_modules():
    pk = types.ModuleType("pykeops")
    pk.set_verbose = lambda *a, **k: None
    pkt = types.ModuleType("pykeops.torch")
    pkt.Genred = object
    pyro = types.ModuleType("pyro")
    pdist = types.ModuleType("pyro.distributions")

    class TorchDistribution(torch.distributions.Distribution):
        pass

    pdist.TorchDistribution = TorchDistribution
    sys.modules.update({"pykeops": pk, "pykeops.torch": pkt, "pyro": pyro, "pyro.distributions": pdist})
    for name in ("tapqir", "tapqir.distributions"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    util = load("tapqir.distributions.util", REF / "tapqir/distributions/util.py")
    ksmogn = load("tapqir.distributions.ksmogn", REF / "tapqir/distributions/ksmogn.py")
    return util, ksmogn


def make_case(seed, N, F, C, K, P, O, gain, dtype=torch.float64):
    """Random but realistic KSMOGN inputs (test-suite constants of test/test_tapqir.py:25-40)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, dtype=dtype)
    case = dict(
        height=500 + 4000 * r(N, F, C, K),
        width=0.8 + 1.4 * r(N, F, C, K),
        x=-6 + 12 * r(N, F, C, K),
        y=-6 + 12 * r(N, F, C, K),
        target_locs=(P - 1) / 2 + (r(N, F, C, 2) - 0.5),
        background=100 + 100 * r(N, F, C),
        gain=torch.tensor(gain, dtype=dtype),
        P=P,
    )
    if O == 3:  # simulate.py:92,103 -- three identical bins
        case["offset_samples"] = torch.full((3,), 90.0, dtype=dtype)
        case["offset_weights"] = torch.ones(3, dtype=dtype) / 3
    else:  # an empirical-looking histogram of integer offsets
        case["offset_samples"] = torch.arange(80, 80 + O, dtype=dtype)
        w = torch.exp(-0.5 * ((case["offset_samples"] - 90) / 4) ** 2) + 1e-3
        case["offset_weights"] = w / w.sum()
    # integer-valued pixels as produced by simulate.py:122 (floor) and glimpse (int)
    mean = case["background"][..., None, None] + 60 * r(N, F, C, P, P)
    case["value"] = torch.floor(mean + 30 * torch.randn(N, F, C, P, P, generator=g, dtype=dtype)) + 90
    if O != 3:
        # a pixel above some offsets and at/below others: exercises the D > delta mask.  (A pixel
        # at/below EVERY offset gives -inf; ingestion prevents it, glimpse_reader.py:407-411.)
        case["value"][0, 0, 0, 0, 0] = case["offset_samples"][O // 2]
    return case


def main():
    torch.set_default_dtype(torch.float64)
    util, ksmogn = load_reference_modules()
    out = {}

    # ---- prior tables (util.py:67-173) ----
    tables = {}
    for K in (2, 3):
        for lam in ([0.15], [0.5, 0.01], [2.5]):
            lt = torch.tensor(lam)
            tables[f"probs_m_K{K}_lam{lam}"] = dict(lamda=lt, K=K, probs_m=util.probs_m(lt, K),
                                                     trunc=util.truncated_poisson_probs(lt, K))
        tables[f"probs_theta_K{K}"] = util.probs_theta(K, torch.device("cpu")).clone()
    pi = torch.tensor([[0.85, 0.15], [0.3, 0.7]])
    tables["expand_offtarget"] = dict(probs=pi, out=util.expand_offtarget(pi))
    out["tables"] = tables

    # ---- gaussian_spots + KSMOGN.log_prob (torch branch) ----
    cases = {}
    for name, kw in {
        "sim_O3": dict(seed=1, N=3, F=4, C=1, K=2, P=14, O=3, gain=7.0),
        "hist_O16_C2": dict(seed=2, N=2, F=3, C=2, K=2, P=14, O=16, gain=11.5),
        "small_P6": dict(seed=3, N=1, F=2, C=1, K=2, P=6, O=5, gain=3.0),
    }.items():
        c = make_case(**kw)
        K = kw["K"]
        m = torch.tensor([[(i >> k) & 1 for k in range(K)] for i in range(2**K)], dtype=torch.float64)
        m_enum = m.reshape(2, 2, 1, 1, 1, K)  # (m_1, m_0, nb, fb, C, K) as enumerated by Pyro
        spots = util.gaussian_spots(c["height"], c["width"], c["x"], c["y"],
                                    c["target_locs"].unsqueeze(-2), c["P"])
        spots_m = util.gaussian_spots(c["height"], c["width"], c["x"], c["y"],
                                      c["target_locs"].unsqueeze(-2), c["P"], m_enum)
        grads = {}
        leaves = {k: c[k].clone().requires_grad_(True) for k in ("height", "width", "x", "y", "background", "gain")}
        dist = ksmogn.KSMOGN(leaves["height"], leaves["width"], leaves["x"], leaves["y"], c["target_locs"],
                             leaves["background"], leaves["gain"], c["offset_samples"],
                             torch.distributions.utils.probs_to_logits(c["offset_weights"]),
                             c["P"], m_enum, use_pykeops=False)
        logp = dist.log_prob(c["value"])  # (2,2,N,F,C)
        # a fixed upstream weight so the backward is a deterministic function of the inputs
        gw = torch.Generator().manual_seed(100 + kw["seed"])
        W = torch.rand(logp.shape, generator=gw, dtype=torch.float64)
        (W * logp).sum().backward()
        grads = {k: v.grad.clone() for k, v in leaves.items()}
        nom = ksmogn.KSMOGN(c["height"], c["width"], c["x"], c["y"], c["target_locs"], c["background"],
                            c["gain"], c["offset_samples"],
                            torch.distributions.utils.probs_to_logits(c["offset_weights"]), c["P"],
                            None, use_pykeops=False)
        cases[name] = dict(inputs=c, m=m, spots=spots, spots_m=spots_m, image=dist.image.detach(),
                           log_prob=logp.detach().reshape((2**K,) + logp.shape[2:]),
                           W=W.reshape((2**K,) + logp.shape[2:]), grads=grads,
                           log_prob_no_m=nom.log_prob(c["value"]))
    out["ksmogn"] = cases
    torch.save(out, OUT / "ref_distributions.pt")
    print("wrote", OUT / "ref_distributions.pt")


if __name__ == "__main__":
    main()
