"""
Generates tests/golden/ref_step.pt by running the REFERENCE's own model code in the build container.

Run as:  python tests/golden/make_golden_step.py    (needs /root/reference; never runs on the GPU box)

Executed verbatim from /root/reference (nothing is copied into the repository, only numeric outputs are stored):
  tapqir/models/cosmos.py        cosmos.__init__, init_parameters/_init_parameters (:464-598), guide (:329-462),
                                 model (:82-327), TraceELBO (:600-607)
  tapqir/models/hmm.py           hmm.__init__, init_parameters (:419-467), guide (:268-417), model (:82-266) in the sequential form
                                 ``vectorized=False`` (pyro.markov + TraceEnum_ELBO; the vectorised form needs funsor)
  tapqir/models/model.py         Model.__init__/to/Q, Model.init (:153-186) and the loop body ``self.svi.step()`` (:212)
  tapqir/utils/dataset.py        CosmosDataset, OffsetData (fetch, median, offset mean / logits)
  tapqir/distributions/*.py      util.py, ksmogn.py (torch branch, use_pykeops=False), affine_beta.py
Third-party packages that are absent and not installable here (pyro-ppl, pyroapi, pykeops; funsor is only imported by
``tapqir/distributions/__init__.py``, which is bypassed) are replaced by tests/golden/minipyro.py, a restatement of the
Pyro semantics this code touches.  So the model/guide/parameter code is the reference's; Pyro's inference machinery
remains restated -- DESIGN.md section 6 states the pinning status accordingly.

What is stored per case: the dataset, the reference's initial unconstrained parameters, and for every SVI iteration the
minibatch indices, the guide's base variates (recovered from the recorded samples: standard-gamma variate = sample x
rate, Beta variate = (sample - low) / scale, the Dirichlet sample itself), the loss and all 20 gradients; finally the
parameters after the last Adam update, and the reference's ``compute_probs`` (cosmos.py:609-672: z_probs, theta_probs from 50
guide particles, with the particles' variates) at those parameters.
"""

import importlib.util
import sys
import tempfile
import types
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))
REF = Path("/root/reference")


def load_reference():
    import minipyro

    minipyro.install()
    pk = types.ModuleType("pykeops")
    pk.set_verbose = lambda *a, **k: None
    pkt = types.ModuleType("pykeops.torch")
    pkt.Genred = object
    sys.modules.update({"pykeops": pk, "pykeops.torch": pkt})
    for name in ("tapqir", "tapqir.distributions", "tapqir.models", "tapqir.utils"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["tapqir"].__version__ = "reference"
    for name in ("matplotlib", "matplotlib.pyplot"):      # utils/stats.py imports them for the rastergram PNGs only
        sys.modules[name] = types.ModuleType(name)
    # SciPy 1.11 renamed ``rv_frozen.interval(alpha=...)`` to ``confidence``; cosmos.compute_params (cosmos.py:774) still
    # passes ``alpha``.  Accept both, as the SciPy versions the reference was written for did.
    from scipy.stats._distn_infrastructure import rv_frozen

    _interval = rv_frozen.interval
    rv_frozen.interval = lambda self, confidence=None, alpha=None: _interval(self, alpha if confidence is None else confidence)

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, REF / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("tapqir.exceptions", "tapqir/exceptions.py")
    load("tapqir.distributions.util", "tapqir/distributions/util.py")
    ks = load("tapqir.distributions.ksmogn", "tapqir/distributions/ksmogn.py")
    ab = load("tapqir.distributions.affine_beta", "tapqir/distributions/affine_beta.py")
    sys.modules["tapqir.distributions"].KSMOGN, sys.modules["tapqir.distributions"].AffineBeta = ks.KSMOGN, ab.AffineBeta
    ds = load("tapqir.utils.dataset", "tapqir/utils/dataset.py")
    load("tapqir.utils.stats", "tapqir/utils/stats.py")
    load("tapqir.models.model", "tapqir/models/model.py")
    cm = load("tapqir.models.cosmos", "tapqir/models/cosmos.py")
    # hmm.py: its vectorised form needs funsor (tapqir/handlers.py, tapqir/infer/); ``vectorized=False`` is the reference's
    # own sequential form of the same model (pyro.markov + TraceEnum_ELBO, hmm.py:126-131, 474-478) and needs neither
    def _logmatmulexp(x, y):      # pyro.distributions.hmm [third party]: log(exp(x) @ exp(y)), max-shifted
        xs, ys = x.detach().max(-1, keepdim=True)[0], y.detach().max(-2, keepdim=True)[0]
        return torch.matmul((x - xs).exp(), (y - ys).exp()).log() + xs + ys

    for name, attrs in {"funsor": {}, "pyro.distributions.hmm": {"_logmatmulexp": _logmatmulexp, "_sequential_index": None},
                        "tapqir.handlers": {"trace": None, "vectorized_markov": None}, "tapqir.infer": {},
                        "tapqir.infer.elbo": {"TraceMarkovEnum_ELBO": None}}.items():
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
    hm = load("tapqir.models.hmm", "tapqir/models/hmm.py")
    load("tapqir.utils.simulate", "tapqir/utils/simulate.py")
    return minipyro, ds, cm, hm


def noise_from_trace(nodes, K):
    """Base variates of the guide's reparameterised sites, in the oracle's / kernels' convention."""
    gam = lambda s: (s["value"] * (s["fn"].base_dist.rate if hasattr(s["fn"], "base_dist") else s["fn"].rate)).detach()
    beta = lambda s: ((s["value"] - s["fn"].low) / s["fn"].scale).detach()
    n = {"gain": gam(nodes["gain"]), "pi": nodes["pi"]["value"].detach(), "lamda": gam(nodes["lamda"]),
         "proximity": beta(nodes["proximity"]), "background": gam(nodes["background"])}
    n["height"] = torch.stack([gam(nodes[f"height_k{k}"]) for k in range(K)])
    for name in ("width", "x", "y"):
        n[name] = torch.stack([beta(nodes[f"{name}_k{k}"]) for k in range(K)])
    return {k: v.clone() for k, v in n.items()}


def run_case(minipyro, ds_mod, cosmos_mod, N, F, C, nb, fb, seed, offsets, perturb, masked, iters):
    from tapqir_b200.utils.simulate import simulate   # input data only (any images would do)

    kw = {}
    if offsets == "hist":
        s = torch.arange(80.0, 96.0)
        w = torch.exp(-0.5 * ((s - 90) / 3) ** 2) + 1e-3
        kw = dict(offset_samples=s, offset_weights=w / w.sum())
    sim = simulate(N, F, C=C, seed=seed, **kw)
    mask = sim.mask.clone()
    if masked is not None:
        mask[masked] = False
    model = cosmos_mod.cosmos(device="cpu", dtype="double", use_pykeops=False)     # sets the default dtype to double
    data = ds_mod.CosmosDataset(sim.images.double(), sim.xy.double(), sim.is_ontarget, mask, None,
                                sim.offset.samples.double(), sim.offset.weights.double())
    model.data = data
    model.run_path = Path(tempfile.mkdtemp())
    torch.manual_seed(seed)
    model.init(lr=0.005, nbatch_size=nb, fbatch_size=fb)
    store = minipyro.get_param_store().unconstrained()
    init_unconstrained = {k: v.detach().clone() for k, v in store.items()}
    if perturb:
        g = torch.Generator().manual_seed(seed + 100)
        with torch.no_grad():
            for v in store.values():
                v.add_(0.3 * torch.randn(v.shape, generator=g, dtype=v.dtype))
    start = {k: v.detach().clone() for k, v in store.items()}
    steps = []
    for _ in range(iters):
        loss = model.svi.step()                        # model.py:212
        nodes = model.elbo.last_guide_trace
        ndx = nodes["aois"]["value"] if "aois" in nodes else torch.arange(N)
        fdx = nodes["frames"]["value"] if "frames" in nodes else torch.arange(F)
        model_nodes = model.elbo.last_model_trace
        shapes = {k: tuple(model_nodes[k]["value"].shape) for k in ("z", "theta", "m_k0", "m_k1")}
        steps.append(dict(ndx=ndx.clone(), fdx=fdx.clone(), noise=noise_from_trace(nodes, model.K), loss=loss,
                          grads={k: v.clone() for k, v in model.svi.last_grads.items()}, enum_shapes=shapes))
    final = {k: v.detach().clone() for k, v in store.items()}
    # posterior of z / theta at the final parameters: the reference's own compute_probs (cosmos.py:609-672), with the
    # batch sizes opened up so that ONE set of 50 guide particles covers all on-target AOIs and frames
    model.nbatch_size, model.fbatch_size = N, F
    minipyro.LAST_TRACES.clear()
    z_probs, theta_probs = model.compute_probs
    gnodes = minipyro.LAST_TRACES[0].nodes            # the first trace compute_probs takes is the guide's
    K = model.K
    gam = lambda s: (s["value"] * (s["fn"].base_dist.rate if hasattr(s["fn"], "base_dist") else s["fn"].rate)).detach()
    beta = lambda s: ((s["value"] - s["fn"].low) / s["fn"].scale).detach()
    lam, prox, pi = gam(gnodes["lamda"]), beta(gnodes["proximity"]), gnodes["pi"]["value"].detach()
    xs = torch.stack([beta(gnodes[f"x_k{k}"]) for k in range(K)], 1)       # (50, K, n_on, F, C)
    ys = torch.stack([beta(gnodes[f"y_k{k}"]) for k in range(K)], 1)
    stack = lambda name, f: torch.stack([f(gnodes[f"{name}_k{k}"]) for k in range(K)], 1)
    particles = dict(pi=pi[:, 0, 0, 0].clone(), lamda=lam[:, 0, 0, 0].clone(), proximity=prox[:, 0, 0, 0].clone(),
                     x=xs.clone(), y=ys.clone(), gain=gam(gnodes["gain"])[:, 0, 0, 0].clone(),
                     background=gam(gnodes["background"]).clone(), height=stack("height", gam),
                     width=stack("width", beta))          # leading axis = the 50 particles
    probs = dict(z_probs=z_probs.clone(), theta_probs=theta_probs.clone(), particles=particles,
                 n_on=int(sim.is_ontarget.sum()))
    assert sim.images.min() >= 0 and sim.images.max() < 65536 and (sim.images == sim.images.floor()).all()
    return dict(config=dict(N=N, F=F, C=C, nb=nb, fb=fb, seed=seed, offsets=offsets, lr=0.005),
                images=sim.images.to(torch.int32), xy=sim.xy.double(), is_ontarget=sim.is_ontarget, mask=mask,
                offset_samples=sim.offset.samples.double(), offset_weights=sim.offset.weights.double(),
                init_unconstrained=init_unconstrained, start=start, steps=steps, final=final, probs=probs)


def run_hmm_case(minipyro, ds_mod, hmm_mod, N, F, C, nb, seed, perturb, iters):
    """hmm.model / hmm.guide / hmm.init_parameters (models/hmm.py:82-467) in the reference's sequential form
    (``vectorized=False``): the guide enumerates the chain z_f and m_k|z_f, the model theta_f; all frames every step."""
    from tapqir_b200.utils.simulate import simulate

    sim = simulate(N, F, C=C, seed=seed, params={"kon": 0.2, "koff": 0.2})
    model = hmm_mod.hmm(device="cpu", dtype="double", use_pykeops=False, vectorized=False)
    model.data = ds_mod.CosmosDataset(sim.images.double(), sim.xy.double(), sim.is_ontarget, sim.mask.clone(), None,
                                      sim.offset.samples.double(), sim.offset.weights.double())
    model.run_path = Path(tempfile.mkdtemp())
    torch.manual_seed(seed)
    model.init(lr=0.005, nbatch_size=nb, fbatch_size=F)
    store = minipyro.get_param_store().unconstrained()
    init_unconstrained = {k: v.detach().clone() for k, v in store.items()}
    if perturb:
        g = torch.Generator().manual_seed(seed + 100)
        with torch.no_grad():
            for v in store.values():
                v.add_(0.3 * torch.randn(v.shape, generator=g, dtype=v.dtype))
    start = {k: v.detach().clone() for k, v in store.items()}
    gam = lambda s: (s["value"] * (s["fn"].base_dist.rate if hasattr(s["fn"], "base_dist") else s["fn"].rate)).detach()
    beta = lambda s: ((s["value"] - s["fn"].low) / s["fn"].scale).detach()
    K = model.K
    steps = []
    for _ in range(iters):
        loss = model.svi.step()
        nodes = model.elbo.last_guide_trace
        ndx = nodes["aois"]["value"] if "aois" in nodes else torch.arange(N)
        frames = lambda name, f: torch.cat([f(nodes[f"{name}_f{i}"]) for i in range(F)], -2)       # (nb, F, C)
        noise = {"gain": gam(nodes["gain"]), "init": nodes["init"]["value"].detach(), "trans": nodes["trans"]["value"].detach(),
                 "lamda": gam(nodes["lamda"]), "proximity": beta(nodes["proximity"]), "background": frames("background", gam),
                 "height": torch.stack([frames(f"height_k{k}", gam) for k in range(K)]),
                 "width": torch.stack([frames(f"width_k{k}", beta) for k in range(K)]),
                 "x": torch.stack([frames(f"x_k{k}", beta) for k in range(K)]),
                 "y": torch.stack([frames(f"y_k{k}", beta) for k in range(K)])}
        steps.append(dict(ndx=ndx.clone(), noise={k: v.clone() for k, v in noise.items()}, loss=loss,
                          grads={k: v.clone() for k, v in model.svi.last_grads.items()}))
    final = {k: v.detach().clone() for k, v in store.items()}
    z_probs = model.z_probs.clone()       # hmm.py:627-633 via its own _sequential_logmatmulexp (:480-533), all AOIs
    # credible intervals of the two sites cosmos does not have: the branches of cosmos.compute_params for "init" / "trans"
    # (cosmos.py:722-725 + :773-778); the method as a whole needs hmm.theta_probs, which is written in funsor terms
    from tapqir.utils.stats import torch_to_scipy_dist

    ci = {}
    for name in ("init", "trans"):
        fn = minipyro.Dirichlet(minipyro.param(f"{name}_mean") * minipyro.param(f"{name}_size"))
        LL, UL = torch_to_scipy_dist(fn).interval(alpha=0.95)
        ci[name] = dict(LL=torch.as_tensor(LL).clone(), UL=torch.as_tensor(UL).clone(), Mean=fn.mean.detach().clone())
    return dict(config=dict(N=N, F=F, C=C, nb=nb, seed=seed, lr=0.005), z_probs=z_probs, ci=ci, CI=0.95, images=sim.images.to(torch.int32), xy=sim.xy.double(),
                is_ontarget=sim.is_ontarget, mask=sim.mask.clone(), offset_samples=sim.offset.samples.double(),
                offset_weights=sim.offset.weights.double(), init_unconstrained=init_unconstrained, start=start, steps=steps,
                final=final)


def run_data_case(ds_mod):
    """``data.tpqr`` as written by the reference's own ``utils/dataset.py::save`` (:195-213), plus what its CosmosDataset /
    OffsetData report for it -- the state-file contract of SURVEY 8(b)."""
    import numpy as np

    from tapqir_b200.utils.simulate import simulate

    N, F, C = 4, 5, 2
    s = torch.arange(85.0, 96.0)
    w = torch.exp(-0.5 * ((s - 90) / 2.5) ** 2)
    sim = simulate(N, F, C=C, seed=11, offset_samples=s, offset_weights=w / w.sum())
    labels = np.zeros((N // 2, F, C), dtype=[("aoi", int), ("frame", int), ("z", int)])      # simulate.py:110-113
    labels["aoi"] = np.arange(N // 2).reshape(-1, 1, 1)
    labels["frame"] = np.arange(F).reshape(-1, 1)
    labels["z"] = (np.arange(N // 2 * F * C).reshape(N // 2, F, C) % 3 == 0)
    mask = torch.tensor([True, True, False, True])
    ref = ds_mod.CosmosDataset(sim.images.double(), sim.xy.double(), sim.is_ontarget, mask, labels, sim.offset.samples.double(),
                               sim.offset.weights.double(), name="golden", time1=torch.arange(F).double(), ttb=None,
                               channels=("green", "red"))
    out_dir = HERE / "ref_data"
    out_dir.mkdir(exist_ok=True)
    ds_mod.save(ref, out_dir)
    facts = dict(N=int(ref.N), Nc=int(ref.Nc), Nt=int(ref.Nt), F=ref.F, C=ref.C, P=ref.P, median=ref.median.clone(),
                 x=ref.x.clone(), y=ref.y.clone(), offset_min=ref.offset.min, offset_max=ref.offset.max,
                 offset_mean=ref.offset.mean, offset_var=ref.offset.var, offset_logits=ref.offset.logits.clone(),
                 channels=ref.channels, name=ref.name)
    ndx, fdx, cdx = torch.tensor([3, 0])[:, None, None], torch.tensor([4, 1, 2])[:, None], torch.arange(C)
    obs, target, ont = ref.fetch(ndx, fdx, cdx)                                                  # dataset.py:140-151
    facts.update(fetch_ndx=ndx, fetch_fdx=fdx, fetch_obs=obs.clone(), fetch_target=target.clone(), fetch_ontarget=ont.clone())
    torch.save(facts, out_dir / "facts.pt")
    print("wrote", out_dir)


def run_c1_fit(minipyro, ds_mod, cosmos_mod, iters=100, seed=0):
    """BASELINE.json configs[0]: cosmos, simulated N=5 AOIs x F=100 frames, P=14, K=2, 100 SVI iterations on the CPU, full
    batch (the reference's test-suite scale, test/test_tapqir.py:22-50,91-93, which asserts only the exit code).  Only the
    losses and the parameters after the last update are stored: with a full batch nothing but the guide's variates consumes
    the random stream, in the guide's site order, so ``torch.manual_seed(seed)`` + the same draws (oracle.draw_noise)
    reproduce the run; the first iteration's variates are stored to check that alignment."""
    from tapqir_b200.utils.simulate import simulate

    N, F = 5, 100
    sim = simulate(N, F, C=1, seed=seed)
    model = cosmos_mod.cosmos(device="cpu", dtype="double", use_pykeops=False)
    model.data = ds_mod.CosmosDataset(sim.images.double(), sim.xy.double(), sim.is_ontarget, sim.mask.clone(), None,
                                      sim.offset.samples.double(), sim.offset.weights.double())
    model.run_path = Path(tempfile.mkdtemp())
    assert sim.images.max() < 32768
    model.init(lr=0.005, nbatch_size=N, fbatch_size=F)
    torch.manual_seed(seed + 1000)
    losses, first = [], None
    for it in range(iters):
        losses.append(model.svi.step())
        if it == 0:
            first = noise_from_trace(model.elbo.last_guide_trace, model.K)
    store = minipyro.get_param_store().unconstrained()
    stats = model.compute_params(0.95)                  # cosmos.py:711-784 (scipy intervals through stats.torch_to_scipy_dist)
    ci = {k: {s: torch.as_tensor(v).clone() for s, v in stats[k].items()} for k in model.ci_params}
    return dict(config=dict(N=N, F=F, C=1, nb=N, fb=F, seed=seed, rng_seed=seed + 1000, lr=0.005, iters=iters), ci=ci, CI=0.95,
                images=sim.images.to(torch.int16), xy=sim.xy.double(), is_ontarget=sim.is_ontarget, mask=sim.mask.clone(),
                offset_samples=sim.offset.samples.double(), offset_weights=sim.offset.weights.double(),
                losses=torch.tensor(losses, dtype=torch.float64), first_noise=first,
                final={k: v.detach().clone() for k, v in store.items()})


def run_simulate_case(cosmos_mod, hmm_mod):
    """The reference's own ``utils/simulate.py::simulate`` (:12-138: Predictive over the unconditioned cosmos model, pixels
    from ``KSMOGN.rsample``) with the constants of its test-suite (test/test_tapqir.py:22-50): what is stored are summary
    statistics of the simulated movie, which the simulator of this repository has to reproduce in distribution."""
    from tapqir.utils.simulate import simulate

    prm = {"pi": 0.15, "width": 1.4, "gain": 7.0, "lamda": 0.15, "proximity": 0.2, "offset": 90.0, "height": 3000, "background": 150}
    N, F, C, P = 40, 100, 1, 14
    model = cosmos_mod.cosmos(device="cpu", dtype="double", use_pykeops=False)
    data = simulate(model, N, F, C, P, seed=4, params=prm)
    img = data.images.double()
    corners = torch.stack([img[..., 0, 0], img[..., 0, P - 1], img[..., P - 1, 0], img[..., P - 1, P - 1]], -1)
    patch_sum = img.sum((-1, -2)) - (prm["background"] + prm["offset"] - 0.5) * P * P      # photons above background
    z = torch.as_tensor(data.labels["z"])
    stats = dict(params=prm, N=N, F=F, C=C, P=P, shape=tuple(img.shape), is_ontarget=data.is_ontarget.clone(),
                 xy_unique=torch.unique(data.xy), offset_samples=data.offset.samples.clone(), offset_weights=data.offset.weights.clone(),
                 images_dtype=str(data.images.dtype), integral=bool((img == img.floor()).all()),
                 z_fraction=z.double().mean().item(), labels_shape=tuple(data.labels.shape), labels_fields=data.labels.dtype.names,
                 corner_mean=corners.mean().item(), corner_var=corners.var().item(), pixel_mean=img.mean().item(),
                 patch_sum_on=patch_sum[: N // 2].mean().item(), patch_sum_off=patch_sum[N // 2:].mean().item(),
                 patch_sum_q=torch.quantile(patch_sum.flatten(), torch.tensor([0.5, 0.9, 0.99], dtype=torch.float64)),
                 min=img.min().item())
    # kinetic recipe (test/test_tapqir.py:30-33): the reference's hmm model in its sequential form, kon = koff = 0.2
    hprm = {k: v for k, v in prm.items() if k != "pi"}
    hprm.update(kon=0.2, koff=0.2)
    hN, hF = 100, 100
    hmodel = hmm_mod.hmm(device="cpu", dtype="double", use_pykeops=False, vectorized=False)
    hdata = simulate(hmodel, hN, hF, C, P, seed=5, params=hprm)
    hz = torch.as_tensor(hdata.labels["z"])[..., 0]                                   # (N/2, F)
    prev, cur = hz[:, :-1], hz[:, 1:]
    himg = hdata.images.double()
    hsum = himg.sum((-1, -2)) - (prm["background"] + prm["offset"] - 0.5) * P * P
    stats["hmm"] = dict(params=hprm, N=hN, F=hF, z_fraction=hz.double().mean().item(), z0_fraction=hz[:, 0].double().mean().item(),
                        p01=((prev == 0) & (cur == 1)).sum().item() / max((prev == 0).sum().item(), 1),
                        p10=((prev == 1) & (cur == 0)).sum().item() / max((prev == 1).sum().item(), 1),
                        patch_sum_on=hsum[: hN // 2].mean().item(), patch_sum_off=hsum[hN // 2:].mean().item(),
                        shape=tuple(himg.shape), labels_shape=tuple(hdata.labels.shape))
    print("simulate hmm:", {k: v for k, v in stats["hmm"].items() if k not in ("params",)})
    torch.save(stats, HERE / "ref_simulate_stats.pt")
    print("simulate:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in stats.items() if k in ("z_fraction", "corner_mean", "corner_var", "pixel_mean", "patch_sum_on", "patch_sum_off")})


def run_glimpse_header_case():
    """``imscroll/glimpse_reader.py::GlimpseDataset.__init__`` (:55-159) run verbatim on a small synthetic glimpse folder
    (tests/golden/ref_glimpse_folder/, written here): header, AOI tables in the three accepted layouts, the cumulative
    drift relative to the frame the AOIs were picked in, a frame range, spot-picker labels.  (``__getitem__`` / the frame loop
    of ``read_glimpse`` add 2**15 to an int16 array, which numpy 2 refuses: they stay restated in oracle/glimpse_oracle.py.)"""
    import numpy as np
    from scipy.io import savemat

    patches = types.ModuleType("matplotlib.patches")
    patches.Rectangle = object
    sys.modules["matplotlib.patches"] = patches
    spec = importlib.util.spec_from_file_location("tapqir.imscroll.glimpse_reader", REF / "tapqir/imscroll/glimpse_reader.py")
    pkg = types.ModuleType("tapqir.imscroll")
    pkg.__path__ = []
    sys.modules["tapqir.imscroll"] = pkg
    gr = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = gr
    spec.loader.exec_module(gr)

    folder = HERE / "ref_glimpse_folder"
    (folder / "glimpse").mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(3)
    F, H, W = 12, 40, 48
    vid = dict(height=float(H), width=float(W), filenumber=np.repeat(np.arange(3), 4).astype(np.int32),
               offset=(np.arange(F) % 4 * H * W * 2).astype(np.int64), ttb=np.arange(F, dtype=float) * 50.0 + 7.0, time1=1234.5)
    savemat(folder / "glimpse" / "header.mat", {"vid": vid})
    dl = np.zeros((F, 4))
    dl[:, 0] = np.arange(1, F + 1)
    dl[:, 1:3] = rng.normal(0, 0.3, size=(F, 2))
    savemat(folder / "driftlist.mat", {"driftlist": dl})

    def table(n, frame):      # frame, ave, y, x, pixnum, aoi (MATLAB indexing)
        xy = rng.uniform(10, 30, size=(n, 2))
        return np.stack([np.full(n, float(frame)), np.full(n, 10.0), xy[:, 1], xy[:, 0], np.full(n, 5.0), np.arange(1, n + 1)], 1)

    on, off = table(4, 5), table(3, 5)
    savemat(folder / "ontarget_aoiinfo2.mat", {"aoiinfo2": on})                       # layout 1: plain aoiinfo2
    savemat(folder / "offtarget_aoifits.mat", {"aoifits": {"aoiinfo2": off}})         # layout 2: inside aoifits
    np.savetxt(folder / "ontarget.dat", on)                                           # layout 3: text
    intervals = np.array([[-2.0, 1, 3, 2, 0, 0, 1], [1.0, 4, 9, 6, 0, 0, 1], [2.0, 10, 12, 3, 0, 0, 1],
                          [-3.0, 1, 6, 6, 0, 0, 3], [0.0, 7, 12, 6, 0, 0, 3], [3.0, 2, 11, 10, 0, 0, 4]])
    savemat(folder / "intervals.mat", {"Intervals": {"CumulativeIntervalArray": intervals}})

    def kwargs(on_file, frame_range, labels):
        return {"name": "green", "glimpse-folder": str(folder / "glimpse"), "driftlist": str(folder / "driftlist.mat"),
                "ontarget-aoiinfo": str(folder / on_file), "offtarget-aoiinfo": str(folder / "offtarget_aoifits.mat"),
                "use-offtarget": True, "frame-range": frame_range, "frame-start": 3, "frame-end": 10, "labels": labels,
                "ontarget-labels": str(folder / "intervals.mat") if labels else None, "offtarget-labels": None,
                "offset-x": 2, "offset-y": 3}

    facts = {}
    for name, kw in {"mat_full": kwargs("ontarget_aoiinfo2.mat", False, True), "text_range": kwargs("ontarget.dat", True, True),
                     "mat_nolabels": kwargs("ontarget_aoiinfo2.mat", True, False)}.items():
        g = gr.GlimpseDataset(**kw)
        facts[name] = dict(
            kwargs={k: (Path(v).name if isinstance(v, str) and "/" in v else v) for k, v in kw.items()},
            height=g.height, width=g.width, dtypes=list(g.dtypes), offset=(g.offset_x, g.offset_y), name=g.name,
            time1=float(g.header["time1"]), filenumber=np.asarray(g.header["filenumber"]).copy(),
            aoiinfo={d: dict(index=g.aoiinfo[d].index.values.copy(), values=g.aoiinfo[d][["frame", "ave", "y", "x", "pixnum"]].values.copy())
                     for d in g.dtypes},
            cumdrift=dict(index=g.cumdrift.index.values.copy(), values=g.cumdrift[["dy", "dx", "ttb"]].values.copy()),
            labels={d: (None if g.labels[d] is None else g.labels[d].copy()) for d in g.dtypes})
    # ---- the whole of read_glimpse (:304-472) on the same folder: frame decoding, AOI cropping, offset histogram ----------
    # Two third-party API changes stand between the reference and this environment, both bridged, not worked around in the
    # reference's code: numpy 2 no longer promotes ``int16_array + 2**15`` (glimpse_reader.py:186) to a wider integer --
    # numpy 1 gave int32 --, and matplotlib (the AOI overview PNGs of GlimpseDataset.plot) is absent.
    class _PromotingInt16(np.ndarray):
        def __add__(self, other):
            return np.asarray(self).astype(np.int32) + other if isinstance(other, int) else np.asarray(self) + other

    class _Numpy1(types.ModuleType):
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def fromfile(*a, **kw):
            return np.fromfile(*a, **kw).view(_PromotingInt16)

    class _Canvas:                              # every pyplot call is accepted and ignored
        def __getattr__(self, name):
            return self

        def __call__(self, *a, **kw):
            return self

    gr.np, gr.plt = _Numpy1("numpy"), _Canvas()
    gr.GlimpseDataset.plot = lambda self, *a, **kw: None
    sys.modules["tapqir.utils.dataset"].quantile = lambda x, q: torch.quantile(x, q)      # pyro.ops.stats.quantile (vmin / vmax)
    decoded = rng.integers(250, 1200, size=(F, H, W))
    decoded[:, 3:13, 2:12] = rng.integers(85, 100, size=(F, 10, 10))                  # the dark corner the offsets are taken from
    for number in range(3):
        with open(folder / "glimpse" / f"{number}.glimpse", "wb") as fid:
            fid.write((decoded[4 * number: 4 * number + 4] - 2 ** 15).astype(">i2").tobytes())
    out_dir = folder / "out"
    out_dir.mkdir(exist_ok=True)
    kw = kwargs("ontarget_aoiinfo2.mat", True, True)
    channel = {k: kw.pop(k) for k in ("name", "glimpse-folder", "driftlist", "ontarget-aoiinfo", "offtarget-aoiinfo",
                                      "ontarget-labels", "offtarget-labels")}
    kw.update({"P": 14, "num-channels": 1, "dataset": "golden-movie", "channels": [channel], "offset-P": 10, "bin-size": 3})
    torch.set_default_dtype(torch.float32)          # `tapqir glimpse` runs under the FloatTensor default (main.py:205)
    try:
        gr.read_glimpse(out_dir, lambda it: it, **kw)
    finally:
        torch.set_default_dtype(torch.float64)
    facts["read_glimpse"] = dict(decoded=torch.as_tensor(decoded, dtype=torch.int32), P=14, offset_P=10, bin_size=3,
                                 frame_start=3, frame_end=10)
    torch.save(facts, folder / "facts.pt")
    print("wrote", folder)


def main():
    minipyro, ds_mod, cosmos_mod, hmm_mod = load_reference()
    run_data_case(ds_mod)
    run_glimpse_header_case()
    run_simulate_case(cosmos_mod, hmm_mod)
    c1 = run_c1_fit(minipyro, ds_mod, cosmos_mod)
    torch.save(c1, HERE / "ref_c1_fit.pt")
    print("c1 fit: loss", c1["losses"][0].item(), "->", c1["losses"][-1].item())
    cases = {
        "c1_initial_point": dict(N=4, F=6, C=1, nb=3, fb=4, seed=0, offsets="sim", perturb=False, masked=None, iters=5),
        "c1_perturbed_masked": dict(N=5, F=6, C=1, nb=4, fb=4, seed=1, offsets="sim", perturb=True, masked=2, iters=5),
        "c2_hist_offsets": dict(N=4, F=5, C=2, nb=3, fb=3, seed=2, offsets="hist", perturb=True, masked=None, iters=4),
        "c1_full_batch": dict(N=3, F=4, C=1, nb=3, fb=4, seed=3, offsets="sim", perturb=True, masked=None, iters=3),
    }
    out = {name: run_case(minipyro, ds_mod, cosmos_mod, **kw) for name, kw in cases.items()}
    hmm_cases = {
        "hmm_c1": dict(N=3, F=4, C=1, nb=2, seed=5, perturb=True, iters=3),
        "hmm_c2_initial_point": dict(N=2, F=3, C=2, nb=2, seed=6, perturb=False, iters=3),
        "hmm_zprobs_only": dict(N=3, F=23, C=2, nb=3, seed=7, perturb=True, iters=0),       # odd chain length for the scan
    }
    hmm_out = {name: run_hmm_case(minipyro, ds_mod, hmm_mod, **kw) for name, kw in hmm_cases.items()}
    torch.save(hmm_out, HERE / "ref_step_hmm.pt")
    for name, c in hmm_out.items():
        print(name, "losses", [round(s["loss"], 4) for s in c["steps"]])
    torch.save(out, HERE / "ref_step.pt")
    for name, c in out.items():
        print(name, "losses", [round(s["loss"], 4) for s in c["steps"]], c["steps"][0]["enum_shapes"])
    print("wrote", HERE / "ref_step.pt", (HERE / "ref_step.pt").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
