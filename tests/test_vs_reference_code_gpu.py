"""
GPU: the CUDA path against tests/golden/ref_step.pt -- SVI iterations and posteriors produced by the reference's OWN
``models/cosmos.py`` (init_parameters, guide, model, compute_probs) and ``models/model.py`` (Model.init, svi.step), run
verbatim in the build container by tests/golden/make_golden_step.py with the absent pyro / pyroapi packages replaced by
tests/golden/minipyro.py.  Replay mode: recorded minibatch indices and guide variates.
"""

import pytest
import torch

from oracle import cosmos_oracle as O
from tapqir_b200.models import layout as L
from tests.step_helpers import check_global_grads, compare_grads, flat_inputs, golden_step_case, masked_loss_constant
from tests.test_step_gpu import make_engine, replay_args

pytestmark = pytest.mark.gpu
CASES = ["c1_initial_point", "c1_perturbed_masked", "c2_hist_offsets", "c1_full_batch"]


@pytest.mark.parametrize("name", CASES)
def test_svi_iterations_match_reference_model_code(name):
    """tests/golden/ref_step.pt: SVI iterations produced by the reference's own cosmos.py (init_parameters, guide,
    model) and model.py (Model.init, svi.step) -- tests/golden/make_golden_step.py.  The fp64 kernels replay the recorded
    minibatches and base variates from the recorded starting parameters: every loss (1e-11; plus the constant Pyro
    adds for masked AOIs, tests/test_oracle.py), the gradients of the first iteration (1e-8 of each tensor's largest
    entry) and the parameters after the last Adam update (1e-8; the host build of the same arithmetic is at 1e-11)."""
    ds, data, case = golden_step_case(name)
    cfg = case["config"]
    start = {k: v.reshape(O.init_constrained(data)[k].shape).clone() for k, v in case["start"].items()}
    eng = make_engine(ds, data, start, cfg["nb"], cfg["fb"], torch.float64, lr=cfg["lr"])
    first = case["steps"][0]
    loss = eng.step(update=False, **replay_args(eng, data, start, first["ndx"], first["fdx"], first["noise"], torch.float64)).item()
    ref_loss = first["loss"] + masked_loss_constant(case, first)
    assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
    ref_grads = {k: g.reshape(start[k].shape) for k, g in first["grads"].items()}
    bad = compare_grads(eng.named_grads(), ref_grads, 1e-8)
    assert not bad, bad
    for it, step in enumerate(case["steps"]):
        loss = eng.step(**replay_args(eng, data, start, step["ndx"], step["fdx"], step["noise"], torch.float64)).item()
        ref_loss = step["loss"] + masked_loss_constant(case, step)
        assert abs(loss - ref_loss) <= 1e-9 * abs(ref_loss), (it, loss, ref_loss)
    assert eng.iteration == len(case["steps"])
    ours = eng.named_unconstrained()
    for k, v in case["final"].items():
        err = (ours[k].double().cpu().reshape(-1) - v.reshape(-1)).abs().max().item()
        assert err <= 1e-8 * max(1.0, v.abs().max().item()), (k, err)


@pytest.mark.parametrize("name", ["c1_perturbed_masked", "c2_hist_offsets"])
def test_compute_probs_matches_reference_model_code(name):
    """z_probs / theta_probs of the reference's own compute_probs (cosmos.py:609-672, 50 guide particles, run by
    tests/golden/make_golden_step.py) from the fp64 kernels fed the same particles' variates; 1e-9 (the reference's
    unmasked x, y terms differ from the masked form at the eps level, tests/test_oracle.py)."""
    ds, data, case = golden_step_case(name)
    cfg, probs = case["config"], case["probs"]
    final = {k: v.reshape(O.init_constrained(data)[k].shape).clone() for k, v in case["final"].items()}
    n_on, part = probs["n_on"], probs["particles"]
    noises = [{k: v[i] for k, v in part.items()} for i in range(part["pi"].shape[0])]
    eng = make_engine(ds, data, final, cfg["nb"], cfg["fb"], torch.float64)
    ndx, fdx = torch.arange(n_on), torch.arange(data.F)
    flat = [flat_inputs(data, final, n, torch.float64) for n in noises]
    z, th = eng.compute_probs(particles=len(noises), ndx=ndx.to(torch.int32).cuda(), fdx=fdx.to(torch.int32).cuda(),
                              local_noise=[f[4].cuda() for f in flat], global_noise=[f[5].cuda() for f in flat])
    assert (z.double().cpu() - probs["z_probs"][:n_on]).abs().max().item() < 1e-9
    assert (th.double().cpu() - probs["theta_probs"][:, :n_on]).abs().max().item() < 1e-9


@pytest.mark.parametrize("name", CASES)
def test_fp32_production_kernels_within_north_star_of_reference_model_code(name):
    """The fp32 production kernels at every recorded iteration (parameters of the reference's trajectory, replayed by
    the oracle's Adam which tests/test_oracle.py pins to the same file), against the reference's fp64 numbers with
    NO rounding on the reference side: the north-star tolerance -- loss 1e-6, gradients 1e-5 of each tensor's largest
    entry (global ones: of their OWN magnitude, or of their forward-error scale where the entry itself is a cancelling
    sum, step_helpers.check_global_grads -- the global parameters are float64 on the device).  Measured on a
    B200 (profiles/r1s5_parity_vs_reference_code.log): loss <= 7.6e-7, local gradients <= 8.9e-6, global <= 1.8e-6;
    replay mode is deterministic (fixed-order reductions), so the margins do not move between runs."""
    ds, data, case = golden_step_case(name)
    cfg = case["config"]
    svi = O.OracleSVI(data, lr=cfg["lr"], nbatch_size=cfg["nb"], fbatch_size=cfg["fb"])
    with torch.no_grad():
        for k, v in svi.params.items():
            v.copy_(case["start"][k].reshape(v.shape))
    eng = None
    for it, step in enumerate(case["steps"]):
        params = {k: v.detach().clone() for k, v in svi.params.items()}
        if eng is None:
            eng = make_engine(ds, data, params, cfg["nb"], cfg["fb"], torch.float32)
        else:
            eng.load_unconstrained(params)
        loss = eng.step(update=False, **replay_args(eng, data, params, step["ndx"], step["fdx"], step["noise"], torch.float32)).item()
        ref_loss = step["loss"] + masked_loss_constant(case, step)
        assert abs(loss - ref_loss) <= 1e-6 * abs(ref_loss), (it, loss, ref_loss)
        ref_grads = {k: g.reshape(params[k].shape) for k, g in step["grads"].items()}
        bad = compare_grads(eng.named_grads(), ref_grads, 1e-5, names=L.LOCAL_NAMES)
        print(f"[{name} it {it}] loss rel {abs(loss - ref_loss) / abs(ref_loss):.2e}; worst local gradient rel "
              f"{max(compare_grads(eng.named_grads(), ref_grads, 0.0, names=L.LOCAL_NAMES).values()):.2e}; worst global (self) "
              f"{max(compare_grads(eng.named_grads(), ref_grads, 0.0, names=L.GLOBAL_NAMES).values()):.2e}")
        # global gradients: 1e-5 of their OWN magnitude where nothing cancels, of their forward-error scale elsewhere
        bad.update(check_global_grads(eng.named_grads(), ref_grads, params, data, step["ndx"], step["fdx"], step["noise"]))
        assert not bad, (it, bad)
        svi.step(step["ndx"], step["fdx"], step["noise"])


# ---- hmm: tests/golden/ref_step_hmm.pt (models/hmm.py run verbatim in its sequential form) -----------------------------
@pytest.mark.parametrize("name", ["hmm_c1", "hmm_c2_initial_point"])
@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float64, 1e-11, 1e-8), (torch.float32, 1e-6, 1e-5)])
def test_hmm_iterations_match_reference_model_code(name, dtype, ltol, gtol):
    """Every recorded iteration of the reference's hmm guide()/model() (brute-force expectation over the enumerated
    chain) from the hmm kernels (chunked scans): loss and all gradients at the reference's parameters of that
    iteration; fp64 additionally follows the Adam trajectory to the recorded final parameters."""
    from oracle import hmm_oracle as H
    from tests.test_hmm_cpu import adam_update, hmm_global_grads, hmm_golden_case
    from tests.test_hmm_gpu import make_engine as make_hmm_engine

    ds, data, case = hmm_golden_case(name)
    shapes = {k: v.shape for k, v in H.init_constrained(data).items()}
    p = {k: v.reshape(shapes[k]).clone() for k, v in case["start"].items()}
    m, v2 = {k: torch.zeros_like(x) for k, x in p.items()}, {k: torch.zeros_like(x) for k, x in p.items()}
    nb = case["config"]["nb"]
    eng = make_hmm_engine(ds, data, p, nb, dtype)
    traj = make_hmm_engine(ds, data, p, nb, dtype) if dtype == torch.float64 else None
    for t, step in enumerate(case["steps"], 1):
        eng.load_unconstrained(p)
        lnoise = L.pack_local_noise(step["noise"], dtype, "cuda")
        gnoise = eng.gl.pack_noise(step["noise"]).cuda()
        ndx = step["ndx"].to(torch.int32).cuda()
        loss = eng.step(update=False, ndx=ndx, local_noise=lnoise, global_noise=gnoise).item()
        assert abs(loss - step["loss"]) <= ltol * abs(step["loss"]), (t, loss, step["loss"])
        ref_grads = {k: g.reshape(shapes[k]) for k, g in step["grads"].items()}
        if dtype == torch.float64:
            bad = compare_grads(eng.named_grads(), ref_grads, gtol)
        else:
            bad = compare_grads(eng.named_grads(), ref_grads, gtol, names=[k for k in ref_grads if k not in H.GLOBAL_PARAMS])
            bad.update(hmm_global_grads(eng.named_grads(), ref_grads, gtol))
        assert not bad, (t, bad)
        if traj is not None:
            traj.step(ndx=ndx, local_noise=lnoise, global_noise=gnoise)
        _, og = H.loss_and_grads(p, data, step["ndx"], step["noise"])
        adam_update(p, og, m, v2, t)
    if traj is not None:
        ours = traj.named_unconstrained()
        for k, v in case["final"].items():
            err = (ours[k].double().cpu().reshape(-1) - v.reshape(-1)).abs().max().item()
            assert err <= 1e-8 * max(1.0, v.abs().max().item()), (k, err)


@pytest.mark.parametrize("name", ["hmm_c1", "hmm_zprobs_only"])
def test_hmm_z_probs_match_reference_model_code(name):
    """hmm.z_probs (hmm.py:627-633, the reference's own ``_sequential_logmatmulexp`` scan) at the recorded parameters."""
    from oracle import hmm_oracle as H
    from tests.test_hmm_cpu import hmm_golden_case
    from tests.test_hmm_gpu import make_engine as make_hmm_engine

    ds, data, case = hmm_golden_case(name)
    shapes = {k: v.shape for k, v in H.init_constrained(data).items()}
    final = {k: v.reshape(shapes[k]).clone() for k, v in case["final"].items()}
    eng = make_hmm_engine(ds, data, final, case["config"]["nb"], torch.float64)
    zp = eng.z_probs().cpu().double()
    assert zp.shape == case["z_probs"].shape and (zp - case["z_probs"]).abs().max().item() <= 1e-12


def test_c1_hundred_iterations_match_the_reference_run():
    """BASELINE configs[0] -- N=5 AOIs x F=100 frames, full batch, 100 SVI iterations -- as run by the reference's own
    cosmos.py / model.py (tests/golden/ref_c1_fit.pt).  The guide's variates are re-drawn from the recorded seed at the
    reference trajectory's parameters (the oracle follows it to 1e-14, tests/test_oracle.py); the fp64 kernels with the
    dense Adam kernel take the same variates and follow their own trajectory: every loss 1e-9, final parameters 1e-7
    (the host build of the same arithmetic: 1e-11 / 1e-8, tests/test_hostcheck_step.py)."""
    from tests.step_helpers import golden_c1_fit

    ds, data, case = golden_c1_fit()
    cfg = case["config"]
    svi = O.OracleSVI(data, lr=cfg["lr"], nbatch_size=cfg["nb"], fbatch_size=cfg["fb"])
    start = {k: v.detach().clone() for k, v in svi.params.items()}
    eng = make_engine(ds, data, start, cfg["nb"], cfg["fb"], torch.float64, lr=cfg["lr"])
    ndx, fdx = torch.arange(cfg["N"]), torch.arange(cfg["F"])
    state = torch.get_rng_state()
    try:
        torch.manual_seed(cfg["rng_seed"])
        for it in range(cfg["iters"]):
            cur = {k: v.detach().clone() for k, v in svi.params.items()}
            noise = O.draw_noise(cur, data, ndx, fdx)
            svi.step(ndx, fdx, noise)
            loss = eng.step(**replay_args(eng, data, cur, ndx, fdx, noise, torch.float64)).item()
            ref = case["losses"][it].item()
            assert abs(loss - ref) <= 1e-9 * abs(ref), (it, loss, ref)
    finally:
        torch.set_rng_state(state)
    assert eng.iteration == cfg["iters"]
    ours = eng.named_unconstrained()
    for k, v in case["final"].items():
        err = (ours[k].double().cpu().reshape(-1) - v.reshape(-1)).abs().max().item()
        assert err <= 1e-7 * max(1.0, v.abs().max().item()), (k, err)


def test_hmm_compute_stats_writes_reference_files(tmp_path):
    """`tapqir stats` for cosmos+hmm: credible intervals of hmm's own latents (``init``, ``trans``; hmm.py:70-81) next to
    the shared ones, posterior summaries and the three files (stats.py:131-258).  The interval arithmetic itself is pinned
    on the CPU (tests/test_stats_cpu.py)."""
    import pandas as pd

    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    save(simulate(4, 12, C=1, P=14, seed=3, params={"kon": 0.2, "koff": 0.2}), tmp_path)
    model = models["cosmos+hmm"](device="cuda", dtype="float")
    model.load(tmp_path)
    model.init(lr=0.005, nbatch_size=4)
    model.run(20, progress_bar=lambda it: it)
    model.compute_stats(CI=0.95, save_matlab=True)
    params = torch.load(tmp_path / "cosmos+hmm_params.tpqr", weights_only=False)
    Nt, F, K, Q = 4, 12, 2, 1
    for name, shape in (("gain", ()), ("init", (Q, 2)), ("trans", (Q, 2, 2)), ("lamda", (Q,)), ("proximity", ()),
                        ("background", (Nt, F, Q)), ("height", (K, Nt, F, Q)), ("x", (K, Nt, F, Q))):
        for key in ("LL", "UL", "Mean"):
            assert tuple(params[name][key].shape) == shape, (name, key)
        assert (params[name]["LL"] <= params[name]["Mean"]).all() and (params[name]["Mean"] <= params[name]["UL"]).all()
    assert params["m_probs"].shape == (K, Nt, F, Q) and params["z_probs"].shape == (Nt, F, Q, 2)
    assert params["theta_probs"].shape == (K, Nt, F, Q) and params["z_map"].shape == (Nt, F, Q)
    assert params["z_trans"].shape == (Nt, F, Q, 2, 2)                                   # hmm.py:669-676
    summary = pd.read_csv(tmp_path / "cosmos+hmm_summary.csv", index_col=0)
    assert {"gain", "proximity", "lamda", "trans", "SNR_0", "MCC"} <= set(summary.index)
    assert (tmp_path / "cosmos+hmm_params.mat").exists()


def test_read_glimpse_matches_the_reference_read_glimpse(tmp_path):
    """The CUDA ingestion (tq_crop_aois, tq_offset_hist) on the synthetic movie of tests/golden/ref_glimpse_folder/ against
    ``out/data.tpqr``, which the reference's own ``read_glimpse`` wrote for it (glimpse_reader.py:304-472, run verbatim by
    tests/golden/make_golden_step.py): patches, target positions, offset samples / weights, labels -- bit for bit."""
    from pathlib import Path

    import numpy as np

    from tapqir_b200.imscroll import read_glimpse
    from tapqir_b200.utils.dataset import load

    folder = Path(__file__).resolve().parent / "golden" / "ref_glimpse_folder"
    ref = load(folder / "out")
    kwargs = {"P": 14, "num-channels": 1, "dataset": "golden-movie", "offset-P": 10, "bin-size": 3, "offset-x": 2, "offset-y": 3,
              "use-offtarget": True, "frame-range": True, "frame-start": 3, "frame-end": 10, "labels": True,
              "channels": [{"name": "green", "glimpse-folder": str(folder / "glimpse"), "driftlist": str(folder / "driftlist.mat"),
                            "ontarget-aoiinfo": str(folder / "ontarget_aoiinfo2.mat"),
                            "offtarget-aoiinfo": str(folder / "offtarget_aoifits.mat"),
                            "ontarget-labels": str(folder / "intervals.mat"), "offtarget-labels": None}]}
    ds = read_glimpse(tmp_path, None, **kwargs)
    assert np.array_equal(ds.images.cpu().numpy().astype(np.int64), ref.images.numpy())
    assert np.array_equal(ds.xy.cpu().numpy(), ref.xy.numpy())
    assert torch.equal(ds.offset.samples.cpu(), ref.offset.samples) and torch.equal(ds.offset.weights.cpu(), ref.offset.weights)
    assert ds.is_ontarget.tolist() == ref.is_ontarget.tolist() and ds.name == ref.name and ds.channels == ref.channels
    assert np.array_equal(ds.labels, ref.labels)
    assert torch.equal(torch.as_tensor(ds.ttb).double().reshape(-1), ref.ttb.reshape(-1))
