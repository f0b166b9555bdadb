"""
The reference's own smoke contract (test/test_tapqir.py:53-93: simulate N=2, F=5, P=14; `fit` one
iteration; exit code 0) re-expressed against the new `cosmos` class, plus the on-disk state formats.
"""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def dataset_path(tmp_path):
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    data = simulate(2, 5, C=1, P=14, seed=0)   # constants of test_tapqir.py:25-40 are simulate()'s defaults
    save(data, tmp_path)
    return tmp_path


@pytest.mark.parametrize("dtype", ["float", "double"])
def test_fit_one_iteration_and_checkpoint(dataset_path, dtype):
    from tapqir_b200.models import models

    model = models["cosmos"](device="cuda", dtype=dtype)
    model.load(dataset_path)
    assert (model.data.Nt, model.data.F, model.data.C, model.data.P) == (2, 5, 1, 14)
    model.init(lr=0.005, nbatch_size=2, fbatch_size=5)
    model.run(1, progress_bar=lambda it: it)
    assert model.iter == 1 and model.iter_loss == model.iter_loss  # finite, not NaN
    ckpt_file = dataset_path / ".tapqir" / "cosmos_model.tpqr"
    assert ckpt_file.exists()
    ckpt = torch.load(ckpt_file, weights_only=False)
    # layout of models/model.py:273-282
    assert set(ckpt) == {"iter", "params", "optimizer", "rolling", "convergence_status"}
    assert set(ckpt["params"]) == {"params", "constraints"}
    names = set(ckpt["params"]["params"])
    assert names == {"pi_mean", "pi_size", "m_probs", "proximity_loc", "proximity_size", "lamda_loc", "lamda_beta",
                     "gain_loc", "gain_beta", "background_mean_loc", "background_std_loc", "b_loc", "b_beta", "h_loc",
                     "h_beta", "w_mean", "w_size", "x_mean", "y_mean", "size"}
    assert ckpt["params"]["params"]["h_loc"].shape == (2, 2, 5, 1)
    assert ckpt["params"]["params"]["background_mean_loc"].shape == (2, 1, 1)
    assert ckpt["params"]["params"]["pi_mean"].shape == (1, 2)
    st = ckpt["optimizer"]["h_loc"]
    assert set(st) == {"state", "param_groups"} and set(st["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert "-ELBO" in ckpt["rolling"]
    assert (dataset_path / ".tapqir" / "logs" / "cosmos").exists()   # tensorboard directory (model.py:209)

    # resume: a fresh model picks the checkpoint up in init() (model.py:173-180)
    again = models["cosmos"](device="cuda", dtype=dtype)
    again.load(dataset_path)
    again.init(lr=0.005, nbatch_size=2, fbatch_size=5)
    assert again.iter == ckpt["iter"]
    a, b = model.engine.named_unconstrained(), again.engine.named_unconstrained()
    # checkpoint was written at iteration 0 (before the counter increments): parameters after 1 step
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(model.engine.lm, again.engine.lm)
    again.run(2, progress_bar=lambda it: it)
    assert again.iter == ckpt["iter"] + 2


def test_missing_data_raises_reference_exception(tmp_path):
    from tapqir_b200.exceptions import TapqirFileNotFoundError
    from tapqir_b200.models import models

    model = models["cosmos"](device="cuda")
    with pytest.raises(TapqirFileNotFoundError) as err:
        model.load(tmp_path)
    assert err.value.name == "data"


def test_nan_parameters_raise_value_error_at_checkpoint(dataset_path):
    """model.py:246-250: NaN/Inf in any parameter -> ValueError at checkpoint time."""
    from tapqir_b200.models import models

    model = models["cosmos"](device="cuda")
    model.load(dataset_path)
    model.init(nbatch_size=2, fbatch_size=5)
    model.step()
    model.engine.named_unconstrained()["gain_loc"].fill_(float("nan"))
    with pytest.raises(ValueError, match="gain_loc"):
        model.save_checkpoint()


def test_minibatch_run_with_device_subsampling(dataset_path):
    """nbatch/fbatch smaller than the data: indices drawn on the device each step, dense Adam."""
    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    save(simulate(6, 40, seed=1), dataset_path)
    model = models["cosmos"](device="cuda")
    model.load(dataset_path)
    model.init(nbatch_size=3, fbatch_size=16)
    before = model.engine.lparams.clone()
    losses = [model.step().item() for _ in range(30)]
    assert all(l == l for l in losses)
    ndx = model.engine.ndx.cpu()
    assert len(set(ndx.tolist())) == 3 and ndx.min() >= 0 and ndx.max() < 6
    fdx = model.engine.fdx.cpu()
    assert len(set(fdx.tolist())) == 16 and fdx.max() < 40
    assert (model.engine.lparams != before).float().mean().item() > 0.9   # dense update: (almost) every entry moved


def test_compute_stats_writes_reference_files(dataset_path):
    """Row N2: compute_stats -> cosmos_params.tpqr / .mat / cosmos_summary.csv with the reference's
    keys and shapes (stats.py:131-258, cosmos.py:711-784), as `tapqir stats --matlab` does."""
    import pandas as pd
    import scipy.stats as st

    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    save(simulate(4, 30, seed=2), dataset_path)
    model = models["cosmos"](device="cuda")
    model.load(dataset_path)
    model.init(nbatch_size=4, fbatch_size=30)
    model.run(50, progress_bar=lambda it: it)
    model.compute_stats(CI=0.95, save_matlab=True)
    params = torch.load(dataset_path / "cosmos_params.tpqr", weights_only=False)
    Nt, F, K, Q = 4, 30, 2, 1
    for name, shape in (("gain", ()), ("pi", (Q, 2)), ("lamda", (Q,)), ("proximity", ()), ("background", (Nt, F, Q)),
                        ("height", (K, Nt, F, Q)), ("width", (K, Nt, F, Q)), ("x", (K, Nt, F, Q)), ("y", (K, Nt, F, Q))):
        assert set(params[name]) >= {"LL", "UL", "Mean"}
        for key in ("LL", "UL", "Mean"):
            assert tuple(params[name][key].shape) == shape, (name, key)
        assert (params[name]["LL"] <= params[name]["Mean"]).all() and (params[name]["Mean"] <= params[name]["UL"]).all()
    assert params["m_probs"].shape == (K, Nt, F, Q) and params["z_probs"].shape == (Nt, F, Q, 2)
    assert params["theta_probs"].shape == (K, Nt, F, Q) and params["z_map"].shape == (Nt, F, Q)
    assert torch.allclose(params["p_specific"], params["theta_probs"].sum(0))
    assert params["chi2"]["values"].shape == (Nt, F, Q) and {"vmin", "vmax"} <= set(params["height"])
    # interval of one Gamma site against scipy directly
    loc, beta = model.param("gain_loc").item(), model.param("gain_beta").item()
    lo, hi = st.gamma(loc * beta, scale=1 / beta).interval(0.95)
    assert abs(params["gain"]["LL"].item() - lo) < 1e-9 and abs(params["gain"]["UL"].item() - hi) < 1e-9
    summary = pd.read_csv(dataset_path / "cosmos_summary.csv", index_col=0)
    assert list(summary.columns) == ["Mean", "95% LL", "95% UL"]
    assert {"gain", "proximity", "lamda", "pi", "SNR_0", "MCC", "Recall", "Precision", "TN", "FP", "FN", "TP",
            "p(specific)"} <= set(summary.index)
    assert (dataset_path / "cosmos_params.mat").exists()
    # load(data_only=False) picks both files up (model.py:108-126)
    again = models["cosmos"](device="cuda")
    again.load(dataset_path, data_only=False)
    assert "z_probs" in again.params and "gain" in again.summary.index


def test_snr_and_chi2_matches_formula():
    from tapqir_b200.utils.stats import snr_and_chi2

    g = torch.Generator().manual_seed(0)
    K, F, Q, P = 2, 5, 1, 14
    r = lambda *s: torch.rand(*s, generator=g)
    h, w, x, y = 1000 + 2000 * r(K, F, Q), 1.2 + r(K, F, Q), 2 * r(K, F, Q) - 1, 2 * r(K, F, Q) - 1
    tgt = torch.full((F, Q, 2), 6.5)
    b = 100 + 50 * r(F, Q)
    data = 200 + 100 * r(F, Q, P, P)
    snr, chi2 = snr_and_chi2(data.cuda(), h.cuda(), w.cuda(), x.cuda(), y.cuda(), tgt.cuda(), b.cuda(), 7.0, 90.0, 0.0, P, None)
    from oracle import cosmos_oracle as O

    ga = O.gaussian_spots(h.double()[..., None], w.double()[..., None], x.double()[..., None], y.double()[..., None],
                          tgt.double()[None, :, :, None, :], P)[..., 0, :, :]   # (K,F,Q,P,P)
    sig = ((data.double() - b.double()[..., None, None] - 90.0) * (ga / h.double()[..., None, None])).sum((-1, -2))
    ref_snr = sig / (0.0 + b.double() * 7.0).sqrt()
    ideal = b.double()[..., None, None] + ga.sum(0)
    ref_chi2 = ((data.double() - ideal - 90.0) ** 2 / ideal).mean((-1, -2))
    assert torch.allclose(snr.double().cpu(), ref_snr, rtol=1e-4) and torch.allclose(chi2.double().cpu(), ref_chi2, rtol=1e-4)


def test_stats_flow_on_a_model_that_was_never_initialised(dataset_path):
    """`tapqir stats` (main.py:566): ``load(); load_checkpoint(param_only=True); compute_stats()`` on a FRESH model --
    no ``init()``: the engine is built from the constructor's defaults."""
    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    save(simulate(4, 12, seed=5), dataset_path)
    fit = models["cosmos"](device="cuda")
    fit.load(dataset_path)
    fit.init(nbatch_size=4, fbatch_size=12)
    fit.run(1, progress_bar=lambda it: it)
    stats = models["cosmos"](device="cuda")
    stats.load(dataset_path)
    stats.load_checkpoint(param_only=True)
    for k, v in fit.engine.named_unconstrained().items():
        assert torch.equal(v, stats.engine.named_unconstrained()[k]), k
    stats.compute_stats(CI=0.95)
    assert (dataset_path / "cosmos_params.tpqr").exists() and (dataset_path / "cosmos_summary.csv").exists()


def test_reported_loss_carries_the_masked_aoi_constant(dataset_path):
    """model.py:285-298 logs ``-ELBO`` as Pyro computes it: for a masked AOI the enumerated sites' log-probabilities are
    zeroed but still summed over -- 4 configurations x ln 6 per masked unit, times the plate scales (tests/test_oracle.py
    pins the constant against the reference's own run).  ``iter_loss`` adds it to the device loss; gradients are unaffected."""
    import math

    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    ds = simulate(4, 10, seed=6)
    ds.mask[1] = False
    save(ds, dataset_path)
    model = models["cosmos"](device="cuda")
    model.load(dataset_path)
    model.init(nbatch_size=4, fbatch_size=5)
    model.step()
    raw = float(model._loss_dev.item())
    assert abs(model.iter_loss - raw - 1 * 5 * 1 * 4 * math.log(6) * (4 / 4) * (10 / 5)) < 1e-9
    model.init(nbatch_size=2, fbatch_size=10)          # AOI minibatch: the constant follows the drawn indices
    model.step()
    n_masked = int((model.engine.ndx.cpu() == 1).sum())
    assert abs(model.iter_loss - float(model._loss_dev.item()) - n_masked * 10 * 4 * math.log(6) * (4 / 2)) < 1e-9


def test_nan_detection_has_its_own_exception_type(dataset_path):
    """Model.run restarts only on NonFiniteParameterError (a ValueError, as in the reference): an argument error of the
    C ABI -- also a ValueError -- must propagate instead of being retried as a divergence."""
    from tapqir_b200.exceptions import NonFiniteParameterError
    from tapqir_b200.models import models

    model = models["cosmos"](device="cuda")
    model.load(dataset_path)
    model.init(nbatch_size=2, fbatch_size=5)
    model.step()
    model.engine.named_unconstrained()["gain_loc"].fill_(float("nan"))
    with pytest.raises(NonFiniteParameterError):
        model.save_checkpoint()
    model.init(nbatch_size=2, fbatch_size=5)

    def broken_step(**kw):
        raise ValueError("tq_subsample: bad sizes")

    model.step = broken_step
    with pytest.raises(ValueError, match="bad sizes"):
        model.run(3, progress_bar=lambda it: it)
