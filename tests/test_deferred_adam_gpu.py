"""
GPU: the deferred AOI-local Adam update (tq_cosmos_sites_adam: the dense update of step t applied by the site kernel of
step t + 1, engine.deferred_adam) against the separate tq_adam_dense launch at the end of every step.  Same formula, same
constants, explicit fused multiply-adds in both kernels: parameters, both moments, global parameters and every loss must
come out BIT-IDENTICAL, whenever the pending update is flushed (parameter reads, gradient-only steps, a change of the
minibatch shape, checkpoints).
"""

import pytest
import torch

pytestmark = pytest.mark.gpu


def make_model(N, F, C=1, deferred=True, use_graph=True, nb=None, fb=None, seed=3):
    from tapqir_b200.models import models
    from tapqir_b200.utils.simulate import simulate

    model = models["cosmos"](device="cuda", dtype="float")
    model.data = simulate(N, F, C=C, P=14, seed=seed, device="cuda")
    model.init(lr=0.005, nbatch_size=nb or N, fbatch_size=fb or F)
    eng = model.engine
    eng.deferred_adam = deferred and eng.deferred_adam
    eng.deferred_min_units = 0   # (by default only launches of >= 2^20 units defer: no gain below)
    eng.use_graph = use_graph
    return model


def snapshot(eng):
    return [t.clone() for t in (eng.lparams, eng.lm, eng.lv, eng.gparams, eng.gm, eng.gv)]


def assert_same(a, b, what):
    for name, x, y in zip(("lparams", "exp_avg", "exp_avg_sq", "gparams", "g_exp_avg", "g_exp_avg_sq"), a, b):
        assert torch.equal(x, y), f"{what}: {name} differs in {(x != y).sum().item()} of {x.numel()} entries"


@pytest.mark.parametrize("shape", [dict(N=5, F=100), dict(N=4, F=130, C=2), dict(N=3, F=257)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_deferred_update_is_bit_identical(shape, use_graph):
    runs = []
    for deferred in (True, False):
        model = make_model(deferred=deferred, use_graph=use_graph, **shape)
        eng = model.engine
        assert eng.deferred_adam == deferred
        losses = []
        for it in range(7):
            losses.append(eng.step().clone())
            assert eng._pending == deferred
            if it == 3:
                mid = snapshot(eng)            # reading the parameters applies the pending update ...
                assert not eng._pending
                eng.step(update=False)         # ... a gradient-only step leaves them alone ...
                assert_same(mid, snapshot(eng), "gradient-only step")
        assert eng.iteration == 7
        runs.append((torch.cat(losses), snapshot(eng), mid))
    (l1, s1, m1), (l0, s0, m0) = runs
    assert torch.equal(l1, l0), (l1 - l0)
    assert_same(m1, m0, "after 4 steps")
    assert_same(s1, s0, "after 7 steps")


def test_deferred_update_moves_every_owned_entry_once():
    """One step from the initial point: every local entry has moved by exactly one Adam step (|dp| = lr for a first step
    with a non-zero gradient), i.e. the site threads cover their range once and tq_adam_dense the rest once."""
    model = make_model(6, 150)
    eng = model.engine
    before = eng.lparams.clone()
    eng.step()
    assert eng._pending
    lo, hi = eng._deferred_range
    raw = eng._lparams   # not flushed: the owned range still holds the old values, the rest is updated
    assert torch.equal(raw[lo:hi], before[lo:hi])
    moved = (raw != before)
    assert bool(moved[:lo].all()) and moved[hi:].float().mean().item() > 0.99
    after = eng.lparams  # flush
    assert not eng._pending
    step = (after - before).abs()
    nz = eng.lgrads.abs() > 1e-3   # (|dp| = lr |g| / (|g| + eps) on the first step)
    assert torch.allclose(step[nz], torch.full_like(step[nz], 0.005), rtol=1e-3)
    again = eng.lparams.clone()   # a second read must not apply anything
    assert torch.equal(after, again)


def test_switching_minibatch_shape_flushes():
    """full-batch steps (deferred), then subsampled steps (dense update at the end of the step), then full-batch again"""
    runs = []
    for deferred in (True, False):
        model = make_model(8, 120, deferred=deferred, use_graph=False)
        eng = model.engine
        for _ in range(3):
            eng.step()
        eng.set_batch(4, 50)
        assert not eng._pending
        for _ in range(3):
            eng.step()
            assert not eng._pending
        eng.set_batch(8, 120)
        for _ in range(2):
            eng.step()
        runs.append(snapshot(eng))
    assert_same(runs[0], runs[1], "full -> subsampled -> full")


def test_large_launch_uses_the_eight_units_per_thread_kernel():
    """enough units for site_fast_kernel<8> (>= 4 waves of 8 blocks per SM): same bits as the separate update"""
    runs = []
    for deferred in (True, False):
        model = make_model(120, 5000, deferred=deferred, use_graph=True, seed=1)
        eng = model.engine
        for _ in range(4):
            loss = eng.step()
        runs.append((loss.clone(), snapshot(eng)))
        del model, eng
        torch.cuda.empty_cache()
    assert torch.equal(runs[0][0], runs[1][0])
    assert_same(runs[0][1], runs[1][1], "600 000 units, 4 steps")


def test_checkpoint_round_trip_with_a_pending_update(tmp_path):
    from tapqir_b200.models import models
    from tapqir_b200.utils.dataset import save
    from tapqir_b200.utils.simulate import simulate

    save(simulate(3, 40, C=1, P=14, seed=0), tmp_path)
    model = models["cosmos"](device="cuda", dtype="float")
    model.load(tmp_path)
    model.init(lr=0.005, nbatch_size=3, fbatch_size=40)
    model.engine.deferred_min_units = 0
    model.run(5, progress_bar=lambda it: it)
    assert model.engine.deferred_adam and model.engine._pending
    model.save_checkpoint()
    ref = snapshot(model.engine)
    again = models["cosmos"](device="cuda", dtype="float")
    again.load(tmp_path)
    again.init(lr=0.005, nbatch_size=3, fbatch_size=40)
    assert again.iter == model.iter
    assert_same(ref, snapshot(again.engine), "restored")
    again.engine.deferred_min_units = 0
    model.run(3, progress_bar=lambda it: it)
    again.run(3, progress_bar=lambda it: it)
    assert_same(snapshot(model.engine), snapshot(again.engine), "continued")
