"""
GPU: the fused step kernel (csrc/cosmos_fused.cu: guide sites -> likelihood -> post in one persistent launch, per-unit
intermediates in shared memory) against the three per-stage kernels it replaces (site_fast / ksmogn_stream / local_post
with their HBM scratch), on the same inputs: the per-unit arithmetic is the same code, so every gradient of a minibatch
unit must come out bit-identical and the cross-unit sums equal to rounding (the two forms add their block partials in
different, each fixed, orders).  The fused kernel is opt-in (TQ_FUSED=1: measured slower, profiles/r2_fused_ab.md);
the oracle parity suites run the per-stage kernels.
"""

import pytest
import torch

from oracle import cosmos_oracle as O
from tapqir_b200.models import layout as L
from tests.step_helpers import flat_inputs, make_problem
from tests.test_step_gpu import make_engine, replay_args

pytestmark = pytest.mark.gpu

CONFIGS = [
    dict(N=7, F=152, C=1, nb=7, fb=152, seed=0),                       # full batch, rows of 2 x 64 + 24 units
    dict(N=9, F=100, C=1, nb=4, fb=37, seed=1),                        # minibatch: gathered AOIs / frames, one partial batch per AOI
    dict(N=5, F=70, C=2, nb=5, fb=70, seed=2),                         # two channels interleaved in a batch
    dict(N=4, F=64, C=1, nb=4, fb=64, seed=3, offsets="hist"),         # 16 offset bins: the tiled-offset form
    dict(N=6, F=92, C=1, nb=6, fb=92, seed=4, perturb=True, scale=0.7),  # far from the initial point: deferred / fallback sites
]


def run_both(cfg, merge_offsets=True, replay=True):
    cfg = dict(cfg)
    scale = cfg.pop("scale", None)
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    if scale:
        g = torch.Generator().manual_seed(99)
        for k in L.LOCAL_NAMES:
            params[k].add_(scale * torch.randn(params[k].shape, generator=g, dtype=params[k].dtype))
        noise = O.draw_noise(params, data, ndx, fdx, g)
    params = {k: v.float().double() for k, v in params.items()}
    noise = {k: v.float().double() for k, v in noise.items()}
    out = []
    for fused in (True, False):
        eng = make_engine(ds, data, params, cfg["nb"], cfg["fb"], torch.float32, merge_offsets=merge_offsets, seed=5)
        eng.fused = fused
        eng.keep_intermediates = True
        if replay:
            loss = eng.step(update=False, **replay_args(eng, data, params, ndx, fdx, noise, torch.float32)).item()
        else:
            n, f = ndx.to(torch.int32).cuda(), fdx.to(torch.int32).cuda()
            loss = eng.step(update=False, ndx=n, fdx=f).item()   # device-drawn variates (Philox keyed by unit identity)
        assert eng.last_step_fused == fused
        out.append((loss, eng.lgrads.clone(), eng.ggrads.clone(), eng.acc.clone(), eng.samples.clone(), eng.Lm.clone()))
    return out


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("replay", [True, False])
def test_fused_kernel_equals_the_per_stage_kernels(cfg, replay):
    (lf, gf, ggf, af, sf, Lf), (ls, gs, ggs, as_, ss, Ls) = run_both(cfg, replay=replay)
    assert bool(torch.isfinite(gs).all()) and bool(torch.isfinite(gf).all())
    assert torch.equal(sf, ss), "guide samples differ"
    ll = L.LocalLayout(cfg["N"], cfg["F"], cfg["C"])
    vf, vs = ll.views(gf), ll.views(gs)
    # A warp sweeps four consecutive units and picks ONE form of the sweep for them (packed pairs / scalar / small-a).
    # When a minibatch AOI holds a multiple of 4 units both kernels group the same units, so everything per unit is
    # bit-identical; otherwise a unit can be swept by the other form (1e-6 differences).
    # (and far from the initial point, cfg["scale"], the two kernels' separately compiled copies of the regime code
    # outside the bulk forms may contract multiply-adds differently: last-bit differences in a site record)
    # (and with more than 4 offset bins the stand-alone kernel runs the one-pass many-bins form, the fused one the two-pass form)
    aligned = (cfg["fb"] * cfg["C"]) % 4 == 0 and "scale" not in cfg and cfg.get("offsets", "sim") == "sim"
    if aligned:
        assert torch.equal(Lf, Ls), "configuration log-likelihoods differ"
        tol = dict(rtol=1e-12, atol=0)
    else:
        assert torch.allclose(Lf, Ls, rtol=2e-6, atol=0)
        tol = dict(rtol=1e-5, atol=0)
    assert abs(lf - ls) <= (1e-12 if aligned else 1e-6) * abs(ls)
    assert torch.allclose(af, as_, rtol=tol["rtol"], atol=tol["rtol"] * as_.abs().max().item())
    assert torch.allclose(ggf, ggs, rtol=0, atol=(1e-10 if aligned else 1e-5) * ggs.abs().max().item())
    for name in L.LOCAL_NAMES:
        if name in ("background_mean_loc", "background_std_loc"):   # sums over an AOI's frames: same terms, other order
            assert torch.allclose(vf[name], vs[name], rtol=2e-6 if aligned else 1e-5, atol=0), name
        elif aligned:
            assert torch.equal(vf[name], vs[name]), name
        else:
            assert torch.allclose(vf[name], vs[name], rtol=0, atol=1e-5 * vs[name].abs().max().item()), name


def test_fused_kernel_with_three_offset_bins_kept():
    (lf, gf, *_), (ls, gs, *_) = run_both(dict(N=5, F=80, C=1, nb=5, fb=80, seed=6), merge_offsets=False)
    assert abs(lf - ls) <= 1e-12 * abs(ls)
    ll = L.LocalLayout(5, 80, 1)
    for name in L.LOCAL_NAMES[2:]:
        assert torch.equal(ll.views(gf)[name], ll.views(gs)[name]), name


def test_fused_training_steps_follow_the_per_stage_trajectory():
    """200 production steps (device RNG, minibatches drawn on the device, dense Adam, CUDA-graph replay) with either
    form from the same seed: same minibatches, same variates, parameters equal to accumulated rounding of the sums."""
    ds, data, params, _, _, _ = make_problem(N=8, F=120, C=1, nb=8, fb=120, seed=7, perturb=False)
    runs = []
    for fused in (True, False):
        eng = make_engine(ds, data, params, 5, 50, torch.float32, seed=11)
        eng.fused = fused
        losses = [eng.step().item() for _ in range(200)]
        assert eng.last_step_fused == fused
        runs.append((losses, eng.lparams.clone(), eng.gparams.clone()))
    (la, pa, ga), (lb, pb, gb) = runs
    assert all(abs(a - b) <= 1e-6 * abs(b) for a, b in zip(la, lb))
    assert torch.allclose(pa, pb, rtol=0, atol=2e-3) and (pa - pb).abs().mean().item() < 1e-5
    assert torch.allclose(ga, gb, rtol=0, atol=1e-4)


def test_fused_step_rejects_what_it_does_not_support():
    import ctypes

    from tapqir_b200 import _lib

    ds, data, params, _, _, _ = make_problem(N=3, F=5, C=1, nb=3, fb=5, seed=8)
    eng = make_engine(ds, data, params, 3, 5, torch.float64)
    view = eng._view(None, None)
    assert eng.lib.tq_cosmos_fused_supported(_lib.TQ_F64, ctypes.byref(view)) == 0
    eng.step(update=False)
    assert eng.last_step_fused is False
