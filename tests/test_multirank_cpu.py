"""
CPU, world_size 2 over gloo: the N>1 host logic of the AOI-sharded step.

Each rank evaluates ITS contiguous AOI block (the kernels' arithmetic compiled for the host,
tests/hostcheck), the (C, 18) accumulator vector is all-reduced -- the only exchange of the path,
SURVEY.md 8(e) -- and the replicated global reverse pass then yields the same loss and global
gradients as the single-process oracle on the whole minibatch; local gradients equal the oracle's
slices for the rank's AOIs.
"""

import ctypes
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cosmos_oracle as O
from tapqir_b200.models import layout as L
from tests.step_helpers import compare_grads, make_problem


def _worker(rank, world, port, cfg, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import hostcheck
        from tests.step_helpers import host_step

        hc = hostcheck.load()
        ds, data, params, ndx, fdx, noise = make_problem(**cfg)
        Nt, per = data.Nt, data.Nt // world
        lo, hi = rank * per, (rank + 1) * per
        # this rank's shard: AOIs [lo, hi) and the minibatch AOIs that fall inside it
        sel = (ndx >= lo) & (ndx < hi)
        shard = O.OracleData(data.images[lo:hi], data.xy[lo:hi], data.is_ontarget[lo:hi], data.mask[lo:hi],
                             data.offset_samples, data.offset_weights)
        sparams = {k: (v[:, lo:hi] if v.dim() == 4 else (v[lo:hi] if v.dim() == 3 else v)) for k, v in params.items()}
        snoise = {k: v for k, v in noise.items()}
        for k in ("background",):
            snoise[k] = noise[k][sel]
        for k in ("height", "width", "x", "y"):
            snoise[k] = noise[k][:, sel]
        sndx = ndx[sel] - lo
        nb_total = len(ndx)
        # host_step derives sN from the shard; override with the global plate scale
        import tests.step_helpers as SH

        ll, gl, lparams, gparams, lnoise, gnoise = SH.flat_inputs(shard, sparams, snoise, torch.float64)
        mc = L.ModelConst.make(O.DEFAULT_PRIORS, shard.P, torch.float64)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        nb, fb = len(sndx), len(fdx)
        U = nb * fb * shard.C
        lgrads = torch.empty_like(lparams)
        ggrads = torch.empty(gl.numel, dtype=torch.float64)
        acc = torch.zeros(shard.C * L.NACC, dtype=torch.float64)
        samples = torch.empty(L.NSAMP, U, dtype=torch.float64)
        sN, sF = Nt / nb_total, data.F / fb
        hc.hc_cosmos_step_f64.restype = ctypes.c_double
        n32, f32 = sndx.to(torch.int32).contiguous(), fdx.to(torch.int32).contiguous()
        pix, xy = shard.images.contiguous(), shard.xy.contiguous()
        ont, mask = shard.is_ontarget.to(torch.uint8).contiguous(), shard.mask.to(torch.uint8).contiguous()
        off_s, off_w = shard.offset_samples.contiguous(), shard.offset_logits.contiguous()
        hc.hc_cosmos_step_f64(nb, fb, shard.Nt, shard.F, shard.C, shard.P, off_s.numel(), p(n32), p(f32), p(pix), p(xy), p(ont),
                              p(mask), p(off_s), p(off_w), ctypes.byref(mc), ctypes.c_double(sN), ctypes.c_double(sF),
                              p(lparams), p(gparams), p(lnoise), p(gnoise), p(lgrads), p(ggrads), p(acc), p(samples))
        dist.all_reduce(acc)   # the one collective of the path
        hc.hc_globals_post.restype = ctypes.c_double
        loss = hc.hc_globals_post(shard.C, ctypes.byref(mc), p(gparams), p(gnoise), p(acc), ctypes.c_double(sN),
                                  ctypes.c_double(sF), p(ggrads))
        # numpy copies: torch tensors travel through mp queues as shared-memory handles that die with the worker
        grads = {k: v.numpy().copy() for k, v in ll.views(lgrads).items()}
        grads.update({k: v.numpy().copy() for k, v in gl.views(ggrads).items()})
        out_q.put((rank, loss, grads, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [dict(N=4, F=6, C=1, nb=4, fb=4, seed=0), dict(N=6, F=5, C=2, nb=4, fb=3, seed=1)])
def test_two_rank_sharded_step_equals_single_process_oracle(cfg):
    world = 2
    ds, data, params, ndx, fdx, noise = make_problem(**cfg)
    ref_loss, ref_grads = O.loss_and_grads(params, data, ndx, fdx, noise)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, loss, grads, lo, hi in results:
        grads = {k: torch.from_numpy(v) for k, v in grads.items()}
        assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)          # identical on every rank
        assert not compare_grads(grads, ref_grads, 1e-8, names=L.GLOBAL_NAMES)
        for k in L.LOCAL_NAMES:
            ref = ref_grads[k][:, lo:hi] if ref_grads[k].dim() == 4 else ref_grads[k][lo:hi]
            err = (grads[k] - ref).abs().max().item()
            assert err <= 1e-8 * max(ref_grads[k].abs().max().item(), 1e-30), (k, err)


# ---- unequal shards (Nt not divisible by the number of ranks): stratified minibatch, per-rank plate scale ---------------------
def _uneven_plan(cfg, world):
    """What models/cosmos.py::_shard_sizes and engine.set_batch derive: balanced contiguous blocks, nb_r = min(nbatch, Nt_r),
    sN_r = Nt_r / nb_r, sN_ref = Nt / sum nb_r; each rank's minibatch = the first nb_r entries of a seeded permutation."""
    N = cfg["N"]
    sizes = [N // world + (1 if r < N % world else 0) for r in range(world)]
    los = [sum(sizes[:r]) for r in range(world)]
    nbs = [min(cfg["nb"], n) for n in sizes]
    g = torch.Generator().manual_seed(cfg["seed"] + 77)
    ndxs = [torch.randperm(n, generator=g)[:b] for n, b in zip(sizes, nbs)]
    fdx = torch.randperm(cfg["F"], generator=g)[:cfg["fb"]]
    return sizes, los, nbs, ndxs, fdx


def _uneven_problem(cfg, world):
    ds, data, params, _, _, _ = make_problem(N=cfg["N"], F=cfg["F"], C=cfg["C"], nb=1, fb=1, seed=cfg["seed"])
    sizes, los, nbs, ndxs, fdx = _uneven_plan(cfg, world)
    g = torch.Generator().manual_seed(cfg["seed"] + 78)
    all_ndx = torch.cat([lo + n for lo, n in zip(los, ndxs)])
    noise = O.draw_noise(params, data, all_ndx, fdx, g)
    return data, params, sizes, los, nbs, ndxs, fdx, noise


def _uneven_worker(rank, world, port, cfg, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tests.step_helpers as SH
        from tests import hostcheck

        hc = hostcheck.load()
        data, params, sizes, los, nbs, ndxs, fdx, noise = _uneven_problem(cfg, world)
        lo, hi = los[rank], los[rank] + sizes[rank]
        pos = sum(nbs[:rank])
        shard = O.OracleData(data.images[lo:hi], data.xy[lo:hi], data.is_ontarget[lo:hi], data.mask[lo:hi],
                             data.offset_samples, data.offset_weights)
        sparams = {k: (v[:, lo:hi] if v.dim() == 4 else (v[lo:hi] if v.dim() == 3 else v)) for k, v in params.items()}
        snoise = dict(noise)
        snoise["background"] = noise["background"][pos:pos + nbs[rank]]
        for k in ("height", "width", "x", "y"):
            snoise[k] = noise[k][:, pos:pos + nbs[rank]]
        ll, gl, lparams, gparams, lnoise, gnoise = SH.flat_inputs(shard, sparams, snoise, torch.float64)
        mc = L.ModelConst.make(O.DEFAULT_PRIORS, shard.P, torch.float64)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        nb, fb = nbs[rank], len(fdx)
        lgrads, ggrads = torch.empty_like(lparams), torch.empty(gl.numel, dtype=torch.float64)
        acc = torch.zeros(shard.C * L.NACC, dtype=torch.float64)
        samples = torch.empty(L.NSAMP, nb * fb * shard.C, dtype=torch.float64)
        sN, sN_ref, sF = sizes[rank] / nb, data.Nt / sum(nbs), data.F / fb      # engine.set_batch
        n32, f32 = ndxs[rank].to(torch.int32).contiguous(), fdx.to(torch.int32).contiguous()
        pix, xy = shard.images.contiguous(), shard.xy.contiguous()
        ont, mask = shard.is_ontarget.to(torch.uint8).contiguous(), shard.mask.to(torch.uint8).contiguous()
        off_s, off_w = shard.offset_samples.contiguous(), shard.offset_logits.contiguous()
        hc.hc_cosmos_step_f64.restype = ctypes.c_double
        hc.hc_cosmos_step_f64(nb, fb, shard.Nt, shard.F, shard.C, shard.P, off_s.numel(), p(n32), p(f32), p(pix), p(xy), p(ont),
                              p(mask), p(off_s), p(off_w), ctypes.byref(mc), ctypes.c_double(sN), ctypes.c_double(sF),
                              p(lparams), p(gparams), p(lnoise), p(gnoise), p(lgrads), p(ggrads), p(acc), p(samples))
        acc.mul_(sN / sN_ref)   # engine._enqueue: acc_weight
        dist.all_reduce(acc)
        hc.hc_globals_post.restype = ctypes.c_double
        loss = hc.hc_globals_post(shard.C, ctypes.byref(mc), p(gparams), p(gnoise), p(acc), ctypes.c_double(sN_ref),
                                  ctypes.c_double(sF), p(ggrads))
        grads = {k: v.numpy().copy() for k, v in ll.views(lgrads).items()}
        grads.update({k: v.numpy().copy() for k, v in gl.views(ggrads).items()})
        out_q.put((rank, loss, grads, lo, hi))
    finally:
        dist.destroy_process_group()


def test_two_rank_unequal_shards_follow_the_stratified_estimator():
    """5 AOIs over 2 ranks (3 + 2), 2 AOIs drawn per rank: every rank's terms carry ITS plate scale Nt_r / nb_r (1.5 and
    1.0) -- one common scale would bias the ELBO towards the smaller shard.  Reference: the oracle's
    terms per shard combined as  e_global + sum_r sN_r (e_aoi_r + sF e_frame_r)  under autograd."""
    cfg, world = dict(N=5, F=6, C=1, nb=2, fb=4, seed=3), 2
    data, params, sizes, los, nbs, ndxs, fdx, noise = _uneven_problem(cfg, world)
    assert sizes == [3, 2] and nbs == [2, 2]
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    total, sF = 0.0, data.F / len(fdx)
    for r in range(world):
        pos = sum(nbs[:r])
        rn = dict(noise)
        rn["background"] = noise["background"][pos:pos + nbs[r]]
        for k in ("height", "width", "x", "y"):
            rn[k] = noise[k][:, pos:pos + nbs[r]]
        _, parts = O.elbo(leaves, data, los[r] + ndxs[r], fdx, rn, return_parts=True)
        if r == 0:
            total = parts["e_global"]
        total = total + (sizes[r] / nbs[r]) * (parts["e_aoi"] + sF * parts["e_frame"])
    (-total).backward()
    ref_loss, ref_grads = -total.item(), {k: v.grad for k, v in leaves.items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_uneven_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, loss, grads, lo, hi in results:
        grads = {k: torch.from_numpy(v) for k, v in grads.items()}
        assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
        assert not compare_grads(grads, ref_grads, 1e-8, names=L.GLOBAL_NAMES)
        for k in L.LOCAL_NAMES:
            ref = ref_grads[k][:, lo:hi] if ref_grads[k].dim() == 4 else ref_grads[k][lo:hi]
            err = (grads[k] - ref).abs().max().item()
            assert err <= 1e-8 * max(ref_grads[k].abs().max().item(), 1e-30), (k, err)


def test_engine_plate_scales_for_unequal_shards():
    """engine.set_batch arithmetic without a GPU: sN_r, sN_ref and the accumulator weights for 3 + 2 AOIs, nbatch 2."""
    from tapqir_b200.models.cosmos import cosmos

    class FakeData:
        Nt = 5

    sizes = []
    for r in range(2):
        m = cosmos.__new__(cosmos)
        m.data, m.world_size, m.rank, m.presharded = FakeData(), 2, r, False
        sizes.append(m._shard_sizes())
        sl = m._shard()
        assert (sl.start, sl.stop) == ((0, 3) if r == 0 else (3, 5))
    assert sizes[0] == sizes[1] == [3, 2]
    nbs = [min(2, n) for n in sizes[0]]
    sN_ref = 5 / sum(nbs)
    weights = [(n / b) / sN_ref for n, b in zip(sizes[0], nbs)]
    assert weights == [1.2, 0.8] and abs(sum(w * b for w, b in zip(weights, nbs)) * sN_ref - 5) < 1e-12


def test_aoi_sharding_covers_every_aoi_once():
    """cosmos._shard: contiguous blocks, disjoint, complete (also when Nt is not divisible)."""
    from tapqir_b200.models.cosmos import cosmos

    class FakeData:
        def __init__(self, Nt):
            self.Nt = Nt

    for Nt, world in [(100, 8), (7, 2), (5, 4), (1000, 8), (3, 4)]:
        seen = []
        for r in range(world):
            m = cosmos.__new__(cosmos)
            m.data, m.world_size, m.rank, m.presharded = FakeData(Nt), world, r, False
            sl = m._shard()
            seen += list(range(sl.start, sl.stop))
            sizes = m._shard_sizes()
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == Nt      # balanced
        assert seen == list(range(Nt))


# ---- hmm variant: the exchange is (C, NACC) + (C, NHACC) doubles (the chain's expected counts ride along) ---------------
def _hmm_worker(rank, world, port, cfg, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import hmm_oracle as H
        from tests import hostcheck
        from tests.test_hmm_cpu import small_problem

        hc = hostcheck.load()
        ds, data, params, _, _ = small_problem(N=cfg["N"], F=cfg["F"], C=cfg["C"], seed=cfg["seed"])
        ndx = torch.tensor(cfg["ndx"])
        noise = H.draw_noise(params, data, ndx, torch.Generator().manual_seed(cfg["seed"] + 9))
        per = data.Nt // world
        lo, hi = rank * per, (rank + 1) * per
        sel = (ndx >= lo) & (ndx < hi)
        shard = O.OracleData(data.images[lo:hi], data.xy[lo:hi], data.is_ontarget[lo:hi], data.mask[lo:hi],
                             data.offset_samples, data.offset_weights)
        aoi_axis = {"m_probs": 2, "z_trans": 0}
        sparams = {}
        for k, v in params.items():
            if k in H.GLOBAL_PARAMS:
                sparams[k] = v
            else:
                ax = aoi_axis.get(k, 1 if v.dim() == 4 else 0)
                sparams[k] = v.narrow(ax, lo, hi - lo).contiguous()
        snoise = dict(noise)
        snoise["background"] = noise["background"][sel]
        for k in ("height", "width", "x", "y"):
            snoise[k] = noise[k][:, sel]
        sndx = (ndx[sel] - lo).to(torch.int32).contiguous()
        ll, gl = L.HmmLocalLayout(shard.Nt, shard.F, shard.C), L.HmmGlobalLayout(shard.C)
        lparams = torch.zeros(ll.numel, dtype=torch.float64)
        ll.load_named(lparams, sparams)
        gparams = gl.pack({k: params[k] for k in gl.shapes}, dtype=torch.float64)
        lnoise, gnoise = L.pack_local_noise(snoise, torch.float64, "cpu"), gl.pack_noise(noise)
        mc = L.ModelConst.make(O.DEFAULT_PRIORS, shard.P, torch.float64)
        nacc, nhacc = ctypes.c_int(), ctypes.c_int()
        hc.hc_hmm_acc_sizes(ctypes.byref(nacc), ctypes.byref(nhacc))
        acc = torch.zeros(shard.C * nacc.value + shard.C * nhacc.value, dtype=torch.float64)     # one buffer, one all-reduce
        hacc = acc[shard.C * nacc.value:]
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        lgrads, ggrads = torch.empty_like(lparams), torch.zeros(gl.numel, dtype=torch.float64)
        pix, xy = shard.images.contiguous(), shard.xy.contiguous()
        ont, mask = shard.is_ontarget.to(torch.uint8).contiguous(), shard.mask.to(torch.uint8).contiguous()
        off_s, off_w = shard.offset_samples.contiguous(), shard.offset_logits.contiguous()
        sN = data.Nt / len(ndx)
        hc.hc_hmm_step_acc_f64.restype = ctypes.c_double
        hc.hc_hmm_step_acc_f64(len(sndx), shard.Nt, shard.F, shard.C, shard.P, off_s.numel(), p(sndx), p(pix), p(xy), p(ont), p(mask),
                               p(off_s), p(off_w), ctypes.byref(mc), ctypes.c_double(sN), p(lparams), p(gparams), p(lnoise), p(gnoise),
                               p(lgrads), p(ggrads), p(acc), p(hacc))
        dist.all_reduce(acc)
        ggrads.zero_()
        hc.hc_hmm_globals_post.restype = ctypes.c_double
        loss = hc.hc_hmm_globals_post(shard.C, ctypes.byref(mc), p(gparams), p(gnoise), p(acc), p(hacc), ctypes.c_double(sN), p(ggrads))
        grads = {k: v.numpy().copy() for k, v in ll.named(lgrads).items()}
        grads.update({k: v.numpy().copy() for k, v in gl.views(ggrads).items()})
        out_q.put((rank, loss, grads, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [dict(N=4, F=5, C=1, seed=0, ndx=[3, 0, 2]), dict(N=4, F=4, C=2, seed=1, ndx=[1, 2, 3, 0])])
def test_two_rank_sharded_hmm_step_equals_single_process_oracle(cfg):
    from oracle import hmm_oracle as H
    from tests.test_hmm_cpu import small_problem

    world = 2
    ds, data, params, _, _ = small_problem(N=cfg["N"], F=cfg["F"], C=cfg["C"], seed=cfg["seed"])
    ndx = torch.tensor(cfg["ndx"])
    noise = H.draw_noise(params, data, ndx, torch.Generator().manual_seed(cfg["seed"] + 9))
    ref_loss, ref_grads = H.loss_and_grads(params, data, ndx, noise)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 300
    procs = [ctx.Process(target=_hmm_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    aoi_axis = {"m_probs": 2, "z_trans": 0}
    for rank, loss, grads, lo, hi in results:
        assert abs(loss - ref_loss) <= 1e-11 * abs(ref_loss)
        for k, ref in ref_grads.items():
            g = torch.from_numpy(grads[k])
            if k not in H.GLOBAL_PARAMS:
                ref = ref.narrow(aoi_axis.get(k, 1 if ref.dim() == 4 else 0), lo, hi - lo)
            err = (g.reshape(ref.shape) - ref).abs().max().item()
            assert err <= 1e-8 * max(ref_grads[k].abs().max().item(), 1e-30), (k, err)
