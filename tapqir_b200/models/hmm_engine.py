"""
One SVI step of the *hmm* variant (reference: tapqir/models/hmm.py) on the kernels of csrc/: the continuous guide
sites and the 4-configuration likelihood kernel are cosmos' own; around them the guide's Markov chain
(tq_hmm_forward / tq_hmm_backward), the per-state emission terms (tq_hmm_local_post) and the init / trans global
sites (tq_hmm_globals_*).  See include/tapqir_b200.h for the call order and csrc/cosmos_hmm.cuh for the maths.
"""

import ctypes
from collections import OrderedDict

import torch

from tapqir_b200 import _lib
from tapqir_b200.models import layout as L
from tapqir_b200.models.engine import CosmosEngine


class HmmEngine(CosmosEngine):
    def __init__(self, store, Nt_local, F, C, P, priors, **kw):
        kw["fbatch_size"] = F   # every frame, every step (hmm.py:127-131)
        super().__init__(store, Nt_local, F, C, P, priors, **kw)
        dev, dtype, f64 = self.device, self.dtype, torch.float64
        self.ll = L.HmmLocalLayout(self.Nt, self.F, self.C)
        assert self.ll.numel == self.lib.tq_hmm_local_numel(self.Nt, self.F, self.C)
        self.gl = L.HmmGlobalLayout(self.C)
        z = lambda n, dt=dtype: torch.zeros(n, dtype=dt, device=dev)
        self.lparams, self.lgrads, self.lm, self.lv = z(self.ll.numel), z(self.ll.numel), z(self.ll.numel), z(self.ll.numel)
        self.gparams, self.ggrads, self.gm, self.gv = (z(self.gl.numel, f64) for _ in range(4))   # always float64 (engine.py)
        self.nh = self.lib.tq_hmm_chain_sums()
        # accumulators and chain sums in one buffer: one cross-rank sum covers both
        self.acc_all = z(self.C * (L.NACC + self.nh), f64)
        self.acc, self.hacc = self.acc_all[:self.C * L.NACC], self.acc_all[self.C * L.NACC:]
        self.set_batch(kw.get("nbatch_size") or self.Nt, F)

    def set_batch(self, nbatch_size, fbatch_size):
        super().set_batch(nbatch_size, self.F)
        if not hasattr(self, "nh"):
            return
        dev = self.device
        self.chain_a = torch.empty(2, self.U, dtype=torch.float64, device=dev)
        self.chain_rows = torch.empty(self.lib.tq_hmm_chain_rows(), self.U, dtype=torch.float64, device=dev)
        self.chain_v = torch.empty(2, self.U, dtype=self.dtype, device=dev)
        self.hpartial = torch.empty(max(self.nb * self.C, 1) * self.nh, dtype=torch.float64, device=dev)

    # ---- parameter access (reference names / shapes) ------------------------------------------------------------------
    def named_unconstrained(self):
        out = OrderedDict(self.ll.named(self.lparams))
        out.update(self.gl.views(self.gparams))
        return out

    def named_grads(self):
        out = OrderedDict(self.ll.named(self.lgrads))
        out.update(self.gl.views(self.ggrads))
        return out

    def load_unconstrained(self, tensors):
        self.ll.load_named(self.lparams, tensors)
        for k, v in self.gl.views(self.gparams).items():
            v.copy_(tensors[k].to(device=self.device, dtype=v.dtype).reshape(v.shape))

    # ---- one step ------------------------------------------------------------------------------------------------------
    def _enqueue(self, ndx=None, fdx=None, local_noise=None, global_noise=None, update=True, time_likelihood=None):
        assert fdx is None, "the hmm variant uses every frame"
        lib, st, code = self.lib, _lib.stream_ptr(self.device), self.code
        p = _lib.ptr
        mc = ctypes.byref(self.mc)
        with torch.cuda.device(self.device):
            if ndx is None and not self.full_n:
                _lib.check(lib.tq_subsample(self.Nt, self.nb, self.seed, p(self.state), 2 + self.rank, p(self.perm_n),
                                            p(self.ndx), st), "tq_subsample")
                ndx = self.ndx
            view = self._view(ndx, None)
            if not self.full_n:
                self.lgrads.zero_()
            main = torch.cuda.current_stream(self.device)
            self._ev_fork0.record(main)
            self._side.wait_event(self._ev_fork0)
            with torch.cuda.stream(self._side):
                sst = _lib.stream_ptr(self.device)
                _lib.check(lib.tq_hmm_globals_sample(code, self.C, p(self.gparams), mc, p(global_noise), self.seed,
                                                     p(self.state), p(self.gstate), p(self.tables), p(self.gain), sst),
                           "tq_hmm_globals_sample")
                self._ev_join0.record(self._side)
                _lib.check(lib.tq_hmm_globals_prepare(code, self.C, p(self.gparams), mc, p(self.gstate), p(self.gprep), sst),
                           "tq_hmm_globals_prepare")
            # sites that leave the fp32 forms are collected in a worklist (the not-yet-written gradient buffer of the
            # likelihood kernel serves as its storage) and redone in double by dense warps
            _lib.check(lib.tq_cosmos_sites_ws(code, view, self.Nt, mc, p(self.lparams), self.aoi_offset, self.seed,
                                              p(self.state), p(local_noise), p(self.samples), p(self.qm), p(self.rec),
                                              p(self.gs), p(self.work_count), st), "tq_cosmos_sites_ws")
            main.wait_event(self._ev_join0)   # the chain's per-frame terms need the sampled init / trans tables
            _lib.check(lib.tq_hmm_forward(code, view, self.Nt, mc, p(self.lparams), p(self.tables), p(self.chain_rows),
                                          p(self.chain_a), p(self.qm), st), "tq_hmm_forward")
            S, G, K = self.samples, self.gs, L.K
            if time_likelihood is not None:
                time_likelihood[0].record()
            _lib.check(lib.tq_ksmogn_fwd_bwd(code, view, p(S[1:1 + K]), p(S[1 + K:1 + 2 * K]), p(S[1 + 2 * K:1 + 3 * K]),
                                             p(S[1 + 3 * K:1 + 4 * K]), p(S[0]), p(self.gain), p(self.mcfg_arg), 4, p(self.qm),
                                             p(self.Lm), p(G[1:1 + K]), p(G[1 + K:1 + 2 * K]), p(G[1 + 2 * K:1 + 3 * K]),
                                             p(G[1 + 3 * K:1 + 4 * K]), p(G[0]), p(self.g_rate), st), "tq_ksmogn_fwd_bwd")
            if time_likelihood is not None:
                time_likelihood[1].record()
            _lib.check(lib.tq_hmm_local_post(code, view, self.Nt, mc, p(self.lparams), p(self.tables), p(self.samples),
                                             p(self.rec), p(self.Lm), p(self.gs), p(self.g_rate), p(self.chain_a), self.sN,
                                             p(self.lgrads), p(self.chain_v), p(self.tickets), p(self.block_partial), p(self.acc),
                                             st), "tq_hmm_local_post")
            _lib.check(lib.tq_hmm_backward(code, view, self.Nt, mc, p(self.lparams), p(self.tables), p(self.chain_rows),
                                           p(self.chain_a), p(self.chain_v), self.sN, p(self.lgrads), p(self.hpartial),
                                           p(self.hacc), st),
                       "tq_hmm_backward")
            if abs(self.acc_weight - 1.0) > 1e-15:
                self.acc_all.mul_(self.acc_weight)  # unequal AOI shards (CosmosEngine.set_batch)
            if self.p2p is not None:
                self.p2p.push(self.acc_all, st)
            self._ev_fork.record(main)
            self._side.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side):
                sst = _lib.stream_ptr(self.device)
                if self.p2p is not None:
                    self.p2p.wait_sum(self.acc_all, sst)
                elif self.world_size > 1:
                    torch.distributed.all_reduce(self.acc_all, group=self.pg)
                _lib.check(lib.tq_hmm_globals_finish(code, self.C, mc, p(self.gstate), p(self.gprep), p(self.acc), p(self.hacc),
                                                     self.sN_ref, p(self.ggrads), p(self.loss), sst), "tq_hmm_globals_finish")
                if update:
                    b1, b2 = self.betas
                    _lib.check(lib.tq_adam_dense(_lib.TQ_F64, self.gl.numel, p(self.gparams), p(self.ggrads), p(self.gm),
                                                 p(self.gv), self.lr, b1, b2, self.adam_eps, p(self.state), sst),
                               "tq_adam_dense")
                self._ev_join.record(self._side)
            if update:
                b1, b2 = self.betas
                _lib.check(lib.tq_adam_dense(code, self.ll.numel, p(self.lparams), p(self.lgrads), p(self.lm), p(self.lv),
                                             self.lr, b1, b2, self.adam_eps, p(self.state), st), "tq_adam_dense")
            main.wait_event(self._ev_join)
            if update:
                _lib.check(lib.tq_step_advance(p(self.state), st), "tq_step_advance")
        return self.loss

    # ---- posterior of the chain (hmm.py:627-633) --------------------------------------------------------------------------
    @torch.no_grad()
    def z_probs(self):
        """Forward marginals a_f(z) of the guide's chain for every local AOI: (Nt, F, C, 2)."""
        lib, p = self.lib, _lib.ptr
        view = _lib.make_view(self.store.pixels, self.store.xy, self.store.offset_samples, self.store.offset_logits, nb=self.Nt,
                              fb=self.F, C=self.C, F=self.F, P=self.P, ndx=None, fdx=None, is_ontarget=self.store.is_ontarget,
                              mask=self.store.mask)
        U = self.Nt * self.F * self.C
        a = torch.empty(2, U, dtype=torch.float64, device=self.device)
        rows = torch.empty(lib.tq_hmm_chain_rows(), U, dtype=torch.float64, device=self.device)
        qm = torch.empty(4, U, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            # (the forward marginals use the guide's rows only; the tables enter the other per-frame terms)
            _lib.check(lib.tq_hmm_forward(self.code, view, self.Nt, ctypes.byref(self.mc), p(self.lparams), p(self.tables),
                                          p(rows), p(a), p(qm), _lib.stream_ptr(self.device)), "tq_hmm_forward")
        return a.view(2, self.Nt, self.F, self.C).permute(1, 2, 3, 0).contiguous()

    @torch.no_grad()
    def compute_theta_probs(self, z_map, aoi_count=None, particles=5, local_noise=None, global_noise=None, seed_offset=1 << 40):
        """
        theta_probs (K, n, F, C) for local AOIs [0, aoi_count): ``q(theta = k | z = z_MAP)`` averaged over ``particles``
        guide draws (reference: hmm.py:541-625, 5 particles).  ``z_map``: (n, F, C) integer / bool CUDA tensor.
        ``local_noise`` / ``global_noise`` (lists, one per particle) replay given draws for the parity test.
        """
        lib, code, p = self.lib, self.code, _lib.ptr
        mc = ctypes.byref(self.mc)
        dev, dtype = self.device, self.dtype
        n = self.Nt if aoi_count is None else int(aoi_count)
        ndx = torch.arange(n, dtype=torch.int32, device=dev)
        U = n * self.F * self.C
        s = self.store
        view = _lib.make_view(s.pixels, s.xy, s.offset_samples, s.offset_logits, nb=n, fb=self.F, C=self.C, F=self.F, P=self.P,
                              ndx=ndx, fdx=None, is_ontarget=s.is_ontarget, mask=s.mask)
        samples = torch.empty(L.NSAMP, U, dtype=dtype, device=dev)
        qm = torch.empty(4, U, dtype=dtype, device=dev)
        rec = torch.empty(lib.tq_site_record_rows(), U, dtype=dtype, device=dev)
        theta_probs = torch.zeros(L.K, n, self.F, self.C, dtype=dtype, device=dev)
        zm = z_map.to(device=dev, dtype=torch.uint8).contiguous()
        assert zm.numel() == U
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            for i in range(particles):
                pstate = torch.tensor([seed_offset + i], dtype=torch.int64, device=dev)
                gn = None if global_noise is None else global_noise[i]
                ln = None if local_noise is None else local_noise[i]
                _lib.check(lib.tq_hmm_globals_sample(code, self.C, p(self.gparams), mc, p(gn), self.seed, p(pstate),
                                                     p(self.gstate), p(self.tables), p(self.gain), st), "tq_hmm_globals_sample")
                _lib.check(lib.tq_cosmos_sites(code, view, self.Nt, mc, p(self.lparams), self.aoi_offset, self.seed,
                                               p(pstate), p(ln), p(samples), p(qm), p(rec), st), "tq_cosmos_sites")
                _lib.check(lib.tq_hmm_theta_probs(code, view, self.Nt, mc, p(self.lparams), p(self.tables), p(samples), p(zm),
                                                  1.0 / particles, p(theta_probs), st), "tq_hmm_theta_probs")
        return theta_probs

    def compute_probs(self, *args, **kwargs):
        raise NotImplementedError("use z_probs() and compute_theta_probs() for the hmm variant")
