"""
Step engine: owns the device buffers of one rank's AOI shard and enqueues the kernels of one
cosmos SVI step through the C ABI (include/tapqir_b200.h).

One step (what ``svi.step()`` does in models/model.py:212):

    subsample -> globals_sample -> sites -> ksmogn_fwd_bwd -> local_post (+ reductions)
    -> [all-reduce of the (C, NACC) accumulators across ranks] -> globals_grad -> Adam (local, global)

Everything runs on torch's current stream; no host synchronisation inside a step (the loss stays on
the device until someone reads it).  There is no CPU path: constructing an engine without the
sm_100a library or without a CUDA device raises.
"""

import ctypes
import os
from collections import OrderedDict

import torch

from tapqir_b200 import _lib
from tapqir_b200.models import layout as L


class CosmosEngine:
    def __init__(self, store, Nt_local, F, C, P, priors, dtype=torch.float32, lr=0.005, betas=(0.9, 0.999),
                 adam_eps=1e-8, nbatch_size=None, fbatch_size=None, seed=0, ref_dtype=torch.float64,
                 Nt_total=None, aoi_offset=0, rank=0, world_size=1, process_group=None, use_graph=True,
                 shard_sizes=None):
        self.lib = _lib.load()
        self.store = store
        self.device = store.pixels.device
        if self.device.type != "cuda":
            raise ValueError("CosmosEngine needs a CUDA device store (no CPU path)")
        self.dtype = dtype
        self.code = _lib.dtype_code(dtype)
        self.Nt, self.F, self.C, self.P = int(Nt_local), int(F), int(C), int(P)
        self.Nt_total = int(Nt_total if Nt_total is not None else Nt_local)
        self.aoi_offset, self.rank, self.world_size, self.pg = int(aoi_offset), int(rank), int(world_size), process_group
        # AOIs held by every rank (None: equal shards).  Unequal shards are fine: see set_batch
        self.shard_sizes = [int(n) for n in shard_sizes] if shard_sizes is not None else [self.Nt] * self.world_size
        if len(self.shard_sizes) != self.world_size or self.shard_sizes[self.rank] != self.Nt or min(self.shard_sizes) < 1:
            raise ValueError(f"shard sizes {self.shard_sizes} do not describe {self.world_size} non-empty shards with "
                             f"{self.Nt} AOIs on rank {self.rank}")
        self.lr, self.betas, self.adam_eps = float(lr), (float(betas[0]), float(betas[1])), float(adam_eps)
        self.seed = int(seed)
        self.mc = L.ModelConst.make(priors, P, ref_dtype)
        assert self.lib.tq_sizeof_model_const() == ctypes.sizeof(self.mc), "ModelConst ABI mismatch"
        self.ll = L.LocalLayout(self.Nt, self.F, self.C)
        self.gl = L.GlobalLayout(self.C)
        dev, f64 = self.device, torch.float64
        z = lambda n, dt=dtype: torch.zeros(n, dtype=dt, device=dev)
        # parameters, gradients, Adam moments (flat; see layout.py)
        # (lparams / lm / lv are properties: reading them first applies a deferred update, see flush_deferred)
        self._pending = False
        self.lparams, self.lgrads, self.lm, self.lv = z(self.ll.numel), z(self.ll.numel), z(self.ll.numel), z(self.ll.numel)
        # the global parameters (4 + 5Q scalars), their gradients and moments are float64 whatever `dtype` is: the
        # global sites are evaluated in double anyway, and rounding gain_loc / gain_beta to fp32 alone moves the
        # gradient of gain_beta by 1e-4 of itself (it cancels ~1e3-fold in the base variate)
        self.gparams, self.ggrads, self.gm, self.gv = (z(self.gl.numel, f64) for _ in range(4))
        # small device-resident state
        # StepState {uint64 step; uint32 pending; float step_size, inv_sqrt_bc2}: [0] is the iteration counter
        assert self.lib.tq_sizeof_step_state() == 24, "StepState ABI mismatch"
        self.state = torch.zeros(3, dtype=torch.int64, device=dev)
        self.tables = torch.zeros(self.lib.tq_sizeof_tables(), dtype=torch.uint8, device=dev)
        self.gstate = torch.zeros(self.lib.tq_sizeof_gstate() // 8, dtype=f64, device=dev)
        self.gain = z(1)
        self.acc = torch.zeros(self.C * L.NACC, dtype=f64, device=dev)
        self.loss = torch.zeros(1, dtype=f64, device=dev)
        self.work_count = torch.zeros(2, dtype=torch.int32, device=dev)   # entries of the double-fallback worklist
        self.gprep = torch.zeros(self.lib.tq_sizeof_gprep() // 8, dtype=f64, device=dev)   # prepared global reverse mode
        self.mcfg = torch.tensor([[(m >> k) & 1 for k in range(L.K)] for m in range(2**L.K)], dtype=dtype, device=dev)
        self.acc_all = self.acc   # everything that is summed across ranks, contiguous (the hmm engine appends its chain sums)
        # cross-rank sum of the accumulators: NVLink peer-memory push (csrc/p2p_allreduce.cu) unless TQ_ALLREDUCE=nccl or the
        # IPC set-up is not possible (then NCCL; 120-160 us per step at 8 GPUs against ~10 us)
        self.p2p = None
        if self.world_size > 1 and os.environ.get("TQ_ALLREDUCE", "p2p") == "p2p":
            from tapqir_b200.models.p2p import P2PAllReduce, PeerMemoryUnavailable

            try:
                # collective: either EVERY rank gets the peer-memory path or every rank raises (the ranks agree on the
                # outcome with an all-reduce inside), so the fallback below is taken by all ranks together
                self.p2p = P2PAllReduce(dev, self.rank, self.world_size, self.pg)
            except PeerMemoryUnavailable as err:   # no peer access / IPC (e.g. containers without it): NCCL does the same sum
                import logging

                logging.getLogger(__name__).warning(f"peer-memory all-reduce unavailable ({err}); using NCCL")
                self.p2p = None
        self.use_graph = use_graph
        # TQ_FUSED=1: sites -> likelihood -> post as ONE persistent kernel (csrc/cosmos_fused.cu; dtype float, P = 14,
        # uint16 pixels).  Parity-green but NOT the default: measured on a B200 it is slower than the three per-stage
        # kernels (C3: 9.7 vs 6.8 ms per step, C2: 0.267 vs 0.185 ms; profiles/r2_fused_ab.md) -- one register budget
        # (128) for three very different phases leaves 16 warps per SM, and the site / post phases, which the
        # stand-alone kernels run at 32 / 12 warps per SM, are latency-bound at 4 warps per block
        self.fused = os.environ.get("TQ_FUSED", "0") == "1" and type(self).__name__ == "CosmosEngine"
        # Deferred local Adam (default; TQ_DEFERRED_ADAM=0 turns it off): on full-batch float32 steps the dense Adam
        # update of the AOI-local parameters -- 28 B of HBM traffic per element, 0.39 of 6.1 ms at 1000 AOIs x 5000
        # frames -- is applied by the NEXT step's site kernel, which is issue-bound and leaves the memory system idle
        # (tq_cosmos_sites_adam): each site's thread updates the parameters it owns just before it reads them.  Bit for
        # bit the same parameters and moments as the separate kernel; anything else that reads them (checkpoints,
        # statistics, a subsampled or gradient-only step) goes through flush_deferred first.
        self.deferred_adam = (os.environ.get("TQ_DEFERRED_ADAM", "1") == "1" and type(self).__name__ == "CosmosEngine"
                              and dtype == torch.float32)
        # measured (B200, profiles/r2_deferred_ab.sh): 5 M units 6.09 -> 5.97 ms, 2 M (two channels) 2.52 -> 2.47 ms,
        # 625 k and 100 k units: no gain (the dense update is a few us there and the split costs one more launch)
        self.deferred_min_units = int(os.environ.get("TQ_DEFERRED_ADAM_MIN_UNITS", 1 << 20))
        begin, end = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self.lib.tq_local_deferred_range(self.Nt, self.F, self.C, ctypes.byref(begin), ctypes.byref(end)),
                   "tq_local_deferred_range")
        self._deferred_range = (int(begin.value), int(end.value))
        self.keep_intermediates = False   # fused path: also write samples / L back to HBM (tests, diagnostics)
        self._side = torch.cuda.Stream(device=self.device)
        self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork0, self._ev_join0 = torch.cuda.Event(), torch.cuda.Event()
        self._graph, self._eager_default_steps = None, 0
        self.mcfg_arg = None  # NULL = built-in enumerated table -> fp32 production kernel; set to self.mcfg for the generic one
        self.set_batch(nbatch_size or self.Nt, fbatch_size or self.F)

    # ---- deferred local Adam ------------------------------------------------------------------------------
    def flush_deferred(self):
        """Apply a deferred AOI-local Adam update now (no-op when none is pending)."""
        if not self._pending:
            return
        self._pending = False
        lo, hi = self._deferred_range
        b1, b2 = self.betas
        p = _lib.ptr
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tq_adam_deferred_flush(self.code, hi - lo, p(self._lparams[lo:hi]), p(self.lgrads[lo:hi]),
                                                       p(self._lm[lo:hi]), p(self._lv[lo:hi]), b1, b2, self.adam_eps,
                                                       p(self.state), _lib.stream_ptr(self.device)), "tq_adam_deferred_flush")

    def _flat_property(name):
        def get(self):
            self.flush_deferred()
            return getattr(self, name)

        def set_(self, value):
            self.flush_deferred()
            setattr(self, name, value)

        return property(get, set_)

    lparams, lm, lv = _flat_property("_lparams"), _flat_property("_lm"), _flat_property("_lv")
    del _flat_property

    # ---- buffers that depend on the minibatch shape ---------------------------------------------------
    def set_batch(self, nbatch_size, fbatch_size):
        self.flush_deferred()
        self._graph, self._eager_default_steps = None, 0  # buffers below are re-allocated: drop any captured graph
        self.nb, self.fb = min(int(nbatch_size), self.Nt), min(int(fbatch_size), self.F)
        dev, dtype = self.device, self.dtype
        U = self.U = self.nb * self.fb * self.C
        self.full_n, self.full_f = self.nb == self.Nt, self.fb == self.F
        self.ndx = None if self.full_n else torch.zeros(self.nb, dtype=torch.int32, device=dev)
        self.fdx = None if self.full_f else torch.zeros(self.fb, dtype=torch.int32, device=dev)
        self.perm_n = torch.arange(self.Nt, dtype=torch.int32, device=dev)
        self.perm_f = torch.arange(self.F, dtype=torch.int32, device=dev)
        e = lambda *s, dt=dtype: torch.empty(*s, dtype=dt, device=dev)
        self.samples, self.gs = e(L.NSAMP, U), e(L.NSAMP, U)
        self.qm, self.Lm, self.g_rate = e(4, U), e(4, U), e(U)
        self.rec = e(self.lib.tq_site_record_rows(), U)   # per-site records (csrc/cosmos_local.cuh SO_*/EX_*)
        # scratch of tq_cosmos_local_post: per-block partial sums, and its (self-resetting) completion tickets
        self.block_partial = e(max(self.lib.tq_local_post_scratch(self.nb, self.fb, self.C),
                                   self.lib.tq_cosmos_fused_scratch(self.nb, self.fb, self.C), 1), dt=torch.float64)
        self.tickets = torch.zeros(max(self.lib.tq_local_post_tickets(self.nb, self.fb, self.C),
                                       self.lib.tq_cosmos_fused_tickets(self.nb, self.fb, self.C), 1), dtype=torch.float64, device=dev)
        # scale factors of the subsampled plates (cosmos.py:194-208).  The AOI minibatch is stratified by shard: rank r
        # draws nb_r = min(nbatch_size, Nt_r) of its Nt_r AOIs, each with inclusion probability nb_r / Nt_r, so its terms
        # carry sN = Nt_r / nb_r.  The cross-rank sum and the global reverse mode use ONE reference scale
        # sN_ref = Nt_total / sum_r nb_r; a rank whose own scale differs (unequal shards) weights its accumulators by
        # sN / sN_ref before the sum.  Equal shards: sN == sN_ref, weight 1.
        nb_all = [min(int(nbatch_size), n) for n in self.shard_sizes]
        self.sN = self.Nt / self.nb
        self.sN_ref = self.Nt_total / sum(nb_all)
        self.acc_weight = self.sN / self.sN_ref
        self.sF = self.F / self.fb

    def _view(self, ndx, fdx):
        s = self.store
        return _lib.make_view(s.pixels, s.xy, s.offset_samples, s.offset_logits, nb=self.nb, fb=self.fb, C=self.C,
                              F=self.F, P=self.P, ndx=ndx, fdx=fdx, is_ontarget=s.is_ontarget, mask=s.mask)

    # ---- parameter access -------------------------------------------------------------------------------
    def named_unconstrained(self):
        out = OrderedDict(self.ll.views(self.lparams))
        out.update(self.gl.views(self.gparams))
        return out

    def named_grads(self):
        out = OrderedDict(self.ll.views(self.lgrads))
        out.update(self.gl.views(self.ggrads))
        return out

    def load_unconstrained(self, tensors):
        views = self.named_unconstrained()
        for k, v in views.items():
            v.copy_(tensors[k].to(device=self.device, dtype=v.dtype).reshape(v.shape))

    # ---- one step ------------------------------------------------------------------------------------------
    def step(self, ndx=None, fdx=None, local_noise=None, global_noise=None, update=True, time_likelihood=None):
        """
        One SVI step.  The default call (device-drawn minibatch and variates, parameter update) is
        captured into a CUDA graph the second time it runs and replayed from then on: every
        step-varying quantity (iteration counter for Philox and Adam, minibatch indices) lives in
        device memory, so the captured launches never change.  Any explicit argument (replay-mode
        tests, kernel timing) takes the eager path :meth:`_enqueue`.
        """
        default = (ndx is None and fdx is None and local_noise is None and global_noise is None and update
                   and time_likelihood is None)
        if not (default and self.use_graph):
            return self._enqueue(ndx, fdx, local_noise, global_noise, update, time_likelihood)
        if self._graph is not None:
            self._graph.replay()
            self._pending = self._graph_defers
            return self.loss
        self._eager_default_steps += 1
        if self._eager_default_steps < 2:
            return self._enqueue(None, None, None, None, True, None)  # first call: lazy CUDA/NCCL init outside capture
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._enqueue(None, None, None, None, True, None)
        self._graph, self._graph_defers = graph, self._pending
        graph.replay()
        return self.loss

    def _enqueue(self, ndx=None, fdx=None, local_noise=None, global_noise=None, update=True, time_likelihood=None):
        """
        Enqueue one SVI step.  ``ndx``/``fdx`` (int32 CUDA tensors of local AOI / frame indices) and
        the base variates (``local_noise`` (NSAMP, U) in ``dtype``; ``global_noise`` float64 in
        GlobalLayout noise order) put the step in *replay* mode for parity tests; by default indices
        and variates are drawn on the device with Philox keyed by (seed, step).
        With ``update=False`` parameters are left untouched (gradients only).
        ``time_likelihood=(start_event, end_event)`` records CUDA events around the likelihood kernel.
        Returns the device tensor holding the loss (-ELBO).
        """
        lib, st, code = self.lib, _lib.stream_ptr(self.device), self.code
        p = _lib.ptr
        mc = ctypes.byref(self.mc)
        # this step leaves its local Adam update to the next site kernel / applies the previous one in its own
        defer = (self.deferred_adam and update and ndx is None and fdx is None and self.full_n and self.full_f
                 and self.U >= max(self.deferred_min_units, 1) and not self.fused)
        if not defer:
            self.flush_deferred()
        with torch.cuda.device(self.device):
            if (ndx is None and not self.full_n and fdx is None and not self.full_f
                    and lib.tq_subsample_pair_supported(self.Nt, self.F)):
                # both draws in one launch (the same frame subset on every rank: its stream id does not depend on the rank)
                _lib.check(lib.tq_subsample_pair(self.Nt, self.nb, 2 + self.rank, p(self.ndx), self.F, self.fb, 1, p(self.fdx),
                                                 self.seed, p(self.state), st), "tq_subsample_pair")
                ndx, fdx = self.ndx, self.fdx
            if ndx is None and not self.full_n:
                _lib.check(lib.tq_subsample(self.Nt, self.nb, self.seed, p(self.state), 2 + self.rank, p(self.perm_n),
                                            p(self.ndx), st), "tq_subsample")
                ndx = self.ndx
            if fdx is None and not self.full_f:
                # same frame subset on every rank: stream id does not depend on the rank
                _lib.check(lib.tq_subsample(self.F, self.fb, self.seed, p(self.state), 1, p(self.perm_f), p(self.fdx), st),
                           "tq_subsample")
                fdx = self.fdx
            view = self._view(ndx, fdx)
            if not (self.full_n and self.full_f):
                self.lgrads.zero_()  # dense zero-filled gradient outside the minibatch
            # sampling the globals (one warp, latency-bound) and evaluating the guide sites (independent
            # of the globals) run concurrently; the likelihood kernel needs both
            main = torch.cuda.current_stream(self.device)
            self._ev_fork0.record(main)
            self._side.wait_event(self._ev_fork0)
            with torch.cuda.stream(self._side):
                _lib.check(lib.tq_cosmos_globals_sample(code, self.C, p(self.gparams), mc, p(global_noise), self.seed,
                                                        p(self.state), p(self.gstate), p(self.tables), p(self.gain),
                                                        _lib.stream_ptr(self.device)), "tq_cosmos_globals_sample")
                self._ev_join0.record(self._side)
                # acc-independent part of the globals' reverse mode: off the critical path, under the likelihood kernel
                _lib.check(lib.tq_cosmos_globals_prepare(code, self.C, p(self.gparams), mc, p(self.gstate), p(self.gprep),
                                                         _lib.stream_ptr(self.device)), "tq_cosmos_globals_prepare")
            fused = (self.fused and self.mcfg_arg is None and self.U > 0
                     and bool(lib.tq_cosmos_fused_supported(code, ctypes.byref(view))))
            self.last_step_fused = fused
            if fused:
                main.wait_event(self._ev_join0)   # the sampled gain and the prior tables
                if time_likelihood is not None:
                    time_likelihood[0].record()
                keep = self.keep_intermediates
                _lib.check(lib.tq_cosmos_fused_step(code, view, self.Nt, mc, p(self.lparams), p(self.tables), p(self.gain),
                                                    self.aoi_offset, self.seed, p(self.state), p(local_noise), self.sN, self.sF,
                                                    p(self.lgrads), p(self.tickets), p(self.block_partial), p(self.acc),
                                                    p(self.samples) if keep else None, p(self.Lm) if keep else None, st),
                           "tq_cosmos_fused_step")
                if time_likelihood is not None:
                    time_likelihood[1].record()
            else:
                self._enqueue_stages(view, local_noise, main, time_likelihood, st, defer)
            # the all-reduce of the (C, 18) accumulators (multi-GPU), finishing the global reverse pass (a few FMAs per
            # parameter) and the global Adam run on the side stream, beside the dense Adam over the AOI-local buffer,
            # which depends on none of them: the collective's latency hides under the local update
            if abs(self.acc_weight - 1.0) > 1e-15:
                self.acc_all.mul_(self.acc_weight)  # unequal AOI shards (set_batch)
            if self.p2p is not None:
                self.p2p.push(self.acc_all, st)     # straight into every peer's buffer, as early as the values exist
            self._ev_fork.record(main)
            self._side.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side):
                sst = _lib.stream_ptr(self.device)
                if self.p2p is not None:
                    self.p2p.wait_sum(self.acc_all, sst)
                elif self.world_size > 1:
                    torch.distributed.all_reduce(self.acc_all, group=self.pg)
                _lib.check(lib.tq_cosmos_globals_finish(code, self.C, mc, p(self.gstate), p(self.gprep), p(self.acc),
                                                        self.sN_ref, self.sF, p(self.ggrads), p(self.loss), sst),
                           "tq_cosmos_globals_finish")
                if update:
                    b1, b2 = self.betas
                    _lib.check(lib.tq_adam_dense(_lib.TQ_F64, self.gl.numel, p(self.gparams), p(self.ggrads), p(self.gm),
                                                 p(self.gv), self.lr, b1, b2, self.adam_eps, p(self.state), sst),
                               "tq_adam_dense")
                self._ev_join.record(self._side)
            if update:
                b1, b2 = self.betas
                lo, hi = self._deferred_range if defer else (0, 0)
                for a, b in ((0, lo), (hi, self.ll.numel)):   # everything, or what the site kernel does not own
                    if b > a:
                        _lib.check(lib.tq_adam_dense(code, b - a, p(self._lparams[a:b]), p(self.lgrads[a:b]), p(self._lm[a:b]),
                                                     p(self._lv[a:b]), self.lr, b1, b2, self.adam_eps, p(self.state), st),
                                   "tq_adam_dense")
            main.wait_event(self._ev_join)
            if update and defer:
                _lib.check(lib.tq_step_advance_deferred(p(self.state), self.lr, b1, b2, st), "tq_step_advance_deferred")
                self._pending = True
            elif update:
                _lib.check(lib.tq_step_advance(p(self.state), st), "tq_step_advance")
        return self.loss

    def _enqueue_stages(self, view, local_noise, main, time_likelihood, st, defer=False):
        """The AOI-local part of a step as three kernels with HBM scratch between them (dtype double, float pixels,
        P != 14, the operator-level mcfg table, TQ_FUSED=0)."""
        lib, code, p = self.lib, self.code, _lib.ptr
        mc = ctypes.byref(self.mc)
        # sites that leave the fp32 forms are collected in a worklist (the not-yet-written gradient buffer of the
        # likelihood kernel serves as its storage) and redone in double by dense warps
        if defer:
            b1, b2 = self.betas
            _lib.check(lib.tq_cosmos_sites_adam(code, view, self.Nt, mc, p(self._lparams), self.aoi_offset, self.seed,
                                                p(self.state), p(local_noise), p(self.samples), p(self.qm), p(self.rec),
                                                p(self.gs), p(self.work_count), p(self.lgrads), p(self._lm), p(self._lv),
                                                b1, b2, self.adam_eps, st), "tq_cosmos_sites_adam")
        else:
            _lib.check(lib.tq_cosmos_sites_ws(code, view, self.Nt, mc, p(self._lparams), self.aoi_offset, self.seed,
                                              p(self.state), p(local_noise), p(self.samples), p(self.qm), p(self.rec),
                                              p(self.gs), p(self.work_count), st), "tq_cosmos_sites_ws")
        main.wait_event(self._ev_join0)
        S, G, K = self.samples, self.gs, L.K
        if time_likelihood is not None:
            time_likelihood[0].record()
        _lib.check(lib.tq_ksmogn_fwd_bwd(code, view, p(S[1:1 + K]), p(S[1 + K:1 + 2 * K]), p(S[1 + 2 * K:1 + 3 * K]),
                                         p(S[1 + 3 * K:1 + 4 * K]), p(S[0]), p(self.gain), p(self.mcfg_arg), 4, p(self.qm),
                                         p(self.Lm), p(G[1:1 + K]), p(G[1 + K:1 + 2 * K]), p(G[1 + 2 * K:1 + 3 * K]),
                                         p(G[1 + 3 * K:1 + 4 * K]), p(G[0]), p(self.g_rate), st), "tq_ksmogn_fwd_bwd")
        if time_likelihood is not None:
            time_likelihood[1].record()
        _lib.check(lib.tq_cosmos_local_post(code, view, self.Nt, mc, p(self._lparams), p(self.tables), p(self.samples),
                                            p(self.rec), p(self.Lm), p(self.gs), p(self.g_rate), self.sN, self.sF, p(self.lgrads),
                                            p(self.tickets), p(self.block_partial), p(self.acc), st),
                   "tq_cosmos_local_post")

    # ---- posterior of the enumerated latents (cosmos.compute_probs, row N1) --------------------------------
    @torch.no_grad()
    def compute_probs(self, aoi_count=None, particles=50, aoi_chunk=None, local_noise=None, global_noise=None,
                      ndx=None, fdx=None, seed_offset=1 << 40):
        """
        z_probs (n, F, C, 2) and theta_probs (K, n, F, C) for local AOIs [0, aoi_count) and all frames
        (reference: cosmos.py:609-672, 50 guide particles).  For every particle the guide is sampled
        (tq_cosmos_globals_sample + tq_cosmos_sites) and tq_cosmos_zprobs accumulates the posterior of
        (z, theta) given that draw.  ``local_noise`` / ``global_noise`` (lists, one per particle) and
        explicit ``ndx`` / ``fdx`` replay given draws for the parity test.
        """
        lib, code, p = self.lib, self.code, _lib.ptr
        mc = ctypes.byref(self.mc)
        dev, dtype = self.device, self.dtype
        n = self.Nt if aoi_count is None else int(aoi_count)
        if ndx is not None:
            chunks = [ndx.to(torch.int32)]
        else:
            step = aoi_chunk or max(1, min(n, (1 << 22) // max(self.F * self.C, 1)))
            chunks = [torch.arange(lo, min(lo + step, n), dtype=torch.int32, device=dev) for lo in range(0, n, step)]
        fb = self.F if fdx is None else len(fdx)
        z_out, th_out = [], []
        rows = lib.tq_site_record_rows()
        for chunk in chunks:
            nbc = len(chunk)
            U = nbc * fb * self.C
            s = self.store
            view = _lib.make_view(s.pixels, s.xy, s.offset_samples, s.offset_logits, nb=nbc, fb=fb, C=self.C, F=self.F,
                                  P=self.P, ndx=chunk, fdx=fdx, is_ontarget=s.is_ontarget, mask=s.mask)
            samples = torch.empty(L.NSAMP, U, dtype=dtype, device=dev)
            qm = torch.empty(4, U, dtype=dtype, device=dev)
            rec = torch.empty(rows, U, dtype=dtype, device=dev)
            z_probs = torch.zeros(nbc, fb, self.C, 2, dtype=dtype, device=dev)
            theta_probs = torch.zeros(L.K, nbc, fb, self.C, dtype=dtype, device=dev)
            with torch.cuda.device(dev):
                st = _lib.stream_ptr(dev)
                for i in range(particles):
                    pstate = torch.tensor([seed_offset + i], dtype=torch.int64, device=dev)
                    gn = None if global_noise is None else global_noise[i]
                    ln = None if local_noise is None else local_noise[i]
                    _lib.check(lib.tq_cosmos_globals_sample(code, self.C, p(self.gparams), mc, p(gn), self.seed, p(pstate),
                                                            p(self.gstate), p(self.tables), p(self.gain), st),
                               "tq_cosmos_globals_sample")
                    _lib.check(lib.tq_cosmos_sites(code, view, self.Nt, mc, p(self.lparams), self.aoi_offset, self.seed,
                                                   p(pstate), p(ln), p(samples), p(qm), p(rec), st), "tq_cosmos_sites")
                    _lib.check(lib.tq_cosmos_zprobs(code, view, self.Nt, mc, p(self.lparams), p(self.tables), p(samples),
                                                    1.0 / particles, p(z_probs), p(theta_probs), st), "tq_cosmos_zprobs")
            z_out.append(z_probs)
            th_out.append(theta_probs)
        return torch.cat(z_out, 0), torch.cat(th_out, 1)

    def close(self):
        """Release what the caching allocator does not own: the captured graph and the peer-memory buffer with its IPC
        mappings (an engine that is replaced, e.g. by the NaN restart of Model.run, must not leak them)."""
        self._graph, self._eager_default_steps = None, 0
        if self.p2p is not None:
            torch.cuda.synchronize(self.device)
            self.p2p.close()
            self.p2p = None

    def release_graph(self):
        """Drop the captured CUDA graph (it keeps references to NCCL work when world_size > 1)."""
        self._graph, self._eager_default_steps = None, 0

    @property
    def iteration(self):
        return int(self.state[0].item())

    def set_iteration(self, it):
        self.flush_deferred()   # a pending update belongs to the count it was recorded with
        self.state[0] = int(it)
