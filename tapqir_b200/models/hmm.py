"""
hmm -- multi-colour hidden-Markov colocalization model (reference: tapqir/models/hmm.py, registry key
"cosmos+hmm"), B200-native SVI path.

Same constructor, parameter names / shapes / constraints (hmm.py:419-467 on top of cosmos') as the reference.  The
ELBO that Pyro's ``TraceEnum_ELBO`` / ``TraceMarkovEnum_ELBO`` assembles for the Markov guide is
evaluated by :class:`tapqir_b200.models.hmm_engine.HmmEngine` (csrc/cosmos_hmm.cuh): forward marginals of the guide's
chain, the cosmos likelihood kernel with weights sum_z a_f(z) q_f(m|z), per-state emission terms, a backward recursion
for the gradients of ``z_trans``.  Built: the SVI step, ``z_probs`` / ``z_map`` / ``m_probs`` / ``pspecific``,
``theta_probs`` (5 guide particles given ``z_MAP``), ``z_sample``, checkpoints.
"""

from collections import OrderedDict

import torch
import torch.distributions.constraints as constraints
from torch.distributions import transform_to

from tapqir_b200.models.cosmos import DEFAULT_PRIORS, cosmos


class hmm(cosmos):
    r"""
    **Multi-Color Hidden Markov Colocalization Model**

    :param vectorized: accepted for signature compatibility (the chain is always evaluated by the recursion kernels).
    """

    name = "cosmos+hmm"
    # the constant Pyro's enumeration adds to the reported loss per masked unit is pinned for cosmos only (no hmm golden
    # with a masked AOI: the chain's enumeration is a different sum); reported as the ELBO of the unmasked AOIs
    masked_unit_constant = 0.0

    def __init__(self, S: int = 1, K: int = 2, Q: int = None, device: str = "cuda", dtype: str = "float",
                 use_pykeops: bool = True, vectorized: bool = True, priors: dict = None, ref_dtype: str = "double"):
        super().__init__(S=S, K=K, Q=Q, device=device, dtype=dtype, use_pykeops=use_pykeops,
                         priors=dict(priors or DEFAULT_PRIORS), ref_dtype=ref_dtype)
        self.vectorized = vectorized
        self._global_params = ["gain", "proximity", "lamda", "trans"]
        self.ci_params = ["gain", "init", "trans", "lamda", "proximity", "background", "height", "width", "x", "y"]

    def constraints(self):
        cons = OrderedDict([("init_mean", constraints.simplex), ("init_size", constraints.positive),
                            ("trans_mean", constraints.simplex), ("trans_size", constraints.positive),
                            ("z_trans", constraints.simplex)])
        for name, c in super().constraints().items():
            if name not in ("pi_mean", "pi_size"):
                cons[name] = c
        return cons

    def build_engine(self, seed=0):
        from tapqir_b200.models.hmm_engine import HmmEngine

        if self.device.type != "cuda":
            raise RuntimeError("tapqir_b200 has no CPU execution path: construct the model with device='cuda'")
        sl = self._shard()
        if sl.stop - sl.start < 1:
            raise ValueError(f"{self.data.Nt} AOIs cannot be sharded over {self.world_size} ranks: rank {self.rank} would "
                             "hold none (use fewer GPUs)")
        store = self.data.device_store(self.device, self.dtype, sl, merge_offsets=getattr(self, "merge_offsets", True))
        presharded = getattr(self, "presharded", False)
        self.engine = HmmEngine(
            store, sl.stop - sl.start, self.data.F, self.data.C, self.data.P, self.priors, dtype=self.dtype, lr=self.lr,
            betas=self.optim_args["betas"], nbatch_size=self.nbatch_size, seed=seed, ref_dtype=self.ref_dtype,
            Nt_total=self.data.Nt * (self.world_size if presharded else 1),
            aoi_offset=self.rank * self.data.Nt if presharded else sl.start, rank=self.rank, world_size=self.world_size,
            process_group=self.process_group, shard_sizes=self._shard_sizes())
        self.nbatch_size, self.fbatch_size = self.engine.nb, self.engine.fb
        return self.engine

    def init_parameters(self):
        """cosmos' initial values (cosmos.py ``_init_parameters``) + hmm.py:419-467, stored unconstrained."""
        eng, data = self.engine, self.data
        dev, dt = self.device, torch.float64
        K, Q, C, F, S = self.K, self.Q, data.C, data.F, self.S
        Nt = eng.Nt
        bg = (data.median.to(dev, dt) - data.offset.mean)
        full = lambda shape, v: torch.full(shape, float(v), dtype=dt, device=dev)
        init = {
            "init_mean": torch.ones(Q, S + 1, dtype=dt, device=dev), "init_size": full((Q, 1), 2),
            "trans_mean": torch.ones(Q, S + 1, S + 1, dtype=dt, device=dev), "trans_size": full((Q, S + 1, 1), 2),
            "z_trans": torch.ones(Nt, F, C, 1 + S, 1 + S, dtype=dt, device=dev),
            "m_probs": full((1 + S, K, Nt, F, C), 0.5),
            "proximity_loc": full((), 0.5), "proximity_size": full((), 100),
            "lamda_loc": full((Q,), 0.5), "lamda_beta": full((Q,), 100),
            "gain_loc": full((), 5), "gain_beta": full((), 100),
            "background_mean_loc": bg.expand(Nt, 1, C), "background_std_loc": full((Nt, 1, C), 1),
            "b_loc": bg.expand(Nt, F, C), "b_beta": full((Nt, F, C), 1),
            "h_loc": full((K, Nt, F, Q), 2000), "h_beta": full((K, Nt, F, Q), 0.001),
            "w_mean": full((K, Nt, F, Q), 1.5), "w_size": full((K, Nt, F, Q), 100),
            "x_mean": full((K, Nt, F, Q), 0), "y_mean": full((K, Nt, F, Q), 0), "size": full((K, Nt, F, Q), 200),
        }
        cons = self.constraints()
        unconstrained = {}
        for name, value in init.items():
            # SoftmaxTransform.inv = log: an all-ones simplex parameter is stored as zeros (uniform)
            unconstrained[name] = value.log() if cons[name] is constraints.simplex else transform_to(cons[name]).inv(value)
        eng.load_unconstrained(unconstrained)
        for buf in (eng.lm, eng.lv, eng.gm, eng.gv, eng.lgrads, eng.ggrads):
            buf.zero_()
        eng.set_iteration(0)

    @property
    def launches_per_step(self):
        # globals_sample, globals_prepare, site_fast, site_fallback, hmm_rows, hmm_forward, hmm_weights, ksmogn, local_post,
        # hmm_backward, hmm_reduce, globals_finish, adam x2, advance (+ subsample)
        eng = self.engine
        return (15 if eng.dtype == torch.float32 else 14) + (0 if eng.full_n else 1)

    # ---- posterior summaries (hmm.py:627-660) -----------------------------------------------------------------------------
    @property
    def z_probs(self) -> torch.Tensor:
        r"""Forward marginals of the guide's chain, :math:`p(z_f)`: ``(Nt, F, Q, 1+S)``."""
        return self.engine.z_probs().to(self.engine.dtype).cpu()

    @property
    def theta_probs(self) -> torch.Tensor:
        r"""Posterior target-specific spot probability :math:`q(\theta = k, z=z_\mathsf{MAP})` from 5 guide particles for
        the on-target AOIs (hmm.py:541-625, 639-644): ``(K, Nt, F, Q)``; off-target AOIs stay 0.  Cached."""
        if getattr(self, "_theta_probs", None) is None:
            eng, data = self.engine, self.data
            ont = data.is_ontarget[self._shard()]
            n_on = int(ont.sum().item())
            assert bool(ont[:n_on].all()), "on-target AOIs must come first (as written by glimpse/simulate)"
            out = torch.zeros(self.K, eng.Nt, data.F, self.Q, dtype=eng.dtype)
            if n_on:
                out[:, :n_on] = eng.compute_theta_probs(self.z_map[:n_on], aoi_count=n_on, particles=5).cpu()
            self._theta_probs = out
        return self._theta_probs

    @property
    def m_probs(self) -> torch.Tensor:
        r"""Posterior spot presence probability :math:`q(m=1, z=z_\mathsf{MAP})`: ``(K, Nt, F, Q)``."""
        mp = self.param("m_probs").detach().cpu()                  # (1+S, K, Nt, F, Q)
        zmap = self.z_map.long()                                   # (Nt, F, Q)
        return torch.gather(mp, 0, zmap[None, None].expand(1, *mp.shape[1:]))[0]

    @torch.no_grad()
    def compute_params(self, CI):
        """cosmos' summaries (with ``init`` / ``trans`` among the credible intervals) + the guide's chain ``z_trans``
        (hmm.py:669-676)."""
        params = super().compute_params(CI)
        params["z_trans"] = self.param("z_trans").detach().cpu()
        return params

    @torch.no_grad()
    def z_sample(self, num_samples):
        """
        ``num_samples`` draws of the guide's chain for the on-target AOIs, ``(num_samples, N, F, Q)`` (reference:
        hmm.py:662-672, which composes per-frame index maps with ``_sequential_index``; here the chain is simply walked).
        """
        zt = self.param("z_trans").detach()[: self.data.N]                      # (N, F, Q, z', z)
        N, F, Q = zt.shape[:3]
        out = torch.empty(num_samples, N, F, Q, dtype=torch.long, device=zt.device)
        prev = torch.zeros(num_samples, N, Q, dtype=torch.long, device=zt.device)   # row 0 is the initial distribution
        for f in range(F):
            rows = zt[:, f].expand(num_samples, N, Q, 2, 2)
            probs = torch.gather(rows, 3, prev[..., None, None].expand(num_samples, N, Q, 1, 2))[..., 0, :]
            prev = torch.multinomial(probs.reshape(-1, 2), 1).reshape(num_samples, N, Q)
            out[:, :, f] = prev
        return out

    def param(self, name):
        return transform_to(self.constraints()[name])(self.engine.named_unconstrained()[name])
