"""
cosmos -- multi-colour time-independent colocalization model, B200-native SVI path.

Same class name, registry key, constructor, variational-parameter names / shapes / constraints and
checkpoint layout as the reference (tapqir/models/cosmos.py:28-80, 464-598).  The generative
model and guide of cosmos.py:82-462 are not re-expressed in an effect DSL: their ELBO (SURVEY.md
App. A.3) and its gradient are evaluated by the kernels in ``csrc/`` through
:class:`tapqir_b200.models.engine.CosmosEngine`.
"""

import math
from collections import OrderedDict

import torch
import torch.distributions.constraints as constraints
from torch.distributions import transform_to

from tapqir_b200.models import layout as L
from tapqir_b200.models.model import Model

DEFAULT_PRIORS = {  # cosmos.py:55-64
    "background_mean_std": 1000.0,
    "background_std_std": 100.0,
    "lamda_rate": 1.0,
    "height_std": 10000.0,
    "width_min": 0.75,
    "width_max": 2.25,
    "proximity_rate": 1.0,
    "gain_std": 50.0,
}


class cosmos(Model):
    r"""
    **Multi-Color Time-Independent Colocalization Model**

    Ordabayev YA, Friedman LJ, Gelles J, Theobald DL. *Bayesian machine learning analysis of
    single-molecule fluorescence colocalization images.* eLife 2022, doi:10.7554/eLife.73860.

    :param K: Maximum number of spots that can be present in a single image (kernels: K = 2).
    :param device: Computation device.
    :param dtype: "float" or "double".
    :param use_pykeops: accepted for signature compatibility; the offset marginalisation always runs
        in the fused CUDA kernel.
    :param priors: Dictionary of parameters of prior distributions.
    :param ref_dtype: dtype whose ``eps``/``tiny`` the clamping conventions follow; the reference CLI
        always runs in double (main.py:428), so that is the default whatever ``dtype`` is.
    """

    name = "cosmos"

    def __init__(self, S: int = 1, K: int = 2, Q: int = None, device: str = "cuda", dtype: str = "float",
                 use_pykeops: bool = True, priors: dict = None, ref_dtype: str = "double"):
        super().__init__(S=S, K=K, Q=Q, device=device, dtype=dtype, priors=dict(priors or DEFAULT_PRIORS))
        if K != L.K or S != L.S:
            raise NotImplementedError(f"the sm_100a kernels are built for K={L.K}, S={L.S}")
        self._global_params = ["gain", "proximity", "lamda", "pi"]
        self.use_pykeops = use_pykeops
        self.ref_dtype = getattr(torch, ref_dtype)
        self.conv_params = ["-ELBO", "proximity_loc", "gain_loc", "lamda_loc"]
        self.ci_params = ["gain", "pi", "lamda", "proximity", "background", "height", "width", "x", "y"]

    # ---- the Pyro-facing hooks of the reference have no counterpart here ------------------------------
    def model(self):
        raise NotImplementedError("evaluated by csrc/cosmos_local.cuh + ksmogn_core.cuh, see DESIGN.md")

    def guide(self):
        raise NotImplementedError("evaluated by csrc/cosmos_local.cuh, see DESIGN.md")

    def TraceELBO(self, jit=False):
        raise NotImplementedError("the enumerated ELBO is assembled in csrc/cosmos_local.cuh::local_post")

    # ---- variational parameters --------------------------------------------------------------------------
    def constraints(self):
        """name -> torch constraint, in the reference's creation order (cosmos.py:471-598)."""
        eps = torch.finfo(self.ref_dtype).eps
        P = self.data.P
        half = (P + 1) / 2
        return OrderedDict([
            ("pi_mean", constraints.simplex),
            ("pi_size", constraints.positive),
            ("m_probs", constraints.unit_interval),
            ("proximity_loc", constraints.interval(0, (P + 1) / math.sqrt(12) - eps)),
            ("proximity_size", constraints.greater_than(2.0)),
            ("lamda_loc", constraints.positive),
            ("lamda_beta", constraints.positive),
            ("gain_loc", constraints.positive),
            ("gain_beta", constraints.positive),
            ("background_mean_loc", constraints.positive),
            ("background_std_loc", constraints.positive),
            ("b_loc", constraints.positive),
            ("b_beta", constraints.positive),
            ("h_loc", constraints.positive),
            ("h_beta", constraints.positive),
            ("w_mean", constraints.interval(0.75 + eps, 2.25 - eps)),
            ("w_size", constraints.greater_than(2.0)),
            ("x_mean", constraints.interval(-half + eps, half - eps)),
            ("y_mean", constraints.interval(-half + eps, half - eps)),
            ("size", constraints.greater_than(2.0)),
        ])

    def _shard_sizes(self):
        """AOIs per rank: balanced contiguous blocks, the first ``Nt % world_size`` ranks hold one more."""
        Nt, w = self.data.Nt, self.world_size
        if getattr(self, "presharded", False):
            return [Nt] * w
        return [Nt // w + (1 if r < Nt % w else 0) for r in range(w)]

    def _shard(self):
        """Contiguous AOI block of this rank."""
        if getattr(self, "presharded", False):
            return slice(0, self.data.Nt)
        sizes = self._shard_sizes()
        lo = sum(sizes[:self.rank])
        return slice(lo, lo + sizes[self.rank])

    def build_engine(self, seed=0):
        from tapqir_b200.models.engine import CosmosEngine

        if self.device.type != "cuda":
            raise RuntimeError("tapqir_b200 has no CPU execution path: construct the model with device='cuda'")
        sl = self._shard()
        if sl.stop - sl.start < 1:
            raise ValueError(f"{self.data.Nt} AOIs cannot be sharded over {self.world_size} ranks: rank {self.rank} would "
                             "hold none (use fewer GPUs)")
        # identical offset bins are merged on upload unless model.merge_offsets is set to False (utils/dataset.py)
        store = self.data.device_store(self.device, self.dtype, sl, merge_offsets=getattr(self, "merge_offsets", True))
        self.engine = CosmosEngine(
            store, sl.stop - sl.start, self.data.F, self.data.C, self.data.P, self.priors, dtype=self.dtype,
            lr=self.lr, betas=self.optim_args["betas"], nbatch_size=self.nbatch_size, fbatch_size=self.fbatch_size,
            seed=seed, ref_dtype=self.ref_dtype,
            Nt_total=self.data.Nt * (self.world_size if getattr(self, "presharded", False) else 1),
            aoi_offset=self.rank * self.data.Nt if getattr(self, "presharded", False) else sl.start, rank=self.rank,
            world_size=self.world_size, process_group=self.process_group, shard_sizes=self._shard_sizes())
        self.nbatch_size, self.fbatch_size = self.engine.nb, self.engine.fb
        return self.engine

    def init_parameters(self):
        """Initial values of cosmos.py:471-598, stored unconstrained (``transform_to(c).inv``)."""
        eng, data = self.engine, self.data
        dev, dt = self.device, torch.float64
        K, Q, C, F = self.K, self.Q, data.C, data.F
        Nt = eng.Nt
        bg = (data.median.to(dev, dt) - data.offset.mean)
        full = lambda shape, v: torch.full(shape, float(v), dtype=dt, device=dev)
        init = {
            "pi_mean": torch.ones(Q, self.S + 1, dtype=dt, device=dev),
            "pi_size": full((Q, 1), 2),
            "m_probs": full((K, Nt, F, Q), 0.5),
            "proximity_loc": full((), 0.5),
            "proximity_size": full((), 100),
            "lamda_loc": full((Q,), 0.5),
            "lamda_beta": full((Q,), 100),
            "gain_loc": full((), 5),
            "gain_beta": full((), 100),
            "background_mean_loc": bg.expand(Nt, 1, C),
            "background_std_loc": full((Nt, 1, C), 1),
            "b_loc": bg.expand(Nt, F, C),
            "b_beta": full((Nt, F, C), 1),
            "h_loc": full((K, Nt, F, Q), 2000),
            "h_beta": full((K, Nt, F, Q), 0.001),
            "w_mean": full((K, Nt, F, Q), 1.5),
            "w_size": full((K, Nt, F, Q), 100),
            "x_mean": full((K, Nt, F, Q), 0),
            "y_mean": full((K, Nt, F, Q), 0),
            "size": full((K, Nt, F, Q), 200),
        }
        cons = self.constraints()
        views = eng.named_unconstrained()
        for name, value in init.items():
            if name == "pi_mean":
                u = value.log()  # SoftmaxTransform.inv
            else:
                u = transform_to(cons[name]).inv(value)
            views[name].copy_(u.to(views[name].dtype).reshape(views[name].shape))
        for buf in (eng.lm, eng.lv, eng.gm, eng.gv, eng.lgrads, eng.ggrads):
            buf.zero_()
        eng.set_iteration(0)

    # ---- stepping ---------------------------------------------------------------------------------------------
    @property
    def launches_per_step(self):
        """Kernels of this library launched by one default step (for bench.py's gpu_launches)."""
        eng = self.engine
        # globals_sample, globals_prepare, site_fast, site_worklist, ksmogn, local_post, globals_finish, adam x2, advance
        # (dtype "double": one site kernel)
        n = 10 if eng.dtype == torch.float32 else 9
        if getattr(eng, "last_step_fused", False):
            n -= 3   # site_fast + site_worklist + ksmogn + local_post -> cosmos_fused_kernel
        if getattr(eng, "_pending", False):
            n += 1   # deferred local Adam: tq_adam_dense over the two ranges the site kernel does not own instead of one launch
        draws = (0 if eng.full_n else 1) + (0 if eng.full_f else 1)
        n += 1 if draws == 2 and eng.lib.tq_subsample_pair_supported(eng.Nt, eng.F) else draws
        return n

    def step_from_host(self, host_pixels, host_xy, loss_host, prefetch_next=None):
        """
        One step fed from HOST buffers, the way the reference feeds every step
        (utils/dataset.py:140-151: gather on the CPU, ``.to(device)``): the step's pinned host pixels +
        target locations are copied to the device, the step runs, the loss is copied back and awaited.

        ``prefetch_next=(pixels, xy)`` (pinned host tensors of the NEXT step) starts their upload on a
        copy stream into a staging buffer while this step computes; the next call then only pays a
        device-to-device move.  Every step's inputs still cross PCIe inside the caller's timed region.
        """
        eng, dev = self.engine, self.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_feed", None) is None:
            self._feed = {
                "copy": torch.cuda.Stream(device=dev),
                "pix": torch.empty_like(eng.store.pixels), "xy": torch.empty_like(eng.store.xy),
                "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "pending": None,
            }
            self._feed["free"].record(main)
        fd = self._feed
        if fd["pending"] is not None and fd["pending"] == (host_pixels.data_ptr(), host_xy.data_ptr()):
            main.wait_event(fd["ready"])                      # upload issued during the previous step
            eng.store.pixels.copy_(fd["pix"], non_blocking=True)
            eng.store.xy.copy_(fd["xy"], non_blocking=True)
        else:
            eng.store.pixels.copy_(host_pixels, non_blocking=True)
            eng.store.xy.copy_(host_xy, non_blocking=True)
        fd["free"].record(main)                               # staging buffers may be overwritten from here on
        fd["pending"] = None
        loss = self.step()
        loss_host.copy_(loss, non_blocking=True)
        if prefetch_next is not None:
            npix, nxy = prefetch_next
            fd["copy"].wait_event(fd["free"])
            with torch.cuda.stream(fd["copy"]):
                fd["pix"].copy_(npix, non_blocking=True)
                fd["xy"].copy_(nxy, non_blocking=True)
                fd["ready"].record(fd["copy"])
            fd["pending"] = (npix.data_ptr(), nxy.data_ptr())
        main.synchronize()
        return float(loss_host.item())

    # ---- posterior summaries -----------------------------------------------------------------------------------
    @property
    def compute_probs(self):
        """
        ``(z_probs (Nt,F,Q,1+S), theta_probs (K,Nt,F,Q))`` from 50 guide particles for the on-target AOIs;
        off-target AOIs stay 0 (reference: cosmos.py:609-672).  Cached; this rank's AOI block only.
        """
        if getattr(self, "_probs", None) is None:
            eng, data = self.engine, self.data
            sl = self._shard()
            ont = data.is_ontarget[sl]
            n_on = int(ont.sum().item())
            assert bool(ont[:n_on].all()), "on-target AOIs must come first (as written by glimpse/simulate)"
            z_probs = torch.zeros(eng.Nt, data.F, self.Q, 1 + self.S, dtype=eng.dtype)
            theta_probs = torch.zeros(self.K, eng.Nt, data.F, self.Q, dtype=eng.dtype)
            if n_on:
                z, th = eng.compute_probs(aoi_count=n_on, particles=50)
                z_probs[:n_on] = z.cpu()
                theta_probs[:, :n_on] = th.cpu()
            self._probs = (z_probs, theta_probs)
        return self._probs

    @property
    def z_probs(self) -> torch.Tensor:
        r"""Probability of there being a target-specific spot :math:`p(z=1)` (cosmos.py:674-679)."""
        return self.compute_probs[0]

    @property
    def theta_probs(self) -> torch.Tensor:
        r"""Posterior target-specific spot probability :math:`q(\theta = k)`, k = 1..K (cosmos.py:681-686)."""
        return self.compute_probs[1]

    @property
    def pspecific(self) -> torch.Tensor:
        return self.z_probs

    @property
    def z_map(self) -> torch.Tensor:
        return torch.argmax(self.z_probs, dim=-1)

    @torch.no_grad()
    def compute_params(self, CI):
        """
        Mean and ``CI`` credible interval of every guide distribution + the posterior summaries
        (reference: cosmos.py:711-784; its scipy inverse CDFs, stats.py:262-293, evaluated on the device over whole arrays).
        """
        from tapqir_b200.utils.stats import credible_intervals

        value = lambda name: self.param(name).detach().double()
        params = credible_intervals(self.ci_params, value, self.data.P, self.priors, CI, self.device)
        params["m_probs"] = self.m_probs.cpu()
        params["z_probs"] = self.z_probs.cpu()
        params["theta_probs"] = self.theta_probs.cpu()
        params["z_map"] = self.z_map.cpu()
        params["p_specific"] = params["theta_probs"].sum(0)
        return params

    @property
    def m_probs(self) -> torch.Tensor:
        r"""Posterior spot presence probability :math:`q(m=1)` (cosmos.py:688-693)."""
        return self.param("m_probs").detach()
