"""
Base model driver with the reference's public interface (tapqir/models/model.py:31-371):
``Model(S, K, Q, device, dtype, priors)``, ``.load``, ``.init``, ``.run``, ``.save_checkpoint``,
``.load_checkpoint``, ``.compute_stats``.  The SVI loop body (``svi.step()``, model.py:212) is the
CUDA step engine instead of Pyro; the checkpoint file keeps the reference's dict layout
(model.py:273-282) so ``.tapqir/<name>_model.tpqr`` files stay interchangeable.

Differences that are deliberate:
* no CPU execution path: ``init``/``run`` need a CUDA device and the sm_100a library;
* ``to()`` does not flip torch's global default tensor type (model.py:82-91 does);
* the loss is left on the device between checkpoints (the reference pays a host sync per step for
  ``loss.item()``); ``iter_loss`` synchronises when read.
"""

import logging
import random
from collections import defaultdict, deque
from pathlib import Path
from typing import Union

import torch

from tapqir_b200 import __version__ as tapqir_version
from tapqir_b200.exceptions import (CudaOutOfMemoryError, NonFiniteParameterError, PeerTimeoutError,
                                    TapqirFileNotFoundError)
from tapqir_b200.utils.dataset import load

logger = logging.getLogger(__name__)


class Model:
    """
    Base class of the cosmos-family models.

    :param S: Number of distinct molecular states for the binder molecules.
    :param K: Maximum number of spots that can be present in a single image.
    :param Q: Number of fluorescent dyes.
    :param device: Computation device ("cuda", "cuda:1", ...; "cpu" only for loading / inspection).
    :param dtype: Floating point precision, "float" (production) or "double" (the reference CLI's).
    :param priors: Dictionary of parameters of prior distributions.
    """

    name = None
    conv_params = ["-ELBO"]

    def __init__(self, S: int = 1, K: int = 2, Q: int = None, device: str = "cuda", dtype: str = "float",
                 priors: dict = None):
        self.S = S
        self.K = K
        self._Q = Q
        self.nbatch_size = None
        self.fbatch_size = None
        self.priors = priors
        self.n = None
        self.f = None
        self.data_path = None
        self.path = None
        self.run_path = None
        self.engine = None
        self.iter = 0
        self.converged = False
        self._loss_dev = None
        # defaults of everything init() sets, so that the reference's `tapqir stats` flow (main.py:566: load();
        # load_checkpoint(param_only=True); compute_stats()) works on a model that was never init()-ed
        self.lr = 0.005
        self.optim_args = {"lr": self.lr, "betas": [0.9, 0.999]}
        self.rank, self.world_size, self.process_group = 0, 1, None
        self.presharded = False
        self.seed = 0
        self._rolling = defaultdict(lambda: deque([], maxlen=100))
        self.to(device, dtype)

    def to(self, device: str, dtype: str = "double") -> None:
        """Change computation device and floating point precision ("double" or "float")."""
        self.dtype = getattr(torch, dtype)
        self.device = torch.device(device)

    @property
    def Q(self):
        return self._Q or self.data.C

    def load(self, path: Union[str, Path], data_only: bool = True) -> None:
        """Load ``data.tpqr`` and optionally the fit results from a Tapqir analysis folder."""
        import pandas as pd

        self.path = Path(path)
        self.run_path = self.path / ".tapqir"
        self.data = load(self.path, self.device)
        logger.debug(f"Loaded data from {self.path / 'data.tpqr'}")
        if not data_only:
            try:
                self.params = torch.load(self.path / f"{self.name}_params.tpqr", weights_only=False)
            except FileNotFoundError:
                raise TapqirFileNotFoundError("parameter", self.path / f"{self.name}_params.tpqr")
            try:
                self.summary = pd.read_csv(self.path / f"{self.name}_summary.csv", index_col=0)
            except FileNotFoundError:
                raise TapqirFileNotFoundError("summary", self.path / f"{self.name}_summary.csv")

    # ---- to be provided by the concrete model ---------------------------------------------------------
    def init_parameters(self):
        raise NotImplementedError

    def build_engine(self, seed):
        raise NotImplementedError

    # ---- SVI ------------------------------------------------------------------------------------------------
    def init(self, lr: float = 0.005, nbatch_size: int = 5, fbatch_size: int = 512, jit: bool = False,
             rank: int = 0, world_size: int = 1, process_group=None, seed: int = 0, presharded: bool = False) -> None:
        """
        Initialize the SVI state (reference: model.py:153-186): Adam(lr, betas (0.9, 0.999)); resume from
        ``.tapqir/<name>_model.tpqr`` if present, else initialise the variational parameters.

        ``rank/world_size/process_group``: AOI-sharded data parallelism -- this process owns the
        contiguous AOI block ``rank`` of ``world_size`` (``presharded=True``: ``self.data`` already IS
        that block, every rank holding the same number of AOIs).  ``jit`` is accepted and ignored.
        """
        self.lr = lr
        self.optim_args = {"lr": lr, "betas": [0.9, 0.999]}
        self.rank, self.world_size, self.process_group = rank, world_size, process_group
        self.presharded = presharded
        self.nbatch_size = min(nbatch_size, self.data.Nt)
        self.fbatch_size = min(fbatch_size, self.data.F)
        self.seed = seed
        if self.engine is not None:
            self.engine.close()   # peer-memory buffers / IPC mappings of the engine being replaced (restart path)
        self.build_engine(seed)
        try:
            self.load_checkpoint()
        except TapqirFileNotFoundError:
            self.iter = 0
            self.converged = False
            self._rolling = defaultdict(lambda: deque([], maxlen=100))
            self.init_parameters()

    def step(self, **kw):
        """One SVI step; returns the device tensor holding the loss (no host sync)."""
        self._loss_dev = self.engine.step(**kw)
        return self._loss_dev

    @property
    def iter_loss(self):
        """-ELBO of the latest step as the reference reports it (model.py:212, 285-298), including the constant Pyro's
        enumeration contributes for masked AOIs (4 configurations x ln 6 states per masked unit, times the plate
        scales: the sites' log-probabilities are zeroed but still summed over; tests/test_oracle.py).  Synchronises;
        with ``world_size > 1`` every rank must read it (one small all-reduce)."""
        if self._loss_dev is None:
            return float("nan")
        return float(self._loss_dev.item()) + self.masked_loss_constant()

    def masked_loss_constant(self):
        import math

        eng = self.engine
        masked = (eng.store.mask == 0)
        if eng.ndx is not None:
            masked = masked[eng.ndx.long()]
        count = masked.sum().to(torch.float64)
        if self.world_size > 1:
            torch.distributed.all_reduce(count, group=self.process_group)
        per_unit = getattr(self, "masked_unit_constant", 4 * math.log(6))
        return float(count.item()) * eng.fb * eng.C * per_unit * eng.sN * eng.sF

    def run(self, num_iter: int = 0, progress_bar=None) -> None:
        """
        Run inference for ``num_iter`` iterations; 0 = until the convergence criterion holds (cap
        100000).  Reference: model.py:188-237.
        """
        from torch.utils.tensorboard import SummaryWriter

        if progress_bar is None:
            from tqdm import tqdm as progress_bar
        use_crit = False
        if not num_iter:
            use_crit = True
            num_iter = 100000
        logger.debug("Tapqir(b200) version - {}".format(tapqir_version))
        logger.debug("Model - {}".format(self.name))
        logger.debug("Device - {}".format(self.device))
        logger.debug("Floating precision - {}".format(self.dtype))
        logger.debug("Optimizer - Adam")
        logger.debug("Learning rate - {}".format(self.lr))
        logger.debug("AOI batch size - {}".format(self.nbatch_size))
        logger.debug("Frame batch size - {}".format(self.fbatch_size))

        writer = SummaryWriter(log_dir=self.run_path / "logs" / self.name) if (self.run_path and self.rank == 0) else None
        try:
            for i in progress_bar(range(num_iter)):
                try:
                    self.step()
                    if not self.iter % 200:  # checkpoint cadence of the reference
                        self.save_checkpoint(writer)
                        if use_crit and self.converged:
                            logger.info(f"Iteration #{self.iter} model converged.")
                            break
                    self.iter += 1
                except NonFiniteParameterError:
                    # NaN/Inf found at checkpoint time (on ANY rank: the scan is all-reduced, so every rank is here): go
                    # back to the last checkpoint with a new seed -- rank 0 draws it, everyone uses it (model.py:220-232)
                    seed_box = [random.randint(0, 100)]
                    if self.world_size > 1:
                        torch.distributed.broadcast_object_list(seed_box, src=0, group=self.process_group)
                    new_seed = seed_box[0]
                    self.init(lr=self.lr, nbatch_size=self.nbatch_size, fbatch_size=self.fbatch_size, rank=self.rank,
                              world_size=self.world_size, process_group=self.process_group, seed=new_seed,
                              presharded=self.presharded)
                    logger.warning(f"Iteration #{self.iter} restarting with a new seed: {new_seed}.")
                except RuntimeError as err:
                    if str(err.args[0]).startswith("CUDA out of memory") or "out of memory" in str(err):
                        raise CudaOutOfMemoryError()
                    raise
            else:
                logger.warning(f"Iteration #{self.iter} model has not converged.")
            if self.world_size > 1 and self.run_path is not None:
                self.consolidate_checkpoint()   # reference-layout file: single-GPU resume, `tapqir stats`
        finally:
            if writer is not None:
                writer.close()

    # ---- parameter store view -----------------------------------------------------------------------------------
    def constraints(self):
        raise NotImplementedError

    def param_state(self):
        """``pyro.get_param_store().get_state()`` layout [third party]: unconstrained tensors + constraints."""
        return {"params": {k: v.detach().clone() for k, v in self.engine.named_unconstrained().items()},
                "constraints": dict(self.constraints())}

    def param(self, name):
        """Constrained value of a variational parameter (``pyro.param(name)`` in the reference)."""
        from torch.distributions import transform_to

        return transform_to(self.constraints()[name])(self.engine.named_unconstrained()[name])

    def save_checkpoint(self, writer=None):
        """
        Reference: model.py:239-323.  NaN/Inf scan (-> ValueError), rolling convergence criterion,
        ``torch.save`` of {iter, params, optimizer, rolling, convergence_status}, tensorboard scalars.
        """
        eng = self.engine
        if eng.p2p is not None:
            seq = eng.p2p.timed_out()
            if seq:
                raise PeerTimeoutError(seq)
        flat_ok = (torch.isfinite(eng.lparams).all() & torch.isfinite(eng.gparams).all()).to(torch.int32)
        if self.world_size > 1:   # every rank takes the restart path together, or none does
            torch.distributed.all_reduce(flat_ok, op=torch.distributed.ReduceOp.MIN, group=self.process_group)
        if not bool(flat_ok):
            for k, v in eng.named_unconstrained().items():
                if not torch.isfinite(v).all():
                    raise NonFiniteParameterError("Iteration #{}. Detected NaN values in {}".format(self.iter, k))
            raise NonFiniteParameterError("Iteration #{}. Detected NaN values on another rank".format(self.iter))
        loss = self.iter_loss
        for name in self.conv_params:
            if name == "-ELBO":
                self._rolling["-ELBO"].append(loss)
            else:
                val = self.param(name)
                if val.ndim == 1:
                    for i in range(len(val)):
                        self._rolling[f"{name}_{i}"].append(val[i].item())
                else:
                    self._rolling[name].append(val.item())
        self.converged = False
        if len(self._rolling["-ELBO"]) == self._rolling["-ELBO"].maxlen:
            crit = all(torch.tensor(v).std() / torch.tensor(v)[-50:].std() < 1.05 for v in self._rolling.values())
            if crit:
                self.converged = True
        if self.run_path is not None:
            self.run_path.mkdir(parents=True, exist_ok=True)
            suffix = "" if self.world_size == 1 else f".rank{self.rank}"
            torch.save(
                {
                    "iter": self.iter,
                    "params": self.param_state(),
                    "optimizer": self.optim_state(),
                    "rolling": dict(self._rolling),
                    "convergence_status": self.converged,
                },
                self.run_path / f"{self.name}_model.tpqr{suffix}",
            )
        if writer is not None:
            writer.add_scalar("-ELBO", loss, self.iter)
            for name in self.constraints():
                val = self.param(name)
                if val.dim() == 0:
                    writer.add_scalar(name, val.item(), self.iter)
                elif val.dim() == 1 and len(val) <= self.Q * 2:
                    writer.add_scalars(name, {str(i): v.item() for i, v in enumerate(val)}, self.iter)
                elif val.dim() == 2 and len(val) <= self.Q * 2:
                    writer.add_scalars(name, {f"{i}_{j}": k.item() for i, v in enumerate(val) for j, k in enumerate(v)},
                                       self.iter)
        logger.debug(f"Iteration #{self.iter}: Successful.")

    def optim_state(self):
        """``pyro.optim.Adam.get_state()`` layout [third party]: {param name: torch Adam state_dict}."""
        eng = self.engine
        local_named = getattr(eng.ll, "named", eng.ll.views)   # reference names / shapes (hmm: m_probs is stacked)
        moments = dict(local_named(eng.lm))
        moments.update(eng.gl.views(eng.gm))
        second = dict(local_named(eng.lv))
        second.update(eng.gl.views(eng.gv))
        step = torch.tensor(float(eng.iteration))
        group = {"lr": self.lr, "betas": (0.9, 0.999), "eps": eng.adam_eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": [0]}
        return {k: {"state": {0: {"step": step.clone(), "exp_avg": moments[k].detach().clone(),
                                  "exp_avg_sq": second[k].detach().clone()}},
                    "param_groups": [dict(group)]} for k in moments}

    # ---- AOI-sharded checkpoints <-> the reference's single file ------------------------------------------------
    def _aoi_axis(self, name, ndim):
        """Axis of the AOI dimension of parameter ``name`` in the reference's shapes (None: a global parameter)."""
        if self.engine is not None and name in self.engine.gl.shapes:
            return None
        if name in ("background_mean_loc", "background_std_loc", "b_loc", "b_beta", "z_trans"):
            return 0
        if ndim == 5:      # hmm m_probs (1+S, K, Nt, F, Q)
            return 2
        return 1 if ndim == 4 else None

    def consolidate_checkpoint(self):
        """
        Multi-GPU fits checkpoint per rank (``<name>_model.tpqr.rank<r>``: every rank writes its AOI block in parallel,
        nothing crosses NVLink).  This merges the latest rank files into ONE file in the reference's layout
        (model.py:273-282: the AOI axes concatenated in rank order), so that a fit made on N GPUs resumes on one GPU,
        on a different number of GPUs, or feeds ``tapqir stats``.  Collective: every rank calls it (``run`` does, when it
        ends); rank 0 writes.
        """
        if self.world_size == 1 or self.run_path is None:
            return
        torch.distributed.barrier(group=self.process_group)
        if self.rank == 0:
            parts = [torch.load(self.run_path / f"{self.name}_model.tpqr.rank{r}", map_location="cpu", weights_only=False)
                     for r in range(self.world_size)]
            if len({p["iter"] for p in parts}) != 1:
                raise RuntimeError(f"rank checkpoints of different iterations: {[p['iter'] for p in parts]}")

            def cat(name, tensors):
                axis = self._aoi_axis(name, tensors[0].dim())
                return tensors[0] if axis is None else torch.cat(tensors, axis)

            merged = parts[0]
            merged["params"]["params"] = {k: cat(k, [p["params"]["params"][k] for p in parts]) for k in merged["params"]["params"]}
            for k, entry in merged["optimizer"].items():
                for key in ("exp_avg", "exp_avg_sq"):
                    entry["state"][0][key] = cat(k, [p["optimizer"][k]["state"][0][key] for p in parts])
            torch.save(merged, self.run_path / f"{self.name}_model.tpqr")
        torch.distributed.barrier(group=self.process_group)

    def load_checkpoint(self, path: Union[str, Path] = None, param_only: bool = False, warnings: bool = False):
        """Reference: model.py:325-357.  With ``world_size > 1`` a rank reads its own ``.rank<r>`` file if there is one,
        else its AOI block of the single reference-layout file (:meth:`consolidate_checkpoint`, or a single-GPU fit)."""
        path = Path(path) if path else self.run_path
        if path is None:
            raise TapqirFileNotFoundError("model", f"{self.name}_model.tpqr")
        model_path = path / f"{self.name}_model.tpqr"
        shard_path = path / f"{self.name}_model.tpqr.rank{self.rank}"
        use_shard = self.world_size > 1 and shard_path.exists()
        try:
            checkpoint = torch.load(shard_path if use_shard else model_path, map_location=self.device, weights_only=False)
        except FileNotFoundError:
            raise TapqirFileNotFoundError("model", model_path)
        if self.engine is None:
            self.build_engine(self.seed)
        eng = self.engine
        if self.world_size > 1 and not use_shard:
            lo = self.rank * eng.Nt if self.presharded else self._shard().start

            def block(name, t):
                axis = self._aoi_axis(name, t.dim())
                return t if axis is None else t.narrow(axis, lo, eng.Nt)

            checkpoint["params"]["params"] = {k: block(k, v) for k, v in checkpoint["params"]["params"].items()}
            for k, entry in checkpoint["optimizer"].items():
                for key in ("exp_avg", "exp_avg_sq"):
                    entry["state"][0][key] = block(k, entry["state"][0][key])
        eng.load_unconstrained(checkpoint["params"]["params"])
        if not param_only:
            self.converged = checkpoint["convergence_status"]
            self._rolling = defaultdict(lambda: deque([], maxlen=100), checkpoint["rolling"])
            self.iter = checkpoint["iter"]
            opt = checkpoint["optimizer"]
            for lflat, gflat, key in ((eng.lm, eng.gm, "exp_avg"), (eng.lv, eng.gv, "exp_avg_sq")):
                if hasattr(eng.ll, "load_named"):
                    eng.ll.load_named(lflat, {k: opt[k]["state"][0][key] for k in opt})
                    views = eng.gl.views(gflat)
                else:
                    views = dict(eng.ll.views(lflat), **eng.gl.views(gflat))
                for k, v in views.items():
                    v.copy_(opt[k]["state"][0][key].to(device=self.device, dtype=v.dtype).reshape(v.shape))
            steps = [int(opt[k]["state"][0]["step"]) for k in opt]
            eng.set_iteration(max(steps) if steps else 0)
            logger.info(f"Iteration #{self.iter}. Loaded a model checkpoint from {model_path}")
        if warnings and not checkpoint["convergence_status"]:
            logger.warning(f"Model at {path} has not been fully trained")

    def compute_stats(self, CI: float = 0.95, save_matlab: bool = False):
        """Credible intervals and summary statistics (reference: model.py:359-371).  Single process: after a multi-GPU
        fit, load the consolidated checkpoint on one GPU (``load``; ``load_checkpoint(param_only=True)``)."""
        from tapqir_b200.utils.stats import save_stats

        if self.world_size != 1:
            raise RuntimeError("compute_stats summarises the whole dataset on one GPU: run it in a single process on the "
                               "consolidated checkpoint (Model.consolidate_checkpoint)")

        try:
            save_stats(self, self.path, CI=CI, save_matlab=save_matlab)
        except RuntimeError as err:
            if "out of memory" in str(err):
                raise CudaOutOfMemoryError()
            raise
        logger.debug("Computing stats: Successful.")
