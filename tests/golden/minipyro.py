"""
Stand-in for the absent third-party packages ``pyro`` / ``pyroapi`` (pyro-ppl >= 1.8.5, reference ``setup.py:69``),
used ONLY by tests/golden/make_golden_step.py in the build container to run the reference's own
``tapqir/models/cosmos.py`` (``init_parameters``, ``guide``, ``model``, ``compute_probs``, ``compute_params``),
``tapqir/models/hmm.py`` (sequential form) and ``tapqir/models/model.py`` (``Model.init``, the ``svi.step()`` of the run
loop) UNMODIFIED.  Test infrastructure; nothing in ``tapqir_b200`` imports it and it never runs on the GPU box
(tests/test_minipyro.py checks it against closed-form answers on the CPU).

It restates, from Pyro's published semantics, exactly the pieces that code touches and nothing else:

* effect handlers: ``trace`` (+ ``get_trace`` / ``compute_log_prob``), ``replay``, ``block``, ``condition``,
  ``uncondition`` (+ ``Predictive`` built on them), ``mask``, ``plate``
  (subsampling as a replayable site, plate scale size / subsample_size, broadcasting of the distribution to the plate
  shape; sequential plates and ``markov`` loops without dimension recycling) and parallel enumeration
  (``infer={"enumerate": "parallel"}``: the support of the site is placed on a fresh tensor dimension to the left
  of ``max_plate_nesting``; the guide allocates first, the model continues to the left);
* ``pyro.param`` with the unconstrained value stored (``transform_to(constraint).inv``) and the constrained one
  returned; ``pyro.ops.indexing.Vindex``; ``pyro.distributions`` as thin subclasses of ``torch.distributions`` plus
  ``Delta`` and ``AffineBeta`` (TransformedDistribution(Beta, Affine) whose ``rsample`` clamps into
  ``[low + eps * scale, high - eps * scale]``);
* ``TraceEnum_ELBO``: per plate context, sum of the model sites' log-probabilities, log-sum-exp over the dimensions
  the MODEL enumerated, minus the guide sites' log-probabilities, weighted by the probabilities of the sites the
  GUIDE enumerated (exact expectation, differentiable), summed over plates and multiplied by the plate scale;
  reparameterised sites contribute pathwise gradients only.  The general tensor-variable-elimination of Pyro is not
  needed because every enumerated site of cosmos (and of hmm's sequential form) sits in the deepest plate context
  (asserted), where the expectation is evaluated by brute force over all enumerated dimensions;
* ``SVI.step`` (loss and gradients, one ``torch.optim.Adam`` per parameter acting on the unconstrained value, zero
  the gradients) and ``optim.Adam``.
"""

import sys
import types
from collections import OrderedDict

import torch
import torch.distributions as td
from torch.distributions import constraints, transform_to

_STACK = []
_PARAMS = OrderedDict()      # name -> (unconstrained leaf, constraint)
_ENUM = {"next": None}       # next free enumeration dimension (negative)


# ---------------------------------------------------------------------------------------------------------------------
# effect handlers
# ---------------------------------------------------------------------------------------------------------------------
class Messenger:
    fn = None

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        assert _STACK.pop() is self

    def process(self, msg):
        pass

    def postprocess(self, msg):
        pass

    def __call__(self, *a, **kw):      # handler(fn) form: run the wrapped callable under this handler
        with self:
            return self.fn(*a, **kw)


def apply_stack(msg):
    reached = 0
    for reached, h in enumerate(reversed(_STACK)):      # innermost handler first
        h.process(msg)
        if msg.get("stop"):
            break
    if msg["value"] is None:
        if msg["type"] == "sample":
            fn = msg["fn"]
            msg["value"] = fn.rsample() if fn.has_rsample else fn.sample()
        elif msg["type"] == "param":
            msg["value"] = _param_value(msg["name"], *msg["args"])
    for h in _STACK[len(_STACK) - reached - 1:] if _STACK else []:
        h.postprocess(msg)
    return msg


class Trace:
    """What ``handlers.trace(fn).get_trace()`` returns: the recorded sites and ``compute_log_prob``."""

    def __init__(self, nodes):
        self.nodes = nodes

    def compute_log_prob(self):
        for s in self.nodes.values():
            if s["type"] == "sample" and not s.get("subsample") and "log_prob" not in s:
                lp = s["fn"].log_prob(s["value"])
                s["unscaled_log_prob"] = lp                      # before scale AND mask, as in Pyro
                if s["mask"] is not None:
                    lp = torch.where(s["mask"], lp, lp.new_zeros(()))
                s["log_prob"] = lp * s["scale"]


LAST_TRACES = []     # every Trace handed out by get_trace(), so that a caller can look at a callee's traces


class trace(Messenger):
    def __init__(self, fn=None, param_only=False):
        self.fn, self.nodes, self.param_only = fn, OrderedDict(), param_only

    def get_trace(self, *a, **kw):
        self(*a, **kw)
        LAST_TRACES.append(Trace(self.nodes))
        return LAST_TRACES[-1]

    def postprocess(self, msg):
        if self.param_only and msg["type"] != "param":
            return
        if msg["type"] == "sample":
            assert msg["name"] not in self.nodes, f"duplicate site {msg['name']}"
        self.nodes[msg["name"]] = dict(msg)


class replay(Messenger):
    def __init__(self, fn=None, trace=None):
        self.fn, self.guide = fn, (trace.nodes if isinstance(trace, Trace) else trace)

    def process(self, msg):
        if msg["type"] == "sample" and msg["name"] in self.guide and not msg["is_observed"]:
            g = self.guide[msg["name"]]
            msg["value"], msg["done"], msg["infer"] = g["value"], True, g["infer"]


class block(Messenger):
    def __init__(self, fn=None, hide=()):
        self.fn, self.hide = fn, set(hide)

    def process(self, msg):
        if msg["name"] in self.hide:
            msg["stop"] = True


class condition(Messenger):
    """Fix the named sites to given values (they become observed)."""

    def __init__(self, fn=None, data=None):
        self.fn, self.data = fn, data

    def process(self, msg):
        if msg["type"] == "sample" and not msg.get("subsample") and msg["name"] in self.data:
            msg["value"], msg["is_observed"] = self.data[msg["name"]], True


class uncondition(Messenger):
    """Turn observed sites back into sampled ones (used to simulate data from a model)."""

    def __init__(self, fn=None):
        self.fn = fn

    def process(self, msg):
        if msg["type"] == "sample" and msg["is_observed"]:
            msg["infer"] = dict(msg["infer"], was_observed=True)
            msg["is_observed"], msg["value"], msg["done"] = False, None, False


class Predictive:
    """``pyro.infer.Predictive(model, posterior_samples=...)`` in its sequential form: for every index i along the leading
    axis of the given samples, run the model with those sites fixed to ``sample[i]`` and sample the rest; returns the
    remaining sample sites stacked along a new leading axis."""

    def __init__(self, model, posterior_samples=None, num_samples=None, **unused):
        self.model, self.samples = model, posterior_samples or {}
        sizes = {v.shape[0] for v in self.samples.values()}
        assert len(sizes) <= 1
        self.num_samples = sizes.pop() if sizes else num_samples

    def __call__(self, *a, **kw):
        out = {}
        for i in range(self.num_samples):
            tr = trace(condition(self.model, data={k: v[i] for k, v in self.samples.items()})).get_trace(*a, **kw)
            for name, site in tr.nodes.items():
                if site["type"] == "sample" and not site.get("subsample") and name not in self.samples:
                    out.setdefault(name, []).append(site["value"])
        return {k: torch.stack(v) for k, v in out.items()}


class mask(Messenger):
    def __init__(self, fn=None, mask=None):
        self.fn, self.mask = fn, mask

    def process(self, msg):
        if msg["type"] == "sample":
            msg["mask"] = self.mask if msg["mask"] is None else (self.mask & msg["mask"])


class enum(Messenger):
    """Parallel enumeration.  ``first_available_dim`` (guide) starts the allocation; ``None`` (model) continues it."""

    def __init__(self, first_available_dim=None):
        self.first = first_available_dim
        self.dims = []

    def __enter__(self):
        if self.first is not None:
            _ENUM["next"] = self.first
        return super().__enter__()

    def process(self, msg):
        if msg["type"] != "sample" or msg["done"] or msg["is_observed"] or msg.get("subsample"):
            return
        if msg["infer"].get("enumerate") != "parallel":
            return
        fn = msg["fn"]
        value = fn.enumerate_support(expand=False)            # (n,) + (1,) * len(batch_shape)
        actual, target = -1 - len(fn.batch_shape), _ENUM["next"]
        _ENUM["next"] -= 1
        assert target <= actual, "enumeration dimension collides with a batch dimension"
        value = value.reshape(value.shape[:1] + (1,) * (actual - target) + value.shape[1:])
        msg["value"], msg["done"] = value, True
        msg["infer"] = dict(msg["infer"], _enum_dim=target)
        self.dims.append(target)


class _Subsample(td.Distribution):
    """Uniform subsample of ``subsample_size`` out of ``size`` indices without replacement (pyro.plate)."""

    has_rsample = False
    arg_constraints = {}

    def __init__(self, size, subsample_size):
        self.size, self.subsample_size = size, subsample_size
        super().__init__(validate_args=False)

    def sample(self, sample_shape=torch.Size()):
        return torch.randperm(self.size)[: self.subsample_size].clone()

    def log_prob(self, x):
        return torch.zeros(())


class plate(Messenger):
    def __init__(self, name, size, subsample_size=None, subsample=None, dim=None, **unused):
        assert dim is None or dim < 0
        self.name, self.size, self.dim = name, size, dim
        if subsample is not None:
            self.indices = subsample
        elif subsample_size is None or subsample_size >= size:
            self.indices = torch.arange(size)
        else:    # a replayable site, so that model and guide see the same minibatch
            msg = _new_msg("sample", name, fn=_Subsample(size, subsample_size))
            msg["subsample"] = True
            self.indices = apply_stack(msg)["value"]
        self.subsample_size = len(self.indices)

    def __enter__(self):
        super().__enter__()
        return self.indices

    def __iter__(self):
        """Sequential plate (``for k in pyro.plate("spots", K)``): plain integers, no tensor dimension, no scaling."""
        assert self.dim is None and self.subsample_size == self.size
        return iter(range(self.size))

    def process(self, msg):
        if msg["type"] != "sample" or msg.get("subsample"):
            return
        msg["cond_indep_stack"] = (self,) + msg["cond_indep_stack"]
        msg["scale"] = msg["scale"] * (self.size / self.subsample_size)
        # broadcast the distribution to the plate shape
        fn = msg["fn"]
        actual = list(fn.batch_shape)
        target = [None if s == 1 else s for s in actual]
        target = [None] * (-self.dim - len(target)) + target
        assert target[self.dim] in (None, self.subsample_size), (msg["name"], self.name, actual)
        target[self.dim] = self.subsample_size
        for i in range(-len(target), 0):
            if target[i] is None:
                target[i] = actual[i] if len(actual) >= -i else 1
        if tuple(target) != tuple(actual):
            msg["fn"] = fn.expand(torch.Size(target))


def _new_msg(type_, name, fn=None, value=None, is_observed=False, infer=None, args=()):
    return {"type": type_, "name": name, "fn": fn, "value": value, "is_observed": is_observed, "infer": dict(infer or {}),
            "mask": None, "scale": 1.0, "cond_indep_stack": (), "done": False, "stop": False, "args": args}


# ---------------------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------------------
def sample(name, fn, obs=None, infer=None):
    msg = _new_msg("sample", name, fn=fn, value=obs, is_observed=obs is not None, infer=infer)
    return apply_stack(msg)["value"]


def _param_value(name, init=None, constraint=constraints.real):
    if name not in _PARAMS:
        assert init is not None, f"parameter {name} has not been initialised"
        value = init() if callable(init) else init
        with torch.no_grad():
            unconstrained = transform_to(constraint).inv(value.detach()).contiguous().clone()
        unconstrained.requires_grad_(True)
        _PARAMS[name] = (unconstrained, constraint)
    unconstrained, constraint = _PARAMS[name]
    return transform_to(constraint)(unconstrained)


def param(name, init_tensor=None, constraint=constraints.real, event_dim=None):
    msg = _new_msg("param", name, args=(init_tensor, constraint))
    return apply_stack(msg)["value"]


def markov(iterable, history=1):
    """``pyro.markov`` only lets enumeration dimensions be recycled ``history + 1`` steps later; without recycling every
    step keeps its own dimensions, which changes tensor shapes but no value."""
    return iterable


def clear_param_store():
    _PARAMS.clear()


class _ParamStore:
    def items(self):
        for k, (u, c) in _PARAMS.items():
            yield k, transform_to(c)(u).detach()

    def unconstrained(self):
        return OrderedDict((k, u) for k, (u, c) in _PARAMS.items())


def get_param_store():
    return _ParamStore()


def set_rng_seed(seed):
    torch.manual_seed(seed)


# ---------------------------------------------------------------------------------------------------------------------
# pyro.ops.indexing.Vindex
# ---------------------------------------------------------------------------------------------------------------------
def _is_batched(a):
    return isinstance(a, torch.Tensor) and a.dim() > 0


def vindex(tensor, args):
    """Advanced indexing in which tensor indices broadcast against each other on the LEFT and every full slice keeps
    its dimension on the right (in order)."""
    if not isinstance(args, tuple):
        return tensor[args]
    if not args:
        return tensor
    if args[0] is Ellipsis:
        args = args[1:]
        if not args:
            return tensor
        old_event_dim = len(args)
        args = (slice(None),) * (tensor.dim() - len(args)) + args
    else:
        args = args + (slice(None),) * (tensor.dim() - len(args))
        old_event_dim = len(args)
    assert len(args) == tensor.dim() and not any(a is Ellipsis for a in args)
    standard = True
    if tensor.dim() > old_event_dim and _is_batched(args[0]):
        standard = False
    elif any(_is_batched(a) for a in args[1:]):
        standard = False
    if standard:
        return tensor[args]
    new_event_dim = sum(isinstance(a, slice) for a in args[-old_event_dim:])
    new_dim = 0
    args = list(args)
    for i in reversed(range(len(args))):
        a = args[i]
        if isinstance(a, slice):
            assert a == slice(None), "only full slices"
            a = torch.arange(tensor.size(i)).reshape((-1,) + (1,) * new_dim)
            new_dim += 1
        elif _is_batched(a):
            a = a.reshape(a.shape + (1,) * new_event_dim)
        args[i] = a
    return tensor[tuple(args)]


class Vindex:
    def __init__(self, tensor):
        self.tensor = tensor

    def __getitem__(self, args):
        return vindex(self.tensor, args)


# ---------------------------------------------------------------------------------------------------------------------
# pyro.distributions
# ---------------------------------------------------------------------------------------------------------------------
class TorchDistributionMixin:
    def to_event(self, n=None):
        return Independent(self, n) if n else self


class TorchDistribution(TorchDistributionMixin, td.Distribution):
    def expand(self, batch_shape, _instance=None):
        if tuple(batch_shape) == tuple(self.batch_shape):
            return self
        raise NotImplementedError(f"{type(self).__name__}.expand to a different shape")


class Independent(TorchDistributionMixin, td.Independent):
    pass


def _wrap(cls):
    return type(cls.__name__, (TorchDistributionMixin, cls), {})


HalfNormal, Dirichlet, Exponential, Gamma = _wrap(td.HalfNormal), _wrap(td.Dirichlet), _wrap(td.Exponential), _wrap(td.Gamma)
Categorical, Bernoulli, Beta = _wrap(td.Categorical), _wrap(td.Bernoulli), _wrap(td.Beta)


class Delta(TorchDistribution):
    has_rsample = True
    arg_constraints = {"v": constraints.dependent, "log_density": constraints.real}
    support = constraints.real

    def __init__(self, v, log_density=0.0, event_dim=0, validate_args=None):
        assert event_dim == 0
        self.v, self.log_density = v, log_density
        super().__init__(v.shape, torch.Size(), validate_args=False)

    def expand(self, batch_shape, _instance=None):
        return Delta(self.v.expand(batch_shape), self.log_density)

    def rsample(self, sample_shape=torch.Size()):
        return self.v.expand(torch.Size(sample_shape) + self.v.shape)

    def log_prob(self, x):
        return (x == self.v).to(x.dtype).log() + self.log_density


class AffineBeta(TorchDistributionMixin, td.TransformedDistribution):
    """pyro.distributions.AffineBeta(concentration1, concentration0, loc, scale)."""

    arg_constraints = {"concentration1": constraints.positive, "concentration0": constraints.positive,
                       "loc": constraints.real, "scale": constraints.positive}

    def __init__(self, concentration1, concentration0, loc, scale, validate_args=None):
        base = td.Beta(concentration1, concentration0, validate_args=validate_args)
        super().__init__(base, td.AffineTransform(loc=loc, scale=scale), validate_args=validate_args)

    def expand(self, batch_shape, _instance=None):
        new = self._get_checked_instance(AffineBeta, _instance)
        return super().expand(batch_shape, _instance=new)

    def rsample(self, sample_shape=torch.Size()):
        x = self.base_dist.rsample(sample_shape)
        for t in self.transforms:
            x = t(x)
        eps = torch.finfo(x.dtype).eps * self.scale
        return torch.min(torch.max(x, torch.as_tensor(self.low + eps, dtype=x.dtype)),
                         torch.as_tensor(self.high - eps, dtype=x.dtype))

    def sample(self, sample_shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(sample_shape)

    loc = property(lambda self: torch.as_tensor(self.transforms[0].loc))
    scale = property(lambda self: torch.as_tensor(self.transforms[0].scale))
    low = property(lambda self: self.loc)
    high = property(lambda self: self.loc + self.scale)
    concentration1 = property(lambda self: self.base_dist.concentration1)
    concentration0 = property(lambda self: self.base_dist.concentration0)
    sample_size = property(lambda self: self.concentration1 + self.concentration0)
    mean = property(lambda self: self.loc + self.scale * self.base_dist.mean)
    variance = property(lambda self: self.scale.pow(2) * self.base_dist.variance)


def broadcast_shape(*shapes, **kw):
    return torch.broadcast_shapes(*[torch.Size(s) if not isinstance(s, torch.Size) else s for s in shapes])


# ---------------------------------------------------------------------------------------------------------------------
# TraceEnum_ELBO, SVI, Adam
# ---------------------------------------------------------------------------------------------------------------------
def _masked_log_prob(site):
    lp = site["fn"].log_prob(site["value"])
    if site["mask"] is not None:
        lp = torch.where(site["mask"], lp, lp.new_zeros(()))
    return lp


class TraceEnum_ELBO:
    def __init__(self, max_plate_nesting, ignore_jit_warnings=False, **unused):
        self.max_plate_nesting = max_plate_nesting

    def _traces(self, model, guide):
        g_enum, m_enum = enum(first_available_dim=-1 - self.max_plate_nesting), enum()
        gt = trace()
        with gt, g_enum:
            guide()
        mt = trace()
        with mt, replay(trace=gt.nodes), m_enum:
            model()
        return gt.nodes, mt.nodes, g_enum.dims, m_enum.dims

    def differentiable_elbo(self, model, guide):
        guide_nodes, model_nodes, guide_dims, model_dims = self._traces(model, guide)
        self.last_guide_trace, self.last_model_trace = guide_nodes, model_nodes
        sites = lambda nodes: [s for s in nodes.values() if s["type"] == "sample" and not s.get("subsample")]
        contexts = OrderedDict()
        for role, nodes in (("model", model_nodes), ("guide", guide_nodes)):
            for s in sites(nodes):
                key = frozenset(p.name for p in s["cond_indep_stack"])
                c = contexts.setdefault(key, {"model": [], "guide": [], "weights": [], "scale": s["scale"]})
                assert abs(c["scale"] - s["scale"]) < 1e-12 * c["scale"], "sites of one plate context share its scale"
                lp = _masked_log_prob(s)
                c[role].append(lp)
                if role == "guide" and "_enum_dim" in s["infer"]:
                    c["weights"].append(lp)
        deepest = max(contexts, key=len)
        rank = self.max_plate_nesting + len(guide_dims) + len(model_dims)
        pad = lambda t: t.reshape((1,) * (rank - t.dim()) + tuple(t.shape))
        elbo = 0.0
        for key, c in contexts.items():
            if key != deepest:
                tensors = c["model"] + c["guide"]
                assert all(t.dim() <= self.max_plate_nesting for t in tensors), "enumerated sites sit inside all plates"
                shape = torch.broadcast_shapes(*[t.shape for t in tensors])
                term = sum(t.expand(shape) for t in c["model"]) - sum(t.expand(shape) for t in c["guide"])
                elbo = elbo + c["scale"] * term.sum()
                continue
            joint = sum(pad(t) for t in c["model"])
            if model_dims:
                joint = torch.logsumexp(joint.expand(torch.broadcast_shapes(joint.shape, (1,) * rank)),
                                        dim=tuple(rank + d for d in model_dims), keepdim=True)
            cost = joint - sum(pad(t) for t in c["guide"])
            weight = sum(pad(t) for t in c["weights"]).exp() if c["weights"] else torch.ones(())
            shape = torch.broadcast_shapes(cost.shape, weight.shape)
            elbo = elbo + c["scale"] * (weight.expand(shape) * cost.expand(shape)).sum()
        return elbo

    def loss_and_grads(self, model, guide):
        loss = -self.differentiable_elbo(model, guide)
        loss.backward()
        return loss.item()


class Adam:
    """pyro.optim.Adam(optim_args): one torch.optim.Adam per parameter, created when the parameter is first seen."""

    def __init__(self, optim_args):
        self.args = dict(optim_args)
        self.args["betas"] = tuple(self.args.get("betas", (0.9, 0.999)))
        self.optims = {}

    def __call__(self, params):
        for p in params:
            if p not in self.optims:
                self.optims[p] = torch.optim.Adam([p], **self.args)
            self.optims[p].step()


class SVI:
    def __init__(self, model, guide, optim, loss):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss

    def step(self):
        with trace(param_only=True) as cap:
            loss = self.loss.loss_and_grads(self.model, self.guide)
        names = [k for k, s in cap.nodes.items() if s["type"] == "param"]
        params = [_PARAMS[k][0] for k in names]
        self.last_grads = OrderedDict((k, (_PARAMS[k][0].grad.detach().clone() if _PARAMS[k][0].grad is not None
                                           else torch.zeros_like(_PARAMS[k][0]))) for k in names)
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        self.optim(params)
        for p in params:
            p.grad = None
        return loss


# ---------------------------------------------------------------------------------------------------------------------
# module objects under the names the reference imports
# ---------------------------------------------------------------------------------------------------------------------
def install():
    """Registers ``pyro``, ``pyro.distributions``, ``pyro.distributions.util``, ``pyro.ops.indexing``, ``pyro.ops.stats`` and
    ``pyroapi`` in ``sys.modules``."""
    me = sys.modules[__name__]

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    dist = module("pyro.distributions", TorchDistribution=TorchDistribution, HalfNormal=HalfNormal, Dirichlet=Dirichlet,
                  Exponential=Exponential, Gamma=Gamma, Categorical=Categorical, Bernoulli=Bernoulli, Beta=Beta, Delta=Delta,
                  AffineBeta=AffineBeta, Independent=Independent)
    dist.__path__ = []
    module("pyro.distributions.util", broadcast_shape=broadcast_shape)
    ops = module("pyro.ops")
    ops.__path__ = []
    module("pyro.ops.indexing", Vindex=Vindex)
    module("pyro.ops.stats", quantile=None, hpdi=None)
    handlers = module("pyro.poutine", mask=mask, trace=trace, replay=replay, enum=enum, block=block, condition=condition,
                      uncondition=uncondition)
    infer = module("pyro.infer", TraceEnum_ELBO=TraceEnum_ELBO, JitTraceEnum_ELBO=TraceEnum_ELBO, SVI=SVI, Predictive=Predictive)
    optim = module("pyro.optim", Adam=Adam)
    pyro = module("pyro", sample=sample, param=param, plate=plate, markov=markov, clear_param_store=clear_param_store,
                  get_param_store=get_param_store, set_rng_seed=set_rng_seed, distributions=dist, poutine=handlers,
                  infer=infer, optim=optim, ops=ops)
    pyro.__path__ = []
    module("pyroapi", distributions=dist, handlers=handlers, infer=infer, optim=optim, pyro=pyro)
    return me
